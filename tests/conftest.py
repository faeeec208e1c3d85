import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


from _solr_b200_import import solr_b200  # noqa: E402,F401  (registers the package dir `sol-r_b200/` as `solr_b200`)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_libs():
    """Build the in-tree libraries when missing (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build(quiet=True)
    yield
