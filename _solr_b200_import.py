"""Registers the package directory `sol-r_b200/` (not a valid Python identifier) as module `solr_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(_ROOT, "sol-r_b200")


def _load():
    if "solr_b200" in sys.modules:
        return sys.modules["solr_b200"]
    spec = importlib.util.spec_from_file_location("solr_b200", os.path.join(_PKG, "__init__.py"),
                                                  submodule_search_locations=[_PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["solr_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


solr_b200 = _load()
