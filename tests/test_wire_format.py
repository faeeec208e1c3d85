"""The wire format at the engine seam: include/solr_b200_types.h vs the ctypes mirrors vs the reference."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import refh
from solr_b200 import wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ctypes_sizes_match_survey_appendix_a():
    for t, n in wire.WIRE_SIZES.items():
        assert C.sizeof(t) == n


def test_header_compiles_as_c_and_cpp_with_static_asserts():
    for lang, cc, std in (("c", "gcc", "-std=c11"), ("c++", "g++", "-std=c++14")):
        with tempfile.TemporaryDirectory() as d:
            src = os.path.join(d, "t.c" if lang == "c" else "t.cpp")
            with open(src, "w") as f:
                f.write('#include "solr_b200.h"\nint main(void){return (int)sizeof(b200_SceneInfo) - 112;}\n')
            exe = os.path.join(d, "t")
            subprocess.check_call(["/usr/bin/" + cc, std, "-I", os.path.join(ROOT, "include"), src, "-o", exe])
            assert subprocess.call([exe]) == 0


@pytest.mark.skipif(not refh.available("cpu"), reason="reference not built (oracle/_ref)")
def test_sizes_match_compiled_reference():
    out = (C.c_int * 10)()
    refh.load("cpu").refh_struct_sizes(out)
    sizes = list(out)
    assert sizes[:9] == [112, 48, 128, 176, 48, 16, 32, 16, 32]


@pytest.mark.skipif(not refh.available("cpu"), reason="reference not built (oracle/_ref)")
def test_field_offsets_against_reference_arrays():
    """Decode the reference's flattened arrays with our mirrors: the fields land where the setters put them."""
    import golden_scenes as gs
    sc, si, *_ = gs.case_setup("mixed_full")
    r = refh.RefScene(si, "cpu")
    sc.replay(r)
    a = r.arrays()
    prims = np.frombuffer(a["primitives"].tobytes(), dtype=np.dtype([
        ("p0", "3f4"), ("p1", "3f4"), ("p2", "3f4"), ("n0", "3f4"), ("n1", "3f4"), ("n2", "3f4"), ("size", "3f4"),
        ("type", "i4"), ("index", "i4"), ("materialId", "i4"), ("vt", "6f4"), ("pad", "2i4")]))
    assert prims.dtype.itemsize == 128
    assert sorted(prims["index"].tolist()) == list(range(sc.nb_primitives))
    for p in prims:
        i = int(p["index"])
        assert p["type"] == sc.prim_type[i] and p["materialId"] == sc.prim_mat[i]
        assert np.allclose(p["p0"], sc.prim_v[i, 0:3])
    r.close()
