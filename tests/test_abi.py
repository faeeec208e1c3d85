"""The C-ABI libraries load without a GPU and export every symbol their headers declare (no compute calls)."""
import ctypes as C
import os
import re

from solr_b200 import engine, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "solr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-zA-Z0-9_]+)\s*\(", text)))


def test_engine_exports_every_declared_symbol():
    lib = engine.load()
    decl = declared_symbols()
    assert len(decl) >= 20
    for s in decl:
        assert hasattr(lib, s), "missing export: " + s
    assert sorted(engine.ABI_SYMBOLS) == decl, "engine.py's symbol list is out of date with include/solr_b200.h"


def test_seam_mirrors_the_reference_ten():
    # CudaRayTracer.h:25-67 has exactly these ten entry points (cudaRender -> b200_render)
    ref = ["initialize_scene", "finalize_scene", "reshape_scene", "h2d_scene", "h2d_materials", "h2d_randoms",
           "h2d_textures", "h2d_lightInformation", "d2h_bitmap", "render"]
    decl = declared_symbols()
    for r in ref:
        assert "b200_" + r in decl


def test_host_library_exports():
    lib = host.load()
    for s in host.ABI_SYMBOLS:
        assert hasattr(lib, s), "missing export: " + s


def test_error_is_latched_not_raised_across_the_seam():
    lib = engine.load()
    lib.b200_clear_error()
    lib.b200_set_partition(3, 2)  # invalid: rank >= world; pure host-side validation, no CUDA call
    buf = C.create_string_buffer(256)
    assert lib.b200_last_error(buf, 256) != 0 and b"rank" in buf.value
    lib.b200_clear_error()
    assert lib.b200_last_error(buf, 256) == 0
    lib.b200_set_partition(0, 1)
