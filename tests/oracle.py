"""ctypes binding of oracle/libsolr_oracle.so — the CPU restatement (checker).  Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

from _solr_b200_import import solr_b200  # noqa: F401
from solr_b200 import wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libsolr_oracle.so")

COUNTER_FIELDS = ["pixels", "rays", "primary_rays", "shadow_rays", "box_tests", "sphere_tests", "cylinder_tests",
                  "cone_tests", "triangle_tests", "plane_tests", "ellipsoid_tests", "accepted_hits", "shade_calls"]


class Counters(C.Structure):
    _fields_ = [(f, C.c_uint64) for f in COUNTER_FIELDS]

    def as_dict(self):
        return {f: int(getattr(self, f)) for f in COUNTER_FIELDS}


class OracleScene(C.Structure):
    _fields_ = [("boxes", C.c_void_p), ("nbBoxes", C.c_int), ("primitives", C.c_void_p), ("nbPrimitives", C.c_int),
                ("materials", C.c_void_p), ("nbMaterials", C.c_int), ("lightInformation", C.c_void_p),
                ("lightInformationSize", C.c_int), ("nbLamps", C.c_int), ("textures", C.c_void_p),
                ("randoms", C.c_void_p), ("randomTableSize", C.c_int)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libsolr_oracle.so"])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oracle_render.argtypes = [C.POINTER(OracleScene), C.POINTER(wire.SceneInfo),
                                       C.POINTER(wire.PostProcessingInfo), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(Counters)]
        _lib.oracle_post_process.argtypes = [C.POINTER(OracleScene), C.POINTER(wire.SceneInfo), C.POINTER(wire.PostProcessingInfo),
                                             C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_algorithmic_flops.argtypes = [C.POINTER(Counters)]
        _lib.oracle_algorithmic_flops.restype = C.c_double
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """Holds the flattened scene arrays (dict as produced by the host scene container or the reference
    harness: uint8 views of BoundingBox[], Primitive[], Material[], LightInformation[]) and the
    progressive per-pixel state."""

    def __init__(self, arrays, width, height, randoms=None, textures=None, random_table_size=wire.REF_MAX_BITMAP_SIZE):
        self.lib = load()
        self.a = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in arrays.items()}
        self.W, self.H = width, height
        self.randoms = np.zeros(random_table_size, np.float32) if randoms is None else np.ascontiguousarray(randoms, np.float32)
        self.random_table_size = random_table_size
        self.textures = textures
        self.post = np.zeros((height, width, 8), np.float32)
        self.ids = np.zeros((height, width, 4), np.int32)
        self.bitmap = np.zeros((height, width, 3), np.uint8)

    def render(self, scene_info, eye, target, angles, post_info=None, rows=None, threads=None):
        post_info = post_info or wire.PostProcessingInfo()
        s = OracleScene(_ptr(self.a["boxes"]), self.a["nbBoxes"], _ptr(self.a["primitives"]), self.a["nbPrimitives"],
                        _ptr(self.a["materials"]), self.a["nbMaterials"], _ptr(self.a["lightInformation"]),
                        self.a["lightInformationSize"], self.a["nbLamps"], _ptr(self.textures), _ptr(self.randoms),
                        self.random_table_size)
        e = np.asarray(eye, np.float32); t = np.asarray(target, np.float32); a = np.asarray(angles, np.float32)
        r0, r1, rs = rows if rows is not None else (0, self.H, 1)
        k = Counters()
        self.lib.oracle_render(C.byref(s), C.byref(scene_info), C.byref(post_info), _ptr(e), _ptr(t), _ptr(a),
                               _ptr(self.post), _ptr(self.ids), _ptr(self.bitmap), r0, r1, rs,
                               threads or os.cpu_count() or 1, C.byref(k))
        if post_info.type != 0:
            assert rows is None, "post-processing effects need the whole frame"
            self.lib.oracle_post_process(C.byref(s), C.byref(scene_info), C.byref(post_info), _ptr(self.post), _ptr(self.ids),
                                         _ptr(self.bitmap))
        self.counters = k
        return self.bitmap, self.ids, self.post, k

    def flops(self, counters=None):
        return float(self.lib.oracle_algorithmic_flops(C.byref(counters or self.counters)))
