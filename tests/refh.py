"""ctypes binding of oracle/_ref/libsolr_ref_{cpu,cuda}.so (the UNMODIFIED reference, built by
oracle/ref_build/Makefile).  Test infrastructure only — the product never loads these libraries."""
import ctypes as C
import os

import numpy as np

from _solr_b200_import import solr_b200  # noqa: F401
from solr_b200 import wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


class RefhScene(C.Structure):
    _fields_ = [("boxes", C.c_void_p), ("nbBoxes", C.c_int), ("primitives", C.c_void_p), ("nbPrimitives", C.c_int),
                ("materials", C.c_void_p), ("nbMaterials", C.c_int), ("lightInformation", C.c_void_p),
                ("lightInformationSize", C.c_int), ("lamps", C.c_void_p), ("nbLamps", C.c_int),
                ("bounds", C.c_float * 6)]


def lib_path(kind="cpu"):
    return os.path.join(REF_DIR, "libsolr_ref_%s.so" % kind)


def available(kind="cpu"):
    return os.path.exists(lib_path(kind))


_libs = {}


def load(kind="cpu"):
    if kind not in _libs:
        lib = C.CDLL(lib_path(kind))
        lib.refh_create.restype = C.c_void_p
        lib.refh_create.argtypes = [C.POINTER(wire.SceneInfo)]
        lib.refh_destroy.argtypes = [C.c_void_p]
        lib.refh_add_materials.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.refh_add_primitives.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.refh_set_primitive_normals.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.refh_set_normals_bulk.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.refh_set_texture.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.refh_set_material_raw.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.refh_compact_boxes.argtypes = [C.c_void_p]
        lib.refh_compact_boxes.restype = C.c_int
        lib.refh_get_scene.argtypes = [C.c_void_p, C.POINTER(RefhScene)]
        lib.refh_render.argtypes = [C.c_void_p, C.POINTER(wire.SceneInfo), C.POINTER(wire.PostProcessingInfo),
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
        lib.refh_struct_sizes.argtypes = [C.c_void_p]
        lib.refh_limits.argtypes = [C.c_int]
        lib.refh_limits.restype = C.c_int
        _libs[kind] = lib
    return _libs[kind]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RefScene:
    """The reference's own scene container + engine, driven through its seam."""

    def __init__(self, scene_info, kind="cpu", limits=None):
        """limits = (max width, max height): only the drop-in build ("b200") takes frames beyond the reference's 1920x1080."""
        self.lib = load(kind)
        if limits is not None:
            assert kind == "b200", "the reference's own engines are limited to MAX_BITMAP_WIDTH x MAX_BITMAP_HEIGHT"
            self.lib.refh_create_limits.restype = C.c_void_p
            self.lib.refh_create_limits.argtypes = [C.POINTER(wire.SceneInfo), C.c_int, C.c_int]
            self.h = self.lib.refh_create_limits(C.byref(scene_info), limits[0], limits[1])
        else:
            self.h = self.lib.refh_create(C.byref(scene_info))

    def close(self):
        if self.h:
            self.lib.refh_destroy(self.h)
            self.h = None

    # builder protocol (scenes.Scene.replay)
    def add_materials(self, mat_f, mat_i):
        mat_f = np.ascontiguousarray(mat_f, np.float32); mat_i = np.ascontiguousarray(mat_i, np.int32)
        self.lib.refh_add_materials(self.h, mat_f.shape[0], _ptr(mat_f), _ptr(mat_i))

    def add_primitives(self, t, v, m):
        t = np.ascontiguousarray(t, np.int32); v = np.ascontiguousarray(v, np.float32)
        m = np.ascontiguousarray(m, np.int32)
        self.lib.refh_add_primitives(self.h, t.shape[0], _ptr(t), _ptr(v), _ptr(m))

    def set_normals(self, idx, n):
        n = np.ascontiguousarray(n, np.float32)
        self.lib.refh_set_primitive_normals(self.h, idx, _ptr(n))

    def set_normals_bulk(self, first, normals):
        normals = np.ascontiguousarray(normals, np.float32)
        self.lib.refh_set_normals_bulk(self.h, first, normals.shape[0], _ptr(normals))

    def set_texture(self, index, texels):
        t = np.ascontiguousarray(texels, np.uint8)
        self._tex = getattr(self, "_tex", []) + [t]
        self.lib.refh_set_texture(self.h, index, _ptr(t), t.shape[1], t.shape[0], t.shape[2])

    def compact_boxes(self, reconstruct=True):
        if reconstruct:
            return self.lib.refh_compact_boxes(self.h)
        self.lib.refh_compact_boxes_mode.argtypes = [C.c_void_p, C.c_int]
        self.lib.refh_compact_boxes_mode.restype = C.c_int
        return self.lib.refh_compact_boxes_mode(self.h, 0)

    def rotate_primitives(self, center, angles):
        c = (C.c_float * 3)(*center); a = (C.c_float * 3)(*angles)
        self.lib.refh_rotate_primitives.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.refh_rotate_primitives.restype = None
        self.lib.refh_rotate_primitives(self.h, c, a)

    # integration/B200Kernel only: the step applied on the device (True) or by the host fallback (False)
    def rotate_primitives_on_device(self, center, angles):
        c = np.asarray(center, np.float32); a = np.asarray(angles, np.float32)
        self.lib.refh_rotate_primitives_on_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.refh_rotate_primitives_on_device.restype = C.c_int
        return bool(self.lib.refh_rotate_primitives_on_device(self.h, _ptr(c), _ptr(a)))

    def translate_primitives_on_device(self, t):
        v = np.asarray(t, np.float32)
        self.lib.refh_translate_primitives_on_device.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.refh_translate_primitives_on_device.restype = C.c_int
        return bool(self.lib.refh_translate_primitives_on_device(self.h, _ptr(v)))

    def sync_from_device(self):
        self.lib.refh_sync_from_device.argtypes = [C.c_void_p]
        self.lib.refh_sync_from_device(self.h)

    def translate_primitives(self, t):
        v = (C.c_float * 3)(*t)
        self.lib.refh_translate_primitives.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.refh_translate_primitives.restype = None
        self.lib.refh_translate_primitives(self.h, v)

    def scale_primitives(self, scale):
        self.lib.refh_scale_primitives.argtypes = [C.c_void_p, C.c_float]
        self.lib.refh_scale_primitives.restype = None
        self.lib.refh_scale_primitives(self.h, scale)

    def arrays(self):
        """Copies of the flattened wire-format arrays the reference engine would receive."""
        s = RefhScene()
        self.lib.refh_get_scene(self.h, C.byref(s))

        def grab(ptr, n, size):
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * size,)).copy()

        return {
            "boxes": grab(s.boxes, s.nbBoxes, 48), "nbBoxes": s.nbBoxes,
            "primitives": grab(s.primitives, s.nbPrimitives, 128), "nbPrimitives": s.nbPrimitives,
            "materials": grab(s.materials, s.nbMaterials, 176), "nbMaterials": s.nbMaterials,
            "lightInformation": grab(s.lightInformation, max(s.lightInformationSize, 1), 48)[: s.lightInformationSize * 48],
            "lightInformationSize": s.lightInformationSize,
            "lamps": np.ctypeslib.as_array(C.cast(s.lamps, C.POINTER(C.c_int)), shape=(max(s.nbLamps, 1),)).copy()[: s.nbLamps],
            "nbLamps": s.nbLamps, "bounds": np.array(list(s.bounds), dtype=np.float32),
        }

    def render(self, scene_info, eye, target, angles, randoms=None, post_info=None, block=(16, 16),
               want_post=True):
        W, H = scene_info.size.x, scene_info.size.y
        if randoms is None:
            randoms = np.zeros(wire.REF_MAX_BITMAP_SIZE, np.float32)
        assert randoms.shape[0] >= wire.REF_MAX_BITMAP_SIZE
        post_info = post_info or wire.PostProcessingInfo()
        bitmap = np.zeros(W * H * 3, np.uint8)
        ids = np.zeros((W * H, 4), np.int32)
        post = np.zeros((W * H, 8), np.float32) if want_post else None
        e = np.asarray(eye, np.float32); t = np.asarray(target, np.float32); a = np.asarray(angles, np.float32)
        b = np.asarray(block, np.int32)
        self.lib.refh_render(self.h, C.byref(scene_info), C.byref(post_info), _ptr(e), _ptr(t), _ptr(a),
                             _ptr(randoms), _ptr(b), _ptr(bitmap), _ptr(ids), _ptr(post) if want_post else None)
        return bitmap.reshape(H, W, 3), ids.reshape(H, W, 4), (post.reshape(H, W, 8) if want_post else None)
