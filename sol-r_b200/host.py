"""ctypes binding of csrc/libsolr_b200_host.so — SceneHost, the host-side mirror of the reference's
GPUKernel (setters, box compaction) and CudaKernel (render_begin / render_end) for the hot path."""
import ctypes as C
import os

import numpy as np

from . import wire

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libsolr_b200_host.so")
_lib = None


class HostScene(C.Structure):
    _fields_ = [("boxes", C.c_void_p), ("nbBoxes", C.c_int), ("primitives", C.c_void_p), ("nbPrimitives", C.c_int),
                ("materials", C.c_void_p), ("nbMaterials", C.c_int), ("lightInformation", C.c_void_p),
                ("lightInformationSize", C.c_int), ("lamps", C.c_void_p), ("nbLamps", C.c_int),
                ("bounds", C.c_float * 6), ("treeDepth", C.c_int)]


ABI_SYMBOLS = ["b200h_create", "b200h_destroy", "b200h_set_scene_info", "b200h_set_post_processing_info", "b200h_set_camera",
               "b200h_add_primitive", "b200h_set_primitive", "b200h_add_primitives", "b200h_set_normals_bulk",
               "b200h_set_texcoords", "b200h_add_material", "b200h_add_materials", "b200h_set_material_raw",
               "b200h_set_texture", "b200h_compact_boxes", "b200h_get_scene", "b200h_set_randoms", "b200h_set_limits", "b200h_set_capacity",
               "b200h_set_partition", "b200h_set_device", "b200h_init_buffers", "b200h_render_begin", "b200h_render_end",
               "b200h_get_bitmap", "b200h_get_primitive_ids", "b200h_get_primitive_at", "b200h_set_lazy_ids", "b200h_set_flat_build",
               "b200h_rotate_primitives", "b200h_translate_primitives", "b200h_scale_primitives",
               "b200h_set_device_animation", "b200h_sync_from_device", "b200h_find_bonds", "b200h_share_frame", "b200h_obj_vertex_pass"]


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("host library missing: %s (run __graft_entry__.build())" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.b200h_create.restype = vp
    lib.b200h_create.argtypes = [C.POINTER(wire.SceneInfo)]
    lib.b200h_destroy.argtypes = [vp]
    lib.b200h_set_scene_info.argtypes = [vp, C.POINTER(wire.SceneInfo)]
    lib.b200h_set_post_processing_info.argtypes = [vp, C.POINTER(wire.PostProcessingInfo)]
    lib.b200h_set_camera.argtypes = [vp, vp, vp, vp]
    lib.b200h_add_primitive.argtypes = [vp, C.c_int]
    lib.b200h_set_primitive.argtypes = [vp, C.c_int, vp, C.c_int]
    lib.b200h_add_primitives.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.b200h_set_normals_bulk.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.b200h_set_texcoords.argtypes = [vp, C.c_int, vp]
    lib.b200h_add_material.argtypes = [vp]
    lib.b200h_add_materials.argtypes = [vp, C.c_int, vp, vp]
    lib.b200h_set_material_raw.argtypes = [vp, C.c_int, vp]
    lib.b200h_set_texture.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int]
    lib.b200h_compact_boxes.argtypes = [vp, C.c_int]
    lib.b200h_compact_boxes.restype = C.c_int
    lib.b200h_get_scene.argtypes = [vp, C.POINTER(HostScene)]
    lib.b200h_set_randoms.argtypes = [vp, vp, C.c_long, C.c_int]
    lib.b200h_set_limits.argtypes = [vp, C.c_int, C.c_int]
    lib.b200h_set_capacity.argtypes = [vp, C.c_long, C.c_long]
    lib.b200h_set_partition.argtypes = [vp, C.c_int, C.c_int]
    lib.b200h_rotate_primitives.argtypes = [vp, vp, vp]
    lib.b200h_rotate_primitives.restype = None
    lib.b200h_translate_primitives.argtypes = [vp, vp]
    lib.b200h_translate_primitives.restype = None
    lib.b200h_scale_primitives.argtypes = [vp, C.c_float]
    lib.b200h_scale_primitives.restype = None
    lib.b200h_set_flat_build.argtypes = [vp, C.c_int]
    lib.b200h_set_device_animation.argtypes = [vp, C.c_int]
    lib.b200h_sync_from_device.argtypes = [vp]
    lib.b200h_set_flat_build.restype = None
    lib.b200h_set_lazy_ids.argtypes = [vp, C.c_int]
    lib.b200h_set_lazy_ids.restype = None
    lib.b200h_set_device.argtypes = [vp, C.c_int]
    lib.b200h_init_buffers.argtypes = [vp]
    lib.b200h_share_frame.argtypes = [vp, C.c_char_p, C.c_int]
    lib.b200h_share_frame.restype = C.c_int
    lib.b200h_render_begin.argtypes = [vp, C.c_float]
    lib.b200h_render_end.argtypes = [vp]
    lib.b200h_get_bitmap.argtypes = [vp]
    lib.b200h_get_bitmap.restype = vp
    lib.b200h_get_primitive_ids.argtypes = [vp]
    lib.b200h_get_primitive_ids.restype = vp
    lib.b200h_get_primitive_at.argtypes = [vp, C.c_int, C.c_int]
    lib.b200h_get_primitive_at.restype = C.c_uint
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def obj_vertex_pass(text):
    """b200h_obj_vertex_pass (csrc/loaders.cpp): the first pass of OBJReader::loadModelFromFile (OBJReader.cpp:440-563) over the text
    of a .obj file.  Returns (vertices float32[n, 3], normals float32[m, 3], tex_coords float32[k, 2], aabb float32[6]); entry i of
    an array is the reference's map entry i + 1."""
    lib = load()
    lib.b200h_obj_vertex_pass.restype = C.c_int
    lib.b200h_obj_vertex_pass.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    data = text if isinstance(text, bytes) else text.encode("latin-1")
    counts = np.zeros(3, np.int32)
    aabb = np.zeros(6, np.float32)
    lib.b200h_obj_vertex_pass(data, len(data), None, 0, None, 0, None, 0, _ptr(counts), _ptr(aabb))
    v = np.zeros((int(counts[0]), 3), np.float32); n = np.zeros((int(counts[1]), 3), np.float32); t = np.zeros((int(counts[2]), 2), np.float32)
    lib.b200h_obj_vertex_pass(data, len(data), _ptr(v), v.shape[0], _ptr(n), n.shape[0], _ptr(t), t.shape[0], _ptr(counts), _ptr(aabb))
    return v, n, t, aabb


def find_bonds(xyz, processed, is_backbone, stick_distance=1.7, backbone_geometry=False):
    """The bond search of PDBReader.cpp:616-660 without its n^2 loop (csrc/loaders.cpp): for atoms in std::map order, (first, partners)
    with partners[first[i]:first[i + 1]] = the atoms atom i gets a stick to, ascending."""
    lib = load()
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    n = xyz.shape[0]
    processed = np.ascontiguousarray(processed, np.int32)
    is_backbone = np.ascontiguousarray(is_backbone, np.uint8)
    first = np.zeros(n + 1, np.int32)
    lib.b200h_find_bonds.restype = C.c_long
    lib.b200h_find_bonds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_long]
    need = lib.b200h_find_bonds(_ptr(xyz), _ptr(processed), _ptr(is_backbone), n, 1 if backbone_geometry else 0, stick_distance, _ptr(first), None, 0)
    need = abs(int(need))
    partners = np.zeros(max(need, 1), np.int32)
    got = lib.b200h_find_bonds(_ptr(xyz), _ptr(processed), _ptr(is_backbone), n, 1 if backbone_geometry else 0, stick_distance, _ptr(first), _ptr(partners), need)
    assert got == need
    return first, partners[:need]


class SceneHost:
    """Same call sequence a Sol-R application makes on solr::GPUKernel."""

    def __init__(self, scene_info, limits=None, rank=0, world=1, device=None, capacity=None):
        """capacity = (max boxes, max primitives) of the flattened arrays; None = the reference's 2.5 M each."""
        self.lib = load()
        self.scene_info = scene_info
        self.h = self.lib.b200h_create(C.byref(scene_info))
        self.limits = limits or (1920, 1080)
        self.lib.b200h_set_limits(self.h, self.limits[0], self.limits[1])
        self.lib.b200h_set_partition(self.h, rank, world)
        if capacity is not None:
            self.lib.b200h_set_capacity(self.h, capacity[0], capacity[1])
        if device is not None:
            self.lib.b200h_set_device(self.h, device)

    def close(self):
        if self.h:
            self.lib.b200h_destroy(self.h)
            self.h = None

    # ---- builder protocol (scenes.Scene.replay) ----
    def add_materials(self, mat_f, mat_i):
        mat_f = np.ascontiguousarray(mat_f, np.float32); mat_i = np.ascontiguousarray(mat_i, np.int32)
        self.lib.b200h_add_materials(self.h, mat_f.shape[0], _ptr(mat_f), _ptr(mat_i))

    def add_primitives(self, t, v, m):
        t = np.ascontiguousarray(t, np.int32); v = np.ascontiguousarray(v, np.float32); m = np.ascontiguousarray(m, np.int32)
        self.lib.b200h_add_primitives(self.h, t.shape[0], _ptr(t), _ptr(v), _ptr(m))

    def set_normals(self, idx, n):
        n = np.ascontiguousarray(n, np.float32).reshape(1, 9)
        self.lib.b200h_set_normals_bulk(self.h, idx, 1, _ptr(n))

    def set_normals_bulk(self, first, normals):
        normals = np.ascontiguousarray(normals, np.float32)
        self.lib.b200h_set_normals_bulk(self.h, first, normals.shape[0], _ptr(normals))

    def set_texture(self, index, texels):
        t = np.ascontiguousarray(texels, np.uint8)
        self.lib.b200h_set_texture(self.h, index, _ptr(t), t.shape[1], t.shape[0], t.shape[2])

    def compact_boxes(self, reconstruct=True):
        return self.lib.b200h_compact_boxes(self.h, 1 if reconstruct else 0)

    def arrays(self):
        s = HostScene()
        self.lib.b200h_get_scene(self.h, C.byref(s))

        def grab(ptr, n, size):
            if n == 0 or not ptr:
                return np.zeros(0, np.uint8)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * size,)).copy()

        lamps = np.ctypeslib.as_array(C.cast(s.lamps, C.POINTER(C.c_int)), shape=(s.nbLamps,)).copy() if s.nbLamps and s.lamps else np.zeros(0, np.int32)
        return {"boxes": grab(s.boxes, s.nbBoxes, 48), "nbBoxes": s.nbBoxes,
                "primitives": grab(s.primitives, s.nbPrimitives, 128), "nbPrimitives": s.nbPrimitives,
                "materials": grab(s.materials, s.nbMaterials, 176), "nbMaterials": s.nbMaterials,
                "lightInformation": grab(s.lightInformation, s.lightInformationSize, 48),
                "lightInformationSize": s.lightInformationSize, "lamps": lamps, "nbLamps": s.nbLamps,
                "bounds": np.array(list(s.bounds), np.float32), "treeDepth": s.treeDepth}

    # ---- frame protocol ----
    def set_scene_info(self, si):
        self.scene_info = si
        self.lib.b200h_set_scene_info(self.h, C.byref(si))

    def set_camera(self, eye, target, angles):
        e = np.asarray(eye, np.float32); t = np.asarray(target, np.float32); a = np.asarray(angles, np.float32)
        self.lib.b200h_set_camera(self.h, _ptr(e), _ptr(t), _ptr(a))

    def set_randoms(self, randoms, timestamp=0):
        r = np.ascontiguousarray(randoms, np.float32)
        self.lib.b200h_set_randoms(self.h, _ptr(r), r.shape[0], timestamp)

    def init_buffers(self):
        self.lib.b200h_init_buffers(self.h)

    def share_frame(self, name, create):
        """SceneHost::shareFrame: the frame and id buffers move into POSIX shared memory `name`; every process's GPU writes its own
        tiles there (partition.SharedHostFrame does the naming and the ordering between processes)."""
        rc = self.lib.b200h_share_frame(self.h, name.encode(), 1 if create else 0)
        if rc != 0:
            raise RuntimeError("b200h_share_frame(%s) failed: %d" % (name, rc))

    def render_begin(self, timer=0.0):
        self.lib.b200h_render_begin(self.h, timer)

    def render_end(self):
        self.lib.b200h_render_end(self.h)

    # animation step (GPUKernel::rotatePrimitives / translatePrimitives / scalePrimitives, then compact_boxes(False))
    def rotate_primitives(self, center, angles):
        c = (C.c_float * 3)(*center); a = (C.c_float * 3)(*angles)
        self.lib.b200h_rotate_primitives(self.h, c, a)

    def translate_primitives(self, t):
        v = (C.c_float * 3)(*t)
        self.lib.b200h_translate_primitives(self.h, v)

    def scale_primitives(self, scale):
        self.lib.b200h_scale_primitives(self.h, scale)

    def set_device_animation(self, on=True):
        """rotate_primitives / translate_primitives move the scene on the device once it is there (csrc/animate.cuh); this container's
        own copy is brought up to date when somebody asks for it (arrays(), any later host-side step)."""
        self.lib.b200h_set_device_animation(self.h, 1 if on else 0)

    def sync_from_device(self):
        self.lib.b200h_sync_from_device(self.h)

    def set_flat_build(self, mode):
        """0 / False: compact_boxes always builds the reference's per-level maps literally; 1: flat sort-and-merge build of a fresh
        container, flattened depth first; 2 / True (default): the same, flattened level by level.  Same arrays; tests compare them."""
        self.lib.b200h_set_flat_build(self.h, 2 if mode is True else int(mode))

    def set_lazy_ids(self, lazy):
        """True (default): render_end copies the pixels only and the id buffer stays on the device until primitive_ids() /
        get_primitive_at() ask for it; False: the reference's protocol, both buffers every frame (CudaKernel.cpp:304-313)."""
        self.lib.b200h_set_lazy_ids(self.h, 1 if lazy else 0)

    def bitmap(self):
        W, H = self.scene_info.size.x, self.scene_info.size.y
        p = self.lib.b200h_get_bitmap(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(H, W, 3))

    def primitive_ids(self):
        W, H = self.scene_info.size.x, self.scene_info.size.y
        p = self.lib.b200h_get_primitive_ids(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int)), shape=(H, W, 4))

    def get_primitive_at(self, x, y):
        return int(self.lib.b200h_get_primitive_at(self.h, x, y))
