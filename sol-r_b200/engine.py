"""ctypes binding of the engine's C ABI (include/solr_b200.h) — the calls the reference's engine host
class makes across its seam (/root/reference/solr/engines/cuda/CudaKernel.cpp:116-145 initializeDevice,
:174-302 render_begin, :304-313 render_end).  There is no CPU fallback: a missing library or a missing
GPU raises."""
import ctypes as C
import os

import numpy as np

from . import wire

HERE = os.path.dirname(os.path.abspath(__file__))
# SOLR_B200_LIB: an experimental build of the engine library (tools/gpu sweeps); the product path is the in-tree default
LIB_PATH = os.environ.get("SOLR_B200_LIB") or os.path.join(HERE, "csrc", "libsolr_b200.so")

_lib = None


class EngineError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError("CUDA engine library missing: %s (run __graft_entry__.build(); there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    I2, I4, F3, F4 = wire.Int2, wire.Int4, wire.Float3, wire.Float4
    SI, PPI = wire.SceneInfo, wire.PostProcessingInfo
    lib.b200_initialize_scene.argtypes = [I2, SI, C.c_int, C.c_int, C.c_int]
    lib.b200_finalize_scene.argtypes = [I2]
    lib.b200_reshape_scene.argtypes = [I2, SI]
    lib.b200_h2d_scene.argtypes = [I2, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.b200_h2d_materials.argtypes = [I2, C.c_void_p, C.c_int]
    lib.b200_h2d_randoms.argtypes = [I2, C.c_void_p]
    lib.b200_h2d_textures.argtypes = [I2, C.c_int, C.c_void_p]
    lib.b200_h2d_lightInformation.argtypes = [I2, C.c_void_p, C.c_int]
    lib.b200_d2h_bitmap.argtypes = [I2, SI, C.c_void_p, C.c_void_p]
    lib.b200_render.argtypes = [I2, I4, SI, I4, PPI, F3, F3, F4]
    lib.b200_last_error.argtypes = [C.c_char_p, C.c_int]
    lib.b200_last_error.restype = C.c_int
    lib.b200_set_device.argtypes = [C.c_int]
    lib.b200_set_stream.argtypes = [C.c_void_p]
    lib.b200_set_limits.argtypes = [C.c_int, C.c_int]
    lib.b200_set_option.argtypes = [C.c_int, C.c_int]
    lib.b200_set_option.restype = None
    lib.b200_set_partition.argtypes = [C.c_int, C.c_int]
    lib.b200_device_buffers.argtypes = [C.POINTER(C.c_void_p)] * 3
    lib.b200_d2h_primitive_id.argtypes = [SI, C.c_int, C.c_int, C.c_void_p]
    lib.b200_d2h_primitive_id.restype = None
    lib.b200_peer_frame_export.argtypes = [C.c_void_p, C.c_int]
    lib.b200_peer_frame_export.restype = C.c_int
    lib.b200_peer_frame_open.argtypes = [C.c_void_p, C.c_int]
    lib.b200_peer_frame_open.restype = C.c_int
    lib.b200_get_counters.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.c_int]
    lib.b200_d2h_post.argtypes = [SI, C.c_void_p]
    lib.b200_d2h_post.restype = None
    lib.b200_debug_relayout_boxes.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.b200_debug_relayout_boxes.restype = C.c_int
    lib.b200_frame_parameter_bytes.restype = C.c_int
    lib.b200_debug_build_unordered.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.b200_debug_build_unordered.restype = C.c_int
    lib.b200_debug_counters.argtypes = [C.c_void_p]
    lib.b200_debug_counters.restype = None
    lib.b200_last_render_ms.restype = C.c_float
    lib.b200_kernel_launches.restype = C.c_ulonglong
    lib.b200_scene_stats.argtypes = [C.POINTER(C.c_int)] * 4
    lib.b200_measure_fp32_peak.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.b200_measure_fp32_peak.restype = C.c_int
    lib.b200_accumulation_clear.restype = C.c_int
    lib.b200_register_host.argtypes = [C.c_void_p, C.c_size_t]
    lib.b200_register_host.restype = C.c_int
    lib.b200_unregister_host.argtypes = [C.c_void_p]
    lib.b200_unregister_host.restype = C.c_int
    lib.b200_frames_streamed.restype = C.c_ulonglong
    lib.b200_stream_target.argtypes = [SI, C.c_void_p, C.c_void_p]
    lib.b200_stream_target.restype = C.c_int
    lib.b200_accumulation_export.argtypes = [C.c_void_p]
    lib.b200_accumulation_export.restype = C.c_int
    lib.b200_accumulation_import_and_pack.argtypes = [C.c_void_p, C.c_int]
    lib.b200_accumulation_import_and_pack.restype = C.c_int
    for f in ("b200_initialize_scene", "b200_finalize_scene", "b200_reshape_scene", "b200_h2d_scene", "b200_h2d_materials",
              "b200_h2d_randoms", "b200_h2d_textures", "b200_h2d_lightInformation", "b200_d2h_bitmap", "b200_render",
              "b200_set_device", "b200_set_option", "b200_set_stream", "b200_set_limits", "b200_set_partition", "b200_device_buffers",
              "b200_get_counters", "b200_scene_stats", "b200_scene_upload_stats", "b200_synchronize", "b200_clear_error"):
        getattr(lib, f).restype = None
    lib.b200_rotate_primitives.argtypes = [wire.Float3, wire.Float3]
    lib.b200_translate_primitives.argtypes = [wire.Float3]
    lib.b200_scale_primitives.argtypes = [C.c_float]
    lib.b200_d2h_scene.argtypes = [C.c_void_p, C.c_void_p]
    for f in ("b200_rotate_primitives", "b200_translate_primitives", "b200_scale_primitives", "b200_d2h_scene"):
        getattr(lib, f).restype = C.c_int
    lib.b200_last_animation_ms.restype = C.c_float
    lib.b200_scene_layout.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    lib.b200_scene_adopt_layout.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    lib.b200_scene_device_arrays.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.c_int]
    for f in ("b200_scene_layout", "b200_scene_adopt_layout", "b200_scene_device_arrays", "b200_scene_adopt_finish"):
        getattr(lib, f).restype = C.c_int
    _lib = lib
    return lib


# every symbol include/solr_b200.h declares (checked by tests/test_abi.py without a GPU)
ABI_SYMBOLS = [
    "b200_initialize_scene", "b200_finalize_scene", "b200_reshape_scene", "b200_h2d_scene", "b200_h2d_materials",
    "b200_h2d_randoms", "b200_h2d_textures", "b200_h2d_lightInformation", "b200_d2h_bitmap", "b200_render",
    "b200_last_error", "b200_clear_error", "b200_set_device", "b200_set_option", "b200_set_stream", "b200_set_limits", "b200_set_partition",
    "b200_device_buffers", "b200_d2h_post", "b200_debug_relayout_boxes", "b200_frame_parameter_bytes", "b200_debug_build_unordered", "b200_debug_build_walk_trees", "b200_debug_counters", "b200_get_counters", "b200_last_render_ms", "b200_kernel_launches", "b200_scene_stats", "b200_scene_upload_stats",
    "b200_scene_layout", "b200_scene_adopt_layout", "b200_scene_device_arrays", "b200_scene_adopt_finish",
    "b200_rotate_primitives", "b200_translate_primitives", "b200_scale_primitives", "b200_d2h_scene", "b200_last_animation_ms",
    "b200_synchronize", "b200_measure_fp32_peak", "b200_register_host", "b200_unregister_host", "b200_frames_streamed", "b200_stream_target", "b200_accumulation_clear", "b200_accumulation_export", "b200_accumulation_import_and_pack", "b200_peer_frame_export", "b200_peer_frame_open", "b200_d2h_primitive_id",
]


def build_walk_trees(arrays):
    """Host-only: (nodes float32[n, 8, 4], prim_leaf int32[m], nb_main, nb_ext) as h2d_scene builds them for the
    order-independent walks."""
    lib = load()
    lib.b200_debug_build_walk_trees.restype = C.c_int
    lib.b200_debug_build_walk_trees.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                C.POINTER(C.c_int), C.POINTER(C.c_int)]
    boxes = np.ascontiguousarray(arrays["boxes"]); prims = np.ascontiguousarray(arrays["primitives"])
    m = int(arrays["nbPrimitives"])
    prim_leaf = np.zeros(max(m, 1), np.int32)
    nb_main, nb_ext = C.c_int(0), C.c_int(0)
    n4 = lib.b200_debug_build_walk_trees(_ptr(boxes), int(arrays["nbBoxes"]), _ptr(prims), m, None, 0, _ptr(prim_leaf),
                                         C.byref(nb_main), C.byref(nb_ext))
    nodes = np.zeros((max(n4, 1), 4), np.float32)
    lib.b200_debug_build_walk_trees(_ptr(boxes), int(arrays["nbBoxes"]), _ptr(prims), m, _ptr(nodes), n4, _ptr(prim_leaf),
                                    C.byref(nb_main), C.byref(nb_ext))
    return nodes[:n4].reshape(-1, 8, 4), prim_leaf[:m], nb_main.value, nb_ext.value


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


BOX_LAYOUT_AUTO, BOX_LAYOUT_LITERAL, BOX_LAYOUT_BVH = 0, 1, 2


def relayout_boxes(boxes_u8, nb_boxes, layout=BOX_LAYOUT_AUTO):
    """Host-only: the compact device box list (float32 [n', 8]) for a flattened reference box array."""
    lib = load()
    lib.b200_set_option(1, layout)
    b = np.ascontiguousarray(boxes_u8)
    out = np.zeros((max(2 * nb_boxes, 1), 8), np.float32)
    n = lib.b200_debug_relayout_boxes(_ptr(b), nb_boxes, _ptr(out), out.shape[0])
    lib.b200_set_option(1, BOX_LAYOUT_AUTO)
    return out[:n].copy()


def build_unordered(boxes_u8, nb_boxes):
    """Host-only: the unordered SAH BVH (binary depth-first list, float32 [n, 8]) over the reference's leaves."""
    lib = load()
    b = np.ascontiguousarray(boxes_u8)
    out = np.zeros((max(2 * nb_boxes, 1), 8), np.float32)
    n = lib.b200_debug_build_unordered(_ptr(b), nb_boxes, _ptr(out), out.shape[0])
    return out[:n].copy()


class Engine:
    """One GPU's engine, driven exactly as CudaKernel drives the reference's (initBuffers -> h2d_* ->
    cudaRender -> d2h_bitmap).  `arrays` is the dict of flattened wire-format arrays (uint8 views)."""

    OCC = wire.Int2(1, 1)

    def __init__(self, scene_info, device=None, limits=None, rank=0, world=1, stream=None):
        self.lib = load()
        if device is not None:
            self.lib.b200_set_device(device)
        if limits is not None:
            self.lib.b200_set_limits(limits[0], limits[1])
        self.limits = limits or (1920, 1080)
        self.lib.b200_clear_error()
        self.lib.b200_set_partition(rank, world)
        self.lib.b200_initialize_scene(self.OCC, scene_info, 0, 0, 0)
        if stream is not None:
            self.lib.b200_set_stream(C.c_void_p(stream))
        self.lib.b200_reshape_scene(self.OCC, scene_info)
        self.check()
        self.objects = wire.Int4(0, 0, 0, 0)
        self._keep = {}

    def check(self):
        buf = C.create_string_buffer(256)
        code = self.lib.b200_last_error(buf, 256)
        if code != 0:
            raise EngineError("solr_b200 engine error %d: %s" % (code, buf.value.decode()))

    def upload(self, arrays, randoms=None, textures=None):
        a = arrays
        self._keep = {k: np.ascontiguousarray(v) for k, v in a.items() if isinstance(v, np.ndarray)}
        k = self._keep
        self.lib.b200_h2d_scene(self.OCC, _ptr(k["boxes"]), a["nbBoxes"], _ptr(k["primitives"]), a["nbPrimitives"],
                                _ptr(k.get("lamps")) if a.get("nbLamps", 0) else None, a.get("nbLamps", 0))
        self.lib.b200_h2d_lightInformation(self.OCC, _ptr(k["lightInformation"]) if a["lightInformationSize"] else None,
                                           a["lightInformationSize"])
        self.lib.b200_h2d_materials(self.OCC, _ptr(k["materials"]), a["nbMaterials"])
        if randoms is not None:
            r = np.ascontiguousarray(randoms, np.float32)
            assert r.shape[0] >= self.limits[0] * self.limits[1]
            self.lib.b200_h2d_randoms(self.OCC, _ptr(r))
        if textures is not None:
            infos, n = textures
            self.lib.b200_h2d_textures(self.OCC, n, C.cast(infos, C.c_void_p))
        self.objects = wire.Int4(a["nbBoxes"], a["nbPrimitives"], a.get("nbLamps", 0), a["lightInformationSize"])
        self.check()

    def upload_small(self, arrays, randoms=None, textures=None):
        """Everything but the boxes and primitives (lights, materials, randoms, textures): what every process of a multi-GPU frame
        split still uploads itself when the scene proper arrives from the root over NVLink (partition.broadcast_scene)."""
        a = arrays
        self._keep = {k: np.ascontiguousarray(v) for k, v in a.items() if isinstance(v, np.ndarray) and k not in ("boxes", "primitives")}
        k = self._keep
        self.lib.b200_h2d_lightInformation(self.OCC, _ptr(k["lightInformation"]) if a["lightInformationSize"] else None,
                                           a["lightInformationSize"])
        self.lib.b200_h2d_materials(self.OCC, _ptr(k["materials"]), a["nbMaterials"])
        if randoms is not None:
            r = np.ascontiguousarray(randoms, np.float32)
            assert r.shape[0] >= self.limits[0] * self.limits[1]
            self.lib.b200_h2d_randoms(self.OCC, _ptr(r))
        if textures is not None:
            infos, n = textures
            self.lib.b200_h2d_textures(self.OCC, n, C.cast(infos, C.c_void_p))
        self.check()

    # ---- the animation step on the device-resident scene (csrc/animate.cuh) ----
    def rotate_primitives(self, center, angles):
        self.lib.b200_rotate_primitives(wire.Float3(*[float(v) for v in center]), wire.Float3(*[float(v) for v in angles]))
        self.check()

    def translate_primitives(self, t):
        self.lib.b200_translate_primitives(wire.Float3(*[float(v) for v in t]))
        self.check()

    def scale_primitives(self, scale):
        self.lib.b200_scale_primitives(float(scale))
        self.check()

    def download_scene(self):
        """The reference arrays as they are on the device (after animation steps): (boxes, primitives) as uint8 arrays."""
        boxes = np.zeros(self.objects.x * 48, np.uint8)
        prims = np.zeros(self.objects.y * 128, np.uint8)
        self.lib.b200_d2h_scene(_ptr(boxes), _ptr(prims))
        self.check()
        return boxes, prims

    def last_animation_ms(self):
        return float(self.lib.b200_last_animation_ms())

    def scene_layout(self):
        v = (C.c_longlong * 32)()
        n = self.lib.b200_scene_layout(v, 32)
        self.check()
        return [int(x) for x in v[:n]]

    def adopt_layout(self, layout, nb_lamps, light_information_size):
        """Device arrays for a scene another process built (layout = its scene_layout())."""
        v = (C.c_longlong * len(layout))(*layout)
        self.lib.b200_scene_adopt_layout(v, len(layout))
        self.objects = wire.Int4(int(layout[0]), int(layout[2]), nb_lamps, light_information_size)
        self.check()

    def scene_device_arrays(self):
        """[(device pointer or None, bytes)] of the scene's device arrays, in the order every process lists them."""
        ptrs = (C.c_void_p * 16)()
        nbytes = (C.c_longlong * 16)()
        n = self.lib.b200_scene_device_arrays(ptrs, nbytes, 16)
        self.check()
        return [(ptrs[i], int(nbytes[i])) for i in range(n)]

    def adopt_finish(self):
        self.lib.b200_scene_adopt_finish()
        self.check()

    def render(self, scene_info, eye, target, angles, post_info=None):
        """cudaRender: asynchronous on the engine's stream."""
        pp = post_info or wire.PostProcessingInfo()
        self.lib.b200_render(self.OCC, wire.Int4(8, 4, 1, 0), scene_info, self.objects, pp,
                             wire.Float3(*[float(v) for v in eye]), wire.Float3(*[float(v) for v in target]),
                             wire.Float4(*[float(v) for v in angles]))

    def readback(self, scene_info, bitmap=None, ids=None):
        """d2h_bitmap into caller-owned host buffers (allocated here if not given)."""
        W, H = scene_info.size.x, scene_info.size.y
        if bitmap is None:
            bitmap = np.zeros((H, W, 3), np.uint8)
        if ids is None:
            ids = np.zeros((H, W, 4), np.int32)
        self.lib.b200_d2h_bitmap(self.OCC, scene_info, _ptr(bitmap), _ptr(ids))
        self.check()
        return bitmap, ids

    def read_post_buffer(self, scene_info):
        """The float accumulation buffer (colorInfo, sceneInfo per pixel)."""
        post = np.zeros((scene_info.size.y, scene_info.size.x, 8), np.float32)
        self.lib.b200_d2h_post(scene_info, _ptr(post))
        self.check()
        return post

    def counters(self, reset=False):
        r, p = C.c_ulonglong(), C.c_ulonglong()
        self.lib.b200_get_counters(C.byref(r), C.byref(p), 1 if reset else 0)
        return int(r.value), int(p.value)

    def last_render_ms(self):
        return float(self.lib.b200_last_render_ms())

    def kernel_launches(self):
        return int(self.lib.b200_kernel_launches())

    def scene_stats(self):
        v = [C.c_int() for _ in range(4)]
        self.lib.b200_scene_stats(*[C.byref(x) for x in v])
        ms, n0, n1, gpu = C.c_float(), C.c_int(), C.c_int(), C.c_int()
        self.lib.b200_scene_upload_stats(C.byref(ms), C.byref(n0), C.byref(n1), C.byref(gpu))
        return {"boxes_in": v[0].value, "boxes_device": v[1].value, "primitives": v[2].value, "resident_ctas": v[3].value,
                "upload_ms": ms.value, "walk_tree_nodes": n0.value, "point_query_tree_nodes": n1.value, "trees_built_on_gpu": gpu.value}

    def device_buffers(self):
        b, i, p = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.lib.b200_device_buffers(C.byref(b), C.byref(i), C.byref(p))
        return b.value, i.value, p.value

    def set_option(self, key, value):
        self.lib.b200_set_option(key, value)

    def register_host(self, array):
        """b200_register_host: pins a host array of the caller in place (the caller keeps it alive and unregisters it before freeing it)."""
        return self.lib.b200_register_host(_ptr(array), array.nbytes)

    def unregister_host(self, array):
        return self.lib.b200_unregister_host(_ptr(array))

    def frames_streamed(self):
        """Frames whose bitmap / ids the ray kernels wrote into registered host buffers themselves (solr_b200.h, option key 12)."""
        return int(self.lib.b200_frames_streamed())

    def set_stream(self, stream):
        self.lib.b200_set_stream(C.c_void_p(stream) if stream else None)

    def synchronize(self):
        self.lib.b200_synchronize()

    def close(self):
        self.lib.b200_finalize_scene(self.OCC)
