"""ctypes mirrors of include/solr_b200_types.h (the wire format at the engine seam).

Layouts follow the reference's CUDA (float3) struct layout, /root/reference/solr/types.h:58-66,93-97,
140-168,183-189,210-251,254-260,264-286,301-308,323-329; sizes are asserted at import time
(SURVEY.md Appendix A).
"""
import ctypes as C


class Float2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Float4(C.Structure):
    _pack_ = 16  # noqa: (documentation only; ctypes honours _align_ on 3.13+, we pad by construction)
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class Int2(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int)]


class Int3(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("z", C.c_int)]


class Int4(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("z", C.c_int), ("w", C.c_int)]


class SceneInfo(C.Structure):
    _fields_ = [
        ("size", Int2), ("cameraType", C.c_int), ("graphicsLevel", C.c_int), ("nbRayIterations", C.c_int),
        ("transparentColor", C.c_float), ("viewDistance", C.c_float), ("shadowIntensity", C.c_float),
        ("eyeSeparation", C.c_float), ("renderBoxes", C.c_int), ("pathTracingIteration", C.c_int),
        ("maxPathTracingIterations", C.c_int), ("frameBufferType", C.c_int), ("timestamp", C.c_int),
        ("atmosphericEffect", C.c_int), ("doubleSidedTriangles", C.c_int), ("extendedGeometry", C.c_int),
        ("advancedIllumination", C.c_int), ("draftMode", C.c_int), ("skyboxRadius", C.c_int),
        ("skyboxMaterialId", C.c_int), ("gradientBackground", C.c_int), ("geometryEpsilon", C.c_float),
        ("rayEpsilon", C.c_float), ("backgroundColor", Float4),
    ]


class LightInformation(C.Structure):
    _fields_ = [("primitiveId", C.c_int), ("materialId", C.c_int), ("location", Float3), ("_pad", C.c_int * 3),
                ("color", Float4)]


class Material(C.Structure):
    _fields_ = [
        ("innerIllumination", Float4), ("color", Float4), ("specular", Float4),
        ("reflection", C.c_float), ("refraction", C.c_float), ("transparency", C.c_float), ("opacity", C.c_float),
        ("attributes", Int4), ("textureMapping", Int4), ("textureOffset", Int4), ("textureIds", Int4),
        ("advancedTextureOffset", Int4), ("advancedTextureIds", Int4), ("mappingOffset", Float2),
        ("_pad", C.c_int * 2),
    ]


class BoundingBox(C.Structure):
    _fields_ = [("parameters", Float3 * 2), ("nbPrimitives", C.c_int), ("startIndex", C.c_int),
                ("indexForNextBox", Int2), ("_pad", C.c_int * 2)]


class Primitive(C.Structure):
    _fields_ = [
        ("p0", Float3), ("p1", Float3), ("p2", Float3), ("n0", Float3), ("n1", Float3), ("n2", Float3),
        ("size", Float3), ("type", C.c_int), ("index", C.c_int), ("materialId", C.c_int),
        ("vt0", Float2), ("vt1", Float2), ("vt2", Float2), ("_pad", C.c_int * 2),
    ]


class TextureInfo(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("offset", C.c_int), ("size", Int3), ("type", C.c_int), ("_pad", C.c_int)]


class PostProcessingInfo(C.Structure):
    _fields_ = [("type", C.c_int), ("param1", C.c_float), ("param2", C.c_float), ("param3", C.c_int)]


class PostProcessingBuffer(C.Structure):
    _fields_ = [("colorInfo", Float4), ("sceneInfo", Float4)]


WIRE_SIZES = {
    SceneInfo: 112, LightInformation: 48, Material: 176, BoundingBox: 48, Primitive: 128, TextureInfo: 32,
    PostProcessingInfo: 16, PostProcessingBuffer: 32, Int4: 16, Float4: 16, Float3: 12,
}
for _t, _n in WIRE_SIZES.items():
    assert C.sizeof(_t) == _n, (_t.__name__, C.sizeof(_t), _n)
assert SceneInfo.rayEpsilon.offset == 92 and SceneInfo.backgroundColor.offset == 96
assert LightInformation.location.offset == 8 and LightInformation.color.offset == 32
assert Material.attributes.offset == 64 and Material.mappingOffset.offset == 160
assert BoundingBox.nbPrimitives.offset == 24 and BoundingBox.indexForNextBox.offset == 32
assert Primitive.size.offset == 72 and Primitive.type.offset == 84 and Primitive.vt0.offset == 96

# enums / constants (types.h:99-138,192-207; Consts.h:27-48)
CT_PERSPECTIVE, CT_ORTHOGRAPHIC, CT_ANAGLYPH, CT_VR, CT_PANORAMIC, CT_ANTIALIASED, CT_VOLUME = range(7)
GL_NO_SHADING, GL_PHONG, GL_PHONG_BLINN, GL_REFLECTIONS, GL_FULL = range(5)
AI_NONE, AI_BASIC, AI_FULL, AI_RANDOM = range(4)
PPE_NONE, PPE_DEPTH_OF_FIELD, PPE_AMBIENT_OCCLUSION, PPE_RADIOSITY, PPE_FILTER, PPE_CARTOON = range(6)
PT_SPHERE, PT_CYLINDER, PT_TRIANGLE, PT_CHECKBOARD, PT_CAMERA, PT_XYPLANE, PT_YZPLANE, PT_XZPLANE, \
    PT_MAGICCARPET, PT_ENVIRONMENT, PT_ELLIPSOID, PT_QUAD, PT_CONE = range(13)
MATERIAL_NONE = -1
TEXTURE_NONE = -1
NB_MAX_ITERATIONS = 10
REF_MAX_BITMAP_SIZE = 1920 * 1080


def default_scene_info(width, height, graphics_level=GL_FULL, nb_ray_iterations=3):
    """Scene defaults of the reference's Scene::initialize (apps/scenes/Scene.cpp:201-228), SURVEY §8(d)."""
    si = SceneInfo()
    si.size = Int2(width, height)
    si.cameraType = CT_PERSPECTIVE
    si.graphicsLevel = graphics_level
    si.nbRayIterations = nb_ray_iterations
    si.transparentColor = 0.0
    si.viewDistance = 50000.0
    si.shadowIntensity = 1.0
    si.eyeSeparation = 380.0
    si.renderBoxes = 0
    si.pathTracingIteration = 0
    si.maxPathTracingIterations = 1
    si.frameBufferType = 0
    si.timestamp = 0
    si.atmosphericEffect = 0
    si.doubleSidedTriangles = 0
    si.extendedGeometry = 1
    si.advancedIllumination = AI_NONE
    si.draftMode = 0
    si.skyboxRadius = 9000
    si.skyboxMaterialId = MATERIAL_NONE
    si.gradientBackground = 0
    si.geometryEpsilon = 0.001
    si.rayEpsilon = 0.05
    si.backgroundColor = Float4(0.0, 0.0, 0.0, 0.5)
    return si
