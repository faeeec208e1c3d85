/* solr_b200_types.h — wire format at the engine seam.
 *
 * These PODs are byte-compatible with the structs the reference passes across its engine boundary in
 * the CUDA (`float3`) layout: /root/reference/solr/types.h:93-97 (PostProcessingBuffer), :140-168
 * (SceneInfo), :183-189 (LightInformation), :210-251 (Material), :254-260 (BoundingBox), :264-286
 * (Primitive), :301-308 (TextureInfo), :323-329 (PostProcessingInfo) and the vecNf/vecNi typedefs at
 * :58-66.  Sizes/offsets are asserted below and re-checked against the compiled
 * reference in tests/test_wire_format.py.  Plain C so that C, C++, CUDA and ctypes all see one layout.
 */
#ifndef SOLR_B200_TYPES_H
#define SOLR_B200_TYPES_H

#include <stddef.h>
#include <stdint.h>

#if defined(_MSC_VER)
#define SOLR_B200_ALIGN16 __declspec(align(16))
#else
#define SOLR_B200_ALIGN16 __attribute__((aligned(16)))
#endif
#if defined(__cplusplus)
#define SOLR_B200_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define SOLR_B200_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

/* vector scalars: same size/alignment as CUDA's float2/float3/float4/int2/int3/int4 */
typedef struct { float x, y; } b200_float2;              /* 8 B, align 4 (CUDA: align 8; only used inside 16-aligned structs at 8-multiples) */
typedef struct { float x, y, z; } b200_float3;            /* 12 B, align 4 */
typedef struct SOLR_B200_ALIGN16 { float x, y, z, w; } b200_float4; /* 16 B, align 16 */
typedef struct { int x, y; } b200_int2;
typedef struct { int x, y, z; } b200_int3;
typedef struct SOLR_B200_ALIGN16 { int x, y, z, w; } b200_int4;

/* enums of types.h:99-138,192-207,289-321 (values only) */
enum { B200_CT_PERSPECTIVE = 0, B200_CT_ORTHOGRAPHIC = 1, B200_CT_ANAGLYPH = 2, B200_CT_VR = 3,
       B200_CT_PANORAMIC = 4, B200_CT_ANTIALIASED = 5, B200_CT_VOLUME = 6 };
enum { B200_FT_RGB = 0, B200_FT_BGR = 1 };
enum { B200_AI_NONE = 0, B200_AI_BASIC = 1, B200_AI_FULL = 2, B200_AI_RANDOM = 3 };
enum { B200_GL_NO_SHADING = 0, B200_GL_PHONG = 1, B200_GL_PHONG_BLINN = 2, B200_GL_REFLECTIONS = 3, B200_GL_FULL = 4 };
enum { B200_AE_NONE = 0, B200_AE_FOG = 1 };
enum { B200_PT_SPHERE = 0, B200_PT_CYLINDER = 1, B200_PT_TRIANGLE = 2, B200_PT_CHECKBOARD = 3, B200_PT_CAMERA = 4,
       B200_PT_XYPLANE = 5, B200_PT_YZPLANE = 6, B200_PT_XZPLANE = 7, B200_PT_MAGICCARPET = 8,
       B200_PT_ENVIRONMENT = 9, B200_PT_ELLIPSOID = 10, B200_PT_QUAD = 11, B200_PT_CONE = 12 };
enum { B200_PPE_NONE = 0, B200_PPE_DEPTH_OF_FIELD = 1, B200_PPE_AMBIENT_OCCLUSION = 2, B200_PPE_RADIOSITY = 3,
       B200_PPE_FILTER = 4, B200_PPE_CARTOON = 5 };

/* constants of Consts.h:27-48 that the path reads */
#define B200_NB_MAX_ITERATIONS 10
#define B200_MATERIAL_NONE (-1)
#define B200_TEXTURE_NONE (-1)
#define B200_TEXTURE_MANDELBROT (-2)
#define B200_TEXTURE_JULIA (-3)
#define B200_COLOR_DEPTH 3
#define B200_NB_MAX_MATERIALS (65506 + 30)
#define B200_NB_MAX_TEXTURES 512
#define B200_NB_MAX_LIGHTINFORMATIONS 512
#define B200_REF_MAX_BITMAP_WIDTH 1920  /* reference cap; this engine's cap is set by b200_set_limits */
#define B200_REF_MAX_BITMAP_HEIGHT 1080

typedef struct SOLR_B200_ALIGN16 {
    b200_int2 size;
    int cameraType;
    int graphicsLevel;
    int nbRayIterations;
    float transparentColor;
    float viewDistance;
    float shadowIntensity;
    float eyeSeparation;
    int renderBoxes;
    int pathTracingIteration;
    int maxPathTracingIterations;
    int frameBufferType;
    int timestamp;
    int atmosphericEffect;
    int doubleSidedTriangles;
    int extendedGeometry;
    int advancedIllumination;
    int draftMode;
    int skyboxRadius;
    int skyboxMaterialId;
    int gradientBackground;
    float geometryEpsilon;
    float rayEpsilon;
    b200_float4 backgroundColor;
} b200_SceneInfo;

typedef struct SOLR_B200_ALIGN16 {
    int primitiveId;
    int materialId;
    b200_float3 location;
    b200_float4 color;
} b200_LightInformation;

typedef struct SOLR_B200_ALIGN16 {
    b200_float4 innerIllumination; /* x emission, y light jitter, z light range, w noise */
    b200_float4 color;
    b200_float4 specular;          /* x value, y power, z -> blinn.w, w coef */
    float reflection;
    float refraction;
    float transparency;
    float opacity;
    b200_int4 attributes;          /* x fast transparency, y procedural, z wireframe, w wireframe width */
    b200_int4 textureMapping;      /* x width, y height, z (deprecated), w depth */
    b200_int4 textureOffset;       /* x diffuse, y normal, z bump, w specular */
    b200_int4 textureIds;          /* x diffuse, y normal, z bump, w specular */
    b200_int4 advancedTextureOffset; /* x reflection, y transparency, z ambient occlusion */
    b200_int4 advancedTextureIds;
    b200_float2 mappingOffset;
} b200_Material;

typedef struct SOLR_B200_ALIGN16 {
    b200_float3 parameters[2]; /* min corner, max corner */
    int nbPrimitives;          /* leaf: count; inner: 0 */
    int startIndex;            /* leaf: first primitive; inner: depth */
    b200_int2 indexForNextBox; /* x: slots to advance on a miss */
} b200_BoundingBox;

typedef struct SOLR_B200_ALIGN16 {
    b200_float3 p0, p1, p2;
    b200_float3 n0, n1, n2;
    b200_float3 size;
    int type;
    int index;      /* original (pre-compaction) primitive id */
    int materialId;
    b200_float2 vt0, vt1, vt2;
} b200_Primitive;

typedef struct SOLR_B200_ALIGN16 {
    unsigned char* buffer; /* host pointer */
    int offset;            /* offset in the device texture atlas */
    b200_int3 size;        /* width, height, depth(bytes/texel) */
    int type;
} b200_TextureInfo;

typedef struct SOLR_B200_ALIGN16 {
    int type;
    float param1;
    float param2;
    int param3;
} b200_PostProcessingInfo;

typedef struct {
    b200_float4 colorInfo; /* rgb accumulation, w = first-hit depth */
    b200_float4 sceneInfo;
} b200_PostProcessingBuffer;

typedef b200_int4 b200_PrimitiveXYIdBuffer; /* x first-hit original id | -1, y iterations, z light accum, w shadow flag */
typedef unsigned char b200_BitmapBuffer;

SOLR_B200_STATIC_ASSERT(sizeof(b200_SceneInfo) == 112, "SceneInfo wire size");
SOLR_B200_STATIC_ASSERT(offsetof(b200_SceneInfo, rayEpsilon) == 92, "SceneInfo.rayEpsilon");
SOLR_B200_STATIC_ASSERT(offsetof(b200_SceneInfo, backgroundColor) == 96, "SceneInfo.backgroundColor");
SOLR_B200_STATIC_ASSERT(sizeof(b200_LightInformation) == 48, "LightInformation wire size");
SOLR_B200_STATIC_ASSERT(offsetof(b200_LightInformation, location) == 8, "LightInformation.location");
SOLR_B200_STATIC_ASSERT(offsetof(b200_LightInformation, color) == 32, "LightInformation.color");
SOLR_B200_STATIC_ASSERT(sizeof(b200_Material) == 176, "Material wire size");
SOLR_B200_STATIC_ASSERT(offsetof(b200_Material, attributes) == 64, "Material.attributes");
SOLR_B200_STATIC_ASSERT(offsetof(b200_Material, mappingOffset) == 160, "Material.mappingOffset");
SOLR_B200_STATIC_ASSERT(sizeof(b200_BoundingBox) == 48, "BoundingBox wire size");
SOLR_B200_STATIC_ASSERT(offsetof(b200_BoundingBox, nbPrimitives) == 24, "BoundingBox.nbPrimitives");
SOLR_B200_STATIC_ASSERT(offsetof(b200_BoundingBox, indexForNextBox) == 32, "BoundingBox.indexForNextBox");
SOLR_B200_STATIC_ASSERT(sizeof(b200_Primitive) == 128, "Primitive wire size");
SOLR_B200_STATIC_ASSERT(offsetof(b200_Primitive, size) == 72, "Primitive.size");
SOLR_B200_STATIC_ASSERT(offsetof(b200_Primitive, type) == 84, "Primitive.type");
SOLR_B200_STATIC_ASSERT(offsetof(b200_Primitive, vt0) == 96, "Primitive.vt0");
SOLR_B200_STATIC_ASSERT(sizeof(b200_TextureInfo) == 32, "TextureInfo wire size");
SOLR_B200_STATIC_ASSERT(sizeof(b200_PostProcessingInfo) == 16, "PostProcessingInfo wire size");
SOLR_B200_STATIC_ASSERT(sizeof(b200_PostProcessingBuffer) == 32, "PostProcessingBuffer wire size");
SOLR_B200_STATIC_ASSERT(sizeof(b200_PrimitiveXYIdBuffer) == 16, "PrimitiveXYIdBuffer wire size");

#endif /* SOLR_B200_TYPES_H */
