// oracle/ref_build/ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
//
// A flat C API over the UNMODIFIED reference (host class solr::CudaKernel + its engine's C-ABI,
// solr/engines/cuda/CudaRayTracer.h:25-67) so tests can
//   1. build a scene with the reference's own setters and grid-hierarchy builder
//      (GPUKernel.cpp:495-684 setPrimitive, :1780-1909 setMaterial, :1041-1083 compactBoxes) and read
//      back the flattened wire-format arrays the engine would receive, and
//   2. render through the reference engine's seam functions with a caller-supplied random table and
//      timestamp.  GPUKernel::render_begin (GPUKernel.cpp:2712-2727) draws both from rand()/time(0)
//      inside an OpenMP loop, which is the only source of non-determinism on the path; calling the
//      seam directly (exactly what CudaKernel::render_begin does at CudaKernel.cpp:197-293) removes it.
// Compiled into oracle/_ref/libsolr_ref_{cpu,cuda}.so by the Makefile in this directory.
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include <cuda_runtime_api.h>

#include <Consts.h>
#include <types.h>
#include <engines/cuda/CudaKernel.h>
#include <engines/cuda/CudaRayTracer.h>
#ifdef REFH_B200
// Variant: the reference's host library driving the B200 engine through integration/B200Kernel (the class a
// maintainer adds to the reference tree).  Same C API; refh_render goes through render_begin/render_end.
#include <B200Kernel.h>
#include <solr_b200.h>
#define REFH_BASE solr::B200Kernel
#else
#define REFH_BASE solr::CudaKernel
#endif
#include <io/PDBReader.h>
#include <io/OBJReader.h>

#ifndef REFH_B200
// Non-static globals of the reference engine (CudaRayTracer.cu:38,40); read back for parity on the
// float accumulation buffer, which the seam itself never returns.
extern PostProcessingBuffer* d_postProcessingBuffer[MAX_GPU_COUNT];
#endif

namespace
{
class HarnessKernel : public REFH_BASE
{
public:
    BoundingBox* boxes() { return m_hBoundingBoxes; }
    Primitive* primitives() { return m_hPrimitives; }
    Material* materials() { return m_hMaterials; }
    int* lamps() { return m_hLamps; }
    LightInformation* lightInformation() { return m_lightInformation; }
    TextureInfo* textures() { return m_hTextures; }
    int nbBoxes() { return m_nbActiveBoxes[m_frame]; }
    int nbPrimitives() { return m_nbActivePrimitives[m_frame]; }
    int nbLamps() { return m_nbActiveLamps[m_frame]; }
    int nbMaterials() { return m_nbActiveMaterials + 1; }
    int lightInformationSize() { return m_lightInformationSize; }
    vec2i occupancy() { return m_occupancyParameters; }
    void sceneBounds(float* out6)
    {
        out6[0] = m_minPos[m_frame].x; out6[1] = m_minPos[m_frame].y; out6[2] = m_minPos[m_frame].z;
        out6[3] = m_maxPos[m_frame].x; out6[4] = m_maxPos[m_frame].y; out6[5] = m_maxPos[m_frame].z;
    }
};
} // namespace

struct RefhScene
{
    const void* boxes; int nbBoxes;
    const void* primitives; int nbPrimitives;
    const void* materials; int nbMaterials;
    const void* lightInformation; int lightInformationSize;
    const int* lamps; int nbLamps;
    float bounds[6];
};

extern "C" {

// The reference allocates its ~1.6 MB kernel object with plain `new`; glibc serves that from fresh
// mmap pages, so the never-initialised scene bounds m_minPos/m_maxPos are zero in
// practice.  calloc + placement new makes that explicit.
void* refh_create(const SceneInfo* sceneInfo)
{
    void* mem = calloc(1, sizeof(HarnessKernel));
    HarnessKernel* k = new (mem) HarnessKernel();
    k->setSceneInfo(*sceneInfo);
    k->initBuffers();
    k->setFrame(0);
    return k;
}

#ifdef REFH_B200
// the drop-in beyond the reference's frame limit: B200Kernel::setLimits before the buffers are made
void* refh_create_limits(const SceneInfo* sceneInfo, int maxWidth, int maxHeight)
{
    void* mem = calloc(1, sizeof(HarnessKernel));
    HarnessKernel* k = new (mem) HarnessKernel();
    k->setLimits(maxWidth, maxHeight);
    k->setSceneInfo(*sceneInfo);
    k->initBuffers();
    k->setFrame(0);
    return k;
}
#endif

void refh_destroy(void* h)
{
    HarnessKernel* k = static_cast<HarnessKernel*>(h);
    k->~HarnessKernel();
    free(h);
}

int refh_add_primitive(void* h, int type)
{
    return static_cast<HarnessKernel*>(h)->addPrimitive(static_cast<PrimitiveType>(type));
}

// v = x0,y0,z0, x1,y1,z1, x2,y2,z2, w,h,d
void refh_set_primitive(void* h, int index, const float* v, int materialId)
{
    static_cast<HarnessKernel*>(h)->setPrimitive(index, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9],
                                                 v[10], v[11], materialId);
}

void refh_set_primitive_normals(void* h, int index, const float* n)
{
    static_cast<HarnessKernel*>(h)->setPrimitiveNormals(index, make_vec3f(n[0], n[1], n[2]),
                                                        make_vec3f(n[3], n[4], n[5]), make_vec3f(n[6], n[7], n[8]));
}

void refh_set_primitive_texcoords(void* h, int index, const float* t)
{
    static_cast<HarnessKernel*>(h)->setPrimitiveTextureCoordinates(index, make_vec2f(t[0], t[1]),
                                                                   make_vec2f(t[2], t[3]), make_vec2f(t[4], t[5]));
}

int refh_add_material(void* h)
{
    return static_cast<HarnessKernel*>(h)->addMaterial();
}

// f = r,g,b,noise,reflection,refraction,transparency,opacity,specValue,specPower,specCoef,
//     innerIllumination,illuminationDiffusion,illuminationPropagation
// i = procedural,wireframe,wireframeWidth,diffuse,normal,bump,specular,reflection,transparent,ao,fastTransparency
void refh_set_material(void* h, int index, const float* f, const int* i)
{
    static_cast<HarnessKernel*>(h)->setMaterial(index, f[0], f[1], f[2], f[3], f[4], f[5], i[0] != 0, i[1] != 0, i[2],
                                                f[6], f[7], i[3], i[4], i[5], i[6], i[7], i[8], i[9], f[8], f[9], f[10],
                                                f[11], f[12], f[13], i[10] != 0);
}

// Bulk forms of the setters above (one FFI call per scene instead of per primitive).
void refh_add_materials(void* h, int n, const float* f, const int* i)
{
    HarnessKernel* k = static_cast<HarnessKernel*>(h);
    for (int m = 0; m < n; ++m)
    {
        const int id = k->addMaterial();
        const float* F = f + 14 * m;
        const int* I = i + 11 * m;
        k->setMaterial(id, F[0], F[1], F[2], F[3], F[4], F[5], I[0] != 0, I[1] != 0, I[2], F[6], F[7], I[3], I[4], I[5],
                       I[6], I[7], I[8], I[9], F[8], F[9], F[10], F[11], F[12], F[13], I[10] != 0);
    }
}

void refh_add_primitives(void* h, int n, const int* types, const float* v, const int* materialIds)
{
    HarnessKernel* k = static_cast<HarnessKernel*>(h);
    for (int p = 0; p < n; ++p)
    {
        const int id = k->addPrimitive(static_cast<PrimitiveType>(types[p]));
        const float* V = v + 12 * p;
        k->setPrimitive(id, V[0], V[1], V[2], V[3], V[4], V[5], V[6], V[7], V[8], V[9], V[10], V[11], materialIds[p]);
    }
}

void refh_set_normals_bulk(void* h, int first, int n, const float* normals)
{
    HarnessKernel* k = static_cast<HarnessKernel*>(h);
    for (int p = 0; p < n; ++p)
    {
        const float* N = normals + 9 * p;
        k->setPrimitiveNormals(first + p, make_vec3f(N[0], N[1], N[2]), make_vec3f(N[3], N[4], N[5]),
                               make_vec3f(N[6], N[7], N[8]));
    }
}

void refh_set_texture(void* h, int index, const unsigned char* texels, int width, int height, int depth)
{
    TextureInfo ti;
    memset(&ti, 0, sizeof(ti));
    ti.buffer = const_cast<unsigned char*>(texels);
    ti.size.x = width; ti.size.y = height; ti.size.z = depth;
    static_cast<HarnessKernel*>(h)->setTexture(index, ti);
}

void refh_set_material_raw(void* h, int index, const Material* m)
{
    static_cast<HarnessKernel*>(h)->setMaterial(index, *m);
}

int refh_compact_boxes(void* h)
{
    return static_cast<HarnessKernel*>(h)->compactBoxes(true);
}

// the animation step of the reference's scenes: move the primitives, refresh the box bounds, flatten again without a rebuild
// (GPUKernel.cpp:1378-1513, :1574-1600; MoleculeScene.cpp:75-81)
int refh_compact_boxes_mode(void* h, int reconstruct)
{
    return static_cast<HarnessKernel*>(h)->compactBoxes(reconstruct != 0);
}
void refh_rotate_primitives(void* h, const float* center3, const float* angles3)
{
    vec3f c = make_vec3f(center3[0], center3[1], center3[2]);
    vec4f a = make_vec4f(angles3[0], angles3[1], angles3[2], 0.f);
    static_cast<HarnessKernel*>(h)->rotatePrimitives(c, a);
}
void refh_translate_primitives(void* h, const float* t3)
{
    vec3f t = make_vec3f(t3[0], t3[1], t3[2]);
    static_cast<HarnessKernel*>(h)->translatePrimitives(t);
}
void refh_scale_primitives(void* h, float scale)
{
    static_cast<HarnessKernel*>(h)->scalePrimitives(scale, 0, 0);
}

#ifdef REFH_B200
// the same steps applied on the device through integration/B200Kernel (returns 1 when the device did it, 0 when the host fallback ran)
int refh_rotate_primitives_on_device(void* h, const float* center3, const float* angles3)
{
    vec3f c = make_vec3f(center3[0], center3[1], center3[2]);
    vec4f a = make_vec4f(angles3[0], angles3[1], angles3[2], 0.f);
    return static_cast<HarnessKernel*>(h)->rotatePrimitivesOnDevice(c, a) ? 1 : 0;
}
int refh_translate_primitives_on_device(void* h, const float* t3)
{
    vec3f t = make_vec3f(t3[0], t3[1], t3[2]);
    return static_cast<HarnessKernel*>(h)->translatePrimitivesOnDevice(t) ? 1 : 0;
}
void refh_sync_from_device(void* h) { static_cast<HarnessKernel*>(h)->syncFromDevice(); }
#endif

int refh_load_molecule(void* h, const char* filename, int geometryType, float atomSize, float stickSize,
                       int materialType, float scale)
{
    HarnessKernel* k = static_cast<HarnessKernel*>(h);
    solr::PDBReader reader;
    reader.loadAtomsFromFile(filename, *k, static_cast<solr::GeometryType>(geometryType), atomSize, stickSize,
                             materialType, make_vec4f(scale, scale, scale));
    return reader.getNbPrimitives();
}

void refh_get_scene(void* h, RefhScene* out)
{
    HarnessKernel* k = static_cast<HarnessKernel*>(h);
    out->boxes = k->boxes(); out->nbBoxes = k->nbBoxes();
    out->primitives = k->primitives(); out->nbPrimitives = k->nbPrimitives();
    out->materials = k->materials(); out->nbMaterials = k->nbMaterials();
    out->lightInformation = k->lightInformation(); out->lightInformationSize = k->lightInformationSize();
    out->lamps = k->lamps(); out->nbLamps = k->nbLamps();
    k->sceneBounds(out->bounds);
}

#ifdef REFH_B200
// One frame through the reference's own frame protocol (GPUKernel::setSceneInfo / setCamera / render_begin /
// render_end / getBitmap) with B200Kernel as the engine host class — the drop-in path.
void refh_render(void* h, const SceneInfo* sceneInfo, const PostProcessingInfo* post, const float* eye,
                 const float* target, const float* angles, const float* randoms, const int* block,
                 unsigned char* bitmap, int* ids, float* postBuffer)
{
    HarnessKernel* k = static_cast<HarnessKernel*>(h);
    SceneInfo si = *sceneInfo;
    si.maxPathTracingIterations = si.pathTracingIteration + 1; // keep m_refresh true for this frame
    k->setSceneInfo(si);
    k->setPostProcessingInfo(*post);
    k->setRandoms(randoms, static_cast<size_t>(MAX_BITMAP_WIDTH) * MAX_BITMAP_HEIGHT, sceneInfo->timestamp);
    k->setCamera(make_vec3f(eye[0], eye[1], eye[2]), make_vec3f(target[0], target[1], target[2]),
                 make_vec4f(angles[0], angles[1], angles[2], angles[3]));
    k->render_begin(0.f);
    k->render_end();
    const size_t px = static_cast<size_t>(sceneInfo->size.x) * sceneInfo->size.y;
    // within the reference's frame limit these are GPUKernel::getBitmap() / getPrimitiveAt(); beyond it, B200Kernel's own buffers
    memcpy(bitmap, k->getFrame(), px * gColorDepth);
    for (int y = 0; y < sceneInfo->size.y; ++y)
        for (int x = 0; x < sceneInfo->size.x; ++x)
            ids[4 * (y * sceneInfo->size.x + x)] = static_cast<int>(k->getPrimitiveIdAt(x, y));
    if (postBuffer) b200_d2h_post(*reinterpret_cast<const b200_SceneInfo*>(sceneInfo), reinterpret_cast<b200_PostProcessingBuffer*>(postBuffer));
}
#else
// One frame through the reference engine's seam.  randoms must hold MAX_BITMAP_WIDTH*MAX_BITMAP_HEIGHT
// floats (h2d_randoms copies exactly that many, CudaRayTracer.cu:1572-1574).  block = {bx, by}; the
// reference's 12x12 (CudaKernel.cpp:85-87) makes threads with x >= width alias the next row's first
// pixels (no x bound check at CudaRayTracer.cu:445-458), so callers pass a block width dividing W.
// bitmap: W*H*3 bytes, ids: W*H int4, post: W*H*8 floats or null.
void refh_render(void* h, const SceneInfo* sceneInfo, const PostProcessingInfo* post, const float* eye,
                 const float* target, const float* angles, const float* randoms, const int* block,
                 unsigned char* bitmap, int* ids, float* postBuffer)
{
    HarnessKernel* k = static_cast<HarnessKernel*>(h);
    const vec2i occ = k->occupancy();
    h2d_scene(occ, k->boxes(), k->nbBoxes(), k->primitives(), k->nbPrimitives(), k->lamps(), k->nbLamps());
    h2d_lightInformation(occ, k->lightInformation(), k->lightInformationSize());
    h2d_randoms(occ, const_cast<float*>(randoms));
    h2d_materials(occ, k->materials(), k->nbMaterials());
    h2d_textures(occ, NB_MAX_TEXTURES, k->textures());

    vec4i objects = make_vec4i(k->nbBoxes(), k->nbPrimitives(), k->nbLamps(), k->lightInformationSize());
    vec4i blockSize = make_vec4i(block[0], block[1], 1, 0);
    cudaRender(occ, blockSize, *sceneInfo, objects, *post, make_vec3f(eye[0], eye[1], eye[2]),
               make_vec3f(target[0], target[1], target[2]), make_vec4f(angles[0], angles[1], angles[2], angles[3]));
    d2h_bitmap(occ, *sceneInfo, bitmap, reinterpret_cast<PrimitiveXYIdBuffer*>(ids));
    cudaDeviceSynchronize();
    if (postBuffer)
        cudaMemcpy(postBuffer, d_postProcessingBuffer[0],
                   sizeof(PostProcessingBuffer) * sceneInfo->size.x * sceneInfo->size.y, cudaMemcpyDeviceToHost);
}

#endif

// Struct sizes as the reference compiles them — lets tests assert the wire format of include/*.h.
void refh_struct_sizes(int* out)
{
    out[0] = sizeof(SceneInfo); out[1] = sizeof(BoundingBox); out[2] = sizeof(Primitive); out[3] = sizeof(Material);
    out[4] = sizeof(LightInformation); out[5] = sizeof(PostProcessingInfo); out[6] = sizeof(PostProcessingBuffer);
    out[7] = sizeof(PrimitiveXYIdBuffer); out[8] = sizeof(TextureInfo); out[9] = sizeof(Ray);
}

int refh_limits(int which)
{
    switch (which)
    {
    case 0: return MAX_BITMAP_WIDTH;
    case 1: return MAX_BITMAP_HEIGHT;
    case 2: return NB_MAX_BOXES;
    case 3: return NB_MAX_PRIMITIVES;
    case 4: return NB_MAX_MATERIALS;
    default: return -1;
    }
}
}
