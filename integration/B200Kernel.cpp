/* B200Kernel.cpp — see B200Kernel.h.  Mirrors solr/engines/cuda/CudaKernel.cpp:81-145 (ctor, initBuffers,
 * cleanup, initializeDevice, releaseDevice), :174-302 (render_begin) and :304-313 (render_end, minus the GL
 * blit, which belongs to the viewer). */
#include "B200Kernel.h"

#include <Consts.h>
#include <Logging.h>
#include <solr_b200.h>

#include <cstdlib>
#include <cstring>
#include <ctime>

namespace
{
/* The wire structs are layout-identical (include/solr_b200_types.h asserts sizes/offsets against SURVEY Appendix A;
 * these asserts tie them to the reference's own definitions at compile time). */
static_assert(sizeof(SceneInfo) == sizeof(b200_SceneInfo), "SceneInfo");
static_assert(sizeof(BoundingBox) == sizeof(b200_BoundingBox), "BoundingBox");
static_assert(sizeof(Primitive) == sizeof(b200_Primitive), "Primitive");
static_assert(sizeof(Material) == sizeof(b200_Material), "Material");
static_assert(sizeof(LightInformation) == sizeof(b200_LightInformation), "LightInformation");
static_assert(sizeof(TextureInfo) == sizeof(b200_TextureInfo), "TextureInfo");
static_assert(sizeof(PostProcessingInfo) == sizeof(b200_PostProcessingInfo), "PostProcessingInfo");
static_assert(sizeof(PrimitiveXYIdBuffer) == sizeof(b200_PrimitiveXYIdBuffer), "PrimitiveXYIdBuffer");

template <typename To, typename From>
To as(const From &f)
{
    static_assert(sizeof(To) == sizeof(From), "wire struct size");
    To t;
    memcpy(&t, &f, sizeof(To));
    return t;
}
const b200_int2 OCC = {1, 1};
} // namespace

namespace solr
{
B200Kernel::B200Kernel()
    : GPUKernel(), m_deviceInitialized(false), m_fixedRandoms(false), m_hostStale(false), m_fixedTimestamp(0), m_maxWidth(MAX_BITMAP_WIDTH),
      m_maxHeight(MAX_BITMAP_HEIGHT)
{
    m_pinned[0] = m_pinned[1] = nullptr;
    m_occupancyParameters.x = 1;
    m_occupancyParameters.y = 1;
    m_gpuDescription = "B200 engine (libsolr_b200)";
}

B200Kernel::~B200Kernel() { releaseDevice(); }

void B200Kernel::setDeviceId(const int device) { b200_set_device(device); }
void B200Kernel::queryDevice() { LOG_INFO(1, "Device: " << m_gpuDescription); }
void B200Kernel::setLimits(int w, int h) { m_maxWidth = w; m_maxHeight = h; }
void B200Kernel::setPartition(int rank, int world) { b200_set_partition(rank, world); }

/* The per-frame animation step of MoleculeScene.cpp:75-81 (rotatePrimitives + compactBoxes(false) + re-upload), applied to the
 * arrays on the device: same arithmetic as GPUKernel.cpp:1378-1513, same arrays, about 2 ms instead of 0.45 s for 216 k primitives. */
bool B200Kernel::rotatePrimitivesOnDevice(const vec3f &rotationCenter, const vec4f &angles)
{
    if (m_deviceInitialized && m_primitivesTransfered)
    {
        const b200_float3 c = {rotationCenter.x, rotationCenter.y, rotationCenter.z}, a = {angles.x, angles.y, angles.z};
        if (b200_rotate_primitives(c, a) == 0) { m_hostStale = true; return true; }
        b200_clear_error();
    }
    syncFromDevice();
    rotatePrimitives(rotationCenter, angles);
    compactBoxes(false);
    return false;
}

bool B200Kernel::translatePrimitivesOnDevice(const vec3f &translation)
{
    if (m_deviceInitialized && m_primitivesTransfered)
    {
        const b200_float3 t = {translation.x, translation.y, translation.z};
        if (b200_translate_primitives(t) == 0) { m_hostStale = true; return true; }
        b200_clear_error();
    }
    syncFromDevice();
    translatePrimitives(translation);
    compactBoxes(false);
    return false;
}

/* the reference arrays come back, the primitives go back into the container by their ids, the boxes are re-fitted the way a
 * host-side step would have left them (GPUKernel.cpp:1394-1460 without the move) */
void B200Kernel::syncFromDevice()
{
    if (!m_hostStale) return;
    m_hostStale = false;
    if (b200_d2h_scene(reinterpret_cast<b200_BoundingBox *>(m_hBoundingBoxes), reinterpret_cast<b200_Primitive *>(m_hPrimitives)) != 0) return;
    PrimitiveContainer &primitives = m_primitives[m_frame];
    for (int i = 0; i < m_nbActivePrimitives[m_frame]; ++i)
    {
        const Primitive &p = m_hPrimitives[i];
        PrimitiveContainer::iterator it = primitives.find(static_cast<unsigned int>(p.index));
        if (it == primitives.end()) continue;
        CPUPrimitive &cp = it->second;
        cp.p0 = p.p0; cp.p1 = p.p1; cp.p2 = p.p2; cp.n0 = p.n0; cp.n1 = p.n1; cp.n2 = p.n2; cp.size = p.size;
    }
    for (BoxContainer::iterator itb = m_boundingBoxes[m_frame][0].begin(); itb != m_boundingBoxes[m_frame][0].end(); ++itb)
    {
        CPUBoundingBox &box = itb->second;
        resetBox(box, false);
        if (!box.primitives.empty()) updateBoundingBox(box);
    }
    for (int b = 1; b < BOUNDING_BOXES_TREE_DEPTH; ++b)
        for (BoxContainer::iterator itb = m_boundingBoxes[m_frame][b].begin(); itb != m_boundingBoxes[m_frame][b].end(); ++itb)
            updateOutterBoundingBox(itb->second, b - 1);
}

void B200Kernel::setRandoms(const float *randoms, size_t count, int timestamp)
{
    const size_t n = static_cast<size_t>(MAX_BITMAP_WIDTH) * MAX_BITMAP_HEIGHT;
    memset(m_hRandoms, 0, n * sizeof(float));
    memcpy(m_hRandoms, randoms, (count < n ? count : n) * sizeof(float));
    m_fixedRandoms = true;
    m_fixedTimestamp = timestamp;
    m_randomsTransfered = false;
}

void B200Kernel::initBuffers()
{
    GPUKernel::initBuffers();
    queryDevice();
    initializeDevice();
}

void B200Kernel::cleanup()
{
    unpin(); /* before GPUKernel frees the buffers */
    GPUKernel::cleanup();
    releaseDevice();
}

void B200Kernel::unpin()
{
    for (int i = 0; i < 2; ++i)
        if (m_pinned[i])
        {
            b200_unregister_host(m_pinned[i]);
            m_pinned[i] = nullptr;
        }
}

unsigned int B200Kernel::getPrimitiveIdAt(int x, int y)
{
    if (m_bigIds.empty())
        return getPrimitiveAt(x, y);
    if (x < 0 || y < 0 || x >= m_sceneInfo.size.x || y >= m_sceneInfo.size.y)
        return 0;
    return m_bigIds[static_cast<size_t>(y) * m_sceneInfo.size.x + x].x;
}

void B200Kernel::initializeDevice()
{
    b200_set_limits(m_maxWidth, m_maxHeight);
    b200_initialize_scene(OCC, as<b200_SceneInfo>(m_sceneInfo), NB_MAX_PRIMITIVES, NB_MAX_LAMPS, NB_MAX_MATERIALS);
    b200_reshape_scene(OCC, as<b200_SceneInfo>(m_sceneInfo));
    unpin();
    if (m_maxWidth * m_maxHeight > static_cast<int>(MAX_BITMAP_SIZE))
    {
        m_bigBitmap.assign(static_cast<size_t>(m_maxWidth) * m_maxHeight * gColorDepth, 0);
        m_bigIds.assign(static_cast<size_t>(m_maxWidth) * m_maxHeight, PrimitiveXYIdBuffer());
        m_bigRandoms.assign(static_cast<size_t>(m_maxWidth) * m_maxHeight, 0.f);
    }
    /* the frame and id buffers live as long as this object's buffers do: pinned in place for direct read-backs */
    void *frame = m_bigBitmap.empty() ? static_cast<void *>(m_bitmap) : static_cast<void *>(m_bigBitmap.data());
    void *ids = m_bigIds.empty() ? static_cast<void *>(m_hPrimitivesXYIds) : static_cast<void *>(m_bigIds.data());
    const size_t px = static_cast<size_t>(m_maxWidth) * m_maxHeight;
    if (frame && b200_register_host(frame, px * gColorDepth) == 0)
        m_pinned[0] = frame;
    if (ids && b200_register_host(ids, px * sizeof(PrimitiveXYIdBuffer)) == 0)
        m_pinned[1] = ids;
    m_deviceInitialized = true;
}

void B200Kernel::releaseDevice()
{
    unpin();
    if (m_deviceInitialized)
        b200_finalize_scene(OCC);
    m_deviceInitialized = false;
}

void B200Kernel::render_begin(const float timer)
{
    if (m_fixedRandoms)
        m_sceneInfo.timestamp = m_fixedTimestamp; /* skip GPUKernel::render_begin's rand()/time(0) draw */
    else if (m_bigRandoms.empty())
        GPUKernel::render_begin(timer);
    else
    {
        /* GPUKernel::render_begin (GPUKernel.cpp:2712-2727) fills size.x * size.y entries of a table that holds MAX_BITMAP_SIZE:
         * beyond the reference's frame limit the same draw goes into the table of this class */
        m_sceneInfo.timestamp = rand() % 10000;
        if (!m_randomsTransfered || m_sceneInfo.pathTracingIteration % 50 == 1)
        {
            m_randomsTransfered = false;
            srand(static_cast<int>(time(0)));
            for (size_t i = 0; i < m_bigRandoms.size(); ++i)
                m_bigRandoms[i] = 0.000005f * (rand() % 2000 - 1000);
        }
    }
    if (m_refresh)
    {
        const int nbBoxes = m_nbActiveBoxes[m_frame];
        const int nbPrimitives = m_nbActivePrimitives[m_frame];
        const int nbLamps = m_nbActiveLamps[m_frame];
        const int nbMaterials = m_nbActiveMaterials + 1;
        if (!m_primitivesTransfered)
        {
            b200_h2d_scene(OCC, reinterpret_cast<const b200_BoundingBox *>(m_hBoundingBoxes), nbBoxes,
                           reinterpret_cast<const b200_Primitive *>(m_hPrimitives), nbPrimitives, m_hLamps, nbLamps);
            b200_h2d_lightInformation(OCC, reinterpret_cast<const b200_LightInformation *>(m_lightInformation),
                                      m_lightInformationSize);
            m_primitivesTransfered = true;
        }
        if (!m_randomsTransfered)
        {
            if (m_maxWidth * m_maxHeight > static_cast<int>(MAX_BITMAP_SIZE))
            {
                /* the engine copies maxWidth*maxHeight floats; the reference's table only has 1920x1080 */
                if (m_fixedRandoms)
                    memcpy(m_bigRandoms.data(), m_hRandoms, sizeof(float) * MAX_BITMAP_SIZE);
                b200_h2d_randoms(OCC, m_bigRandoms.data());
            }
            else
                b200_h2d_randoms(OCC, m_hRandoms);
            m_randomsTransfered = true;
        }
        if (!m_materialsTransfered)
        {
            realignTexturesAndMaterials();
            b200_h2d_materials(OCC, reinterpret_cast<const b200_Material *>(m_hMaterials), nbMaterials);
            m_materialsTransfered = true;
        }
        if (!m_texturesTransfered)
        {
            b200_h2d_textures(OCC, NB_MAX_TEXTURES, reinterpret_cast<const b200_TextureInfo *>(m_hTextures));
            m_texturesTransfered = true;
        }
        b200_int4 objects = {nbBoxes, nbPrimitives, nbLamps, m_lightInformationSize};
        SceneInfo sceneInfo = m_sceneInfo;
        if (m_sceneInfo.draftMode && m_sceneInfo.pathTracingIteration == 0)
            sceneInfo.graphicsLevel = glNoShading;
        if (m_sceneInfo.draftMode && m_sceneInfo.pathTracingIteration == m_sceneInfo.maxPathTracingIterations)
            sceneInfo.cameraType = ctAntialiazed;
        const b200_int4 blockSize = {8, 4, 1, 0};
        b200_render(OCC, blockSize, as<b200_SceneInfo>(sceneInfo), objects, as<b200_PostProcessingInfo>(m_postProcessingInfo),
                    as<b200_float3>(m_viewPos), as<b200_float3>(m_viewDir), as<b200_float4>(m_angles));
    }
    m_refresh = (m_sceneInfo.pathTracingIteration < m_sceneInfo.maxPathTracingIterations);
}

void B200Kernel::render_end()
{
    unsigned char *bitmap = m_bigBitmap.empty() ? m_bitmap : m_bigBitmap.data();
    PrimitiveXYIdBuffer *ids = m_bigIds.empty() ? m_hPrimitivesXYIds : m_bigIds.data();
    b200_d2h_bitmap(OCC, as<b200_SceneInfo>(m_sceneInfo), bitmap, reinterpret_cast<b200_PrimitiveXYIdBuffer *>(ids));
    /* The viewer's GL blit (CudaKernel.cpp:313-388) is unchanged and stays in the application layer. */
}
}
