// scene_host.cpp — see scene_host.h.  Literal restatement of the reference's host-side scene container,
// grid-hierarchy builder and frame protocol (paths relative to /root/reference/solr/engines/).
#include "scene_host.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <thread>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace solr_b200
{
namespace
{
const unsigned int AABB_MAGIC_NUMBER = 6400; // GPUKernel.cpp:69
const size_t REF_NB_MAX_BOXES = 2500000;     // Consts.h:32
const size_t REF_NB_MAX_PRIMITIVES = 2500000; // Consts.h:33
const int NB_MAX_LAMPS = 512;                // Consts.h:34

b200_float3 v3(float x, float y, float z) { b200_float3 r = {x, y, z}; return r; }
b200_float3 min2(b200_float3 a, b200_float3 b) { return v3(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)); }
b200_float3 max2(b200_float3 a, b200_float3 b) { return v3(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)); }
b200_float3 min3(b200_float3 a, b200_float3 b, b200_float3 c)
{
    return v3(std::min(std::min(a.x, b.x), c.x), std::min(std::min(a.y, b.y), c.y), std::min(std::min(a.z, b.z), c.z));
}
b200_float3 max3(b200_float3 a, b200_float3 b, b200_float3 c)
{
    return v3(std::max(std::max(a.x, b.x), c.x), std::max(std::max(a.y, b.y), c.y), std::max(std::max(a.z, b.z), c.z));
}
void normalizeVector(b200_float3& v) // GPUKernel.cpp:142-151
{
    float l = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    if (l != 0.f) { v.x /= l; v.y /= l; v.z /= l; }
}
// Static partition of [0, n) over host threads: chunk t gets the same range whatever the timing, so results do not depend on it.
unsigned int hostThreads()
{
    static const unsigned int n = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    return n;
}
template <class F> void parallelFor(size_t n, F&& body) // body(begin, end)
{
    const unsigned int T = (n < 65536) ? 1u : hostThreads();
    if (T == 1) { body((size_t)0, n); return; }
    std::vector<std::thread> pool;
    for (unsigned int t = 0; t < T; ++t)
    {
        const size_t b = n * t / T, e = n * (t + 1) / T;
        pool.emplace_back([&body, b, e]() { body(b, e); });
    }
    for (auto& th : pool) th.join();
}
// stable sort by key: chunks sorted on their own threads, then merged pairwise (std::inplace_merge keeps equal keys in order)
template <class E> void parallelStableSortByKey(std::vector<E>& v)
{
    auto less = [](const E& a, const E& b) { return a.key < b.key; };
    const size_t n = v.size();
    unsigned int T = (n < 65536) ? 1u : hostThreads();
    if (T == 1) { std::stable_sort(v.begin(), v.end(), less); return; }
    std::vector<size_t> cut(T + 1);
    for (unsigned int t = 0; t <= T; ++t) cut[t] = n * t / T;
    {
        std::vector<std::thread> pool;
        for (unsigned int t = 0; t < T; ++t) pool.emplace_back([&, t]() { std::stable_sort(v.begin() + cut[t], v.begin() + cut[t + 1], less); });
        for (auto& th : pool) th.join();
    }
    for (unsigned int width = 1; width < T; width *= 2)
    {
        std::vector<std::thread> pool;
        for (unsigned int t = 0; t + width < T; t += 2 * width)
        {
            const size_t b = cut[t], m = cut[t + width], e = cut[std::min(T, t + 2 * width)];
            pool.emplace_back([&, b, m, e]() { std::inplace_merge(v.begin() + b, v.begin() + m, v.begin() + e, less); });
        }
        for (auto& th : pool) th.join();
    }
}

b200_float3 crossProduct(const b200_float3& b, const b200_float3& c) // GPUKernel.cpp:153-160
{
    return v3(b.y * c.z - b.z * c.y, b.z * c.x - b.x * c.z, b.x * c.y - b.y * c.x);
}
} // namespace

SceneHost::SceneHost(const b200_SceneInfo& sceneInfo)
    : m_sceneInfo(sceneInfo), m_treeDepth(2), m_nbActiveTextures(0), m_nbActiveBoxes(0), m_nbActivePrimitives(0),
      m_nbActiveLamps(0), m_nbActiveMaterials(-1), m_lightInformationSize(0), m_maxBoxes(REF_NB_MAX_BOXES),
      m_maxPrimitives(REF_NB_MAX_PRIMITIVES), m_primitivesTransfered(false), m_materialsTransfered(false),
      m_texturesTransfered(false), m_randomsTransfered(false), m_refresh(true), m_deviceInitialised(false),
      m_maxWidth(B200_REF_MAX_BITMAP_WIDTH), m_maxHeight(B200_REF_MAX_BITMAP_HEIGHT), m_rank(0), m_world(1), m_device(-1)
{
    memset(&m_postProcessingInfo, 0, sizeof(m_postProcessingInfo));
    m_viewPos = v3(0.f, 0.f, 0.f);
    m_viewDir = v3(0.f, 0.f, 0.f);
    m_angles.x = m_angles.y = m_angles.z = m_angles.w = 0.f;
    // The reference never initialises m_minPos/m_maxPos before the first scene (only cleanup() does,
    // GPUKernel.cpp:394-399); its ~1.6 MB kernel object comes from fresh zero pages, so they start at 0.
    m_minPos = v3(0.f, 0.f, 0.f);
    m_maxPos = v3(0.f, 0.f, 0.f);
    b200_Material zero;
    memset(&zero, 0, sizeof(zero));
    m_hMaterials.assign(B200_NB_MAX_MATERIALS + 1, zero); // GPUKernel.cpp:307-308
    m_textures.resize(B200_NB_MAX_TEXTURES);
    for (auto& t : m_textures) { t.offset = 0; t.size.x = t.size.y = t.size.z = 0; }
}

void SceneHost::unpinBuffers()
{
    if (m_pinnedBitmap) { b200_unregister_host(m_pinnedBitmap); m_pinnedBitmap = nullptr; }
    if (m_pinnedIds) { b200_unregister_host(m_pinnedIds); m_pinnedIds = nullptr; }
}

void SceneHost::dropSharedFrame()
{
    if (!m_shared) return;
    b200_SceneInfo none = m_sceneInfo;
    b200_stream_target(none, nullptr, nullptr); // waits for a frame in flight, then forgets the buffers
    unpinBuffers();
    munmap(m_shared, m_sharedBytes);
    if (m_sharedOwner) shm_unlink(m_sharedName.c_str());
    m_shared = nullptr; m_sharedBytes = 0; m_sharedOwner = false;
    m_bitmapPtr = m_bitmap.data(); m_idsPtr = m_primitivesXYIds.data();
}

int SceneHost::shareFrame(const char* name, bool create)
{
    if (!m_deviceInitialised || !name) return -4;
    dropSharedFrame();
    const size_t px = (size_t)m_maxWidth * m_maxHeight;
    const size_t bitmapBytes = (px * B200_COLOR_DEPTH + 4095) & ~(size_t)4095; // the id buffer starts on a page
    const size_t bytes = bitmapBytes + px * sizeof(b200_PrimitiveXYIdBuffer);
    const int fd = shm_open(name, create ? (O_CREAT | O_RDWR | O_TRUNC) : O_RDWR, 0600);
    if (fd < 0) return -13;
    if (create && ftruncate(fd, (off_t)bytes) != 0) { close(fd); shm_unlink(name); return -13; }
    void* base = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);
    close(fd);
    if (base == MAP_FAILED) { if (create) shm_unlink(name); return -13; }
    unpinBuffers();
    m_shared = base; m_sharedBytes = bytes; m_sharedName = name; m_sharedOwner = create;
    m_bitmapPtr = static_cast<unsigned char*>(base);
    m_idsPtr = reinterpret_cast<b200_PrimitiveXYIdBuffer*>(static_cast<unsigned char*>(base) + bitmapBytes);
    if (b200_register_host(m_bitmapPtr, px * B200_COLOR_DEPTH) == 0) m_pinnedBitmap = m_bitmapPtr;
    if (b200_register_host(m_idsPtr, px * sizeof(b200_PrimitiveXYIdBuffer)) == 0) m_pinnedIds = m_idsPtr;
    if (!m_pinnedBitmap || !m_pinnedIds) { dropSharedFrame(); return -12; }
    b200_SceneInfo si = m_sceneInfo;
    const int rc = b200_stream_target(si, m_bitmapPtr, m_idsPtr);
    if (rc != 0) { dropSharedFrame(); return rc; }
    return 0;
}

SceneHost::~SceneHost()
{
    dropFlat();
    dropSharedFrame();
    unpinBuffers();
    if (m_deviceInitialised)
    {
        b200_int2 occ = {1, 1};
        b200_finalize_scene(occ);
    }
}

void SceneHost::setCamera(const b200_float3& eye, const b200_float3& dir, const b200_float4& angles) // GPUKernel.cpp:481-493
{
    m_viewPos = eye; m_viewDir = dir; m_angles = angles;
    m_refresh = true;
}

int SceneHost::addPrimitive(int type) // GPUKernel.cpp:495-516
{
    HostPrimitive p;
    memset(&p, 0, sizeof(p));
    p.type = type;
    const int index = static_cast<int>(m_primitives.size());
    m_primitives[index] = p;
    return index;
}

// GPUKernel.cpp:540-684
void SceneHost::setPrimitive(int index, float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2,
                             float w, float h, float d, int materialId)
{
    const float scale = 1.f;
    m_primitivesTransfered = false;
    if (!(index >= 0 && (size_t)index <= m_primitives.size()))
    {
        fprintf(stderr, "[solr_b200] SceneHost::setPrimitive: out of bounds (%d)\n", index);
        return;
    }
    HostPrimitive& p = m_primitives[index];
    p.p0 = v3(x0 * scale, y0 * scale, z0 * scale);
    p.p1 = v3(x1 * scale, y1 * scale, z1 * scale);
    p.p2 = v3(x2 * scale, y2 * scale, z2 * scale);
    p.size = v3(w * scale, h * scale, d * scale);
    p.n0 = p.n1 = p.n2 = v3(0.f, 0.f, 0.f);
    p.vt0.x = p.vt0.y = p.vt1.x = p.vt1.y = p.vt2.x = p.vt2.y = 0.f;
    p.materialId = materialId;
    switch (p.type)
    {
    case B200_PT_SPHERE: p.size = v3(w * scale, w * scale, w * scale); break;
    case B200_PT_ELLIPSOID: p.size = v3(w * scale, h * scale, d * scale); break;
    case B200_PT_CYLINDER:
    case B200_PT_CONE:
    {
        b200_float3 axis = v3(x1 * scale - x0 * scale, y1 * scale - y0 * scale, z1 * scale - z0 * scale);
        float len = sqrt(axis.x * axis.x + axis.y * axis.y + axis.z * axis.z);
        if (len != 0.f) { axis.x /= len; axis.y /= len; axis.z /= len; }
        p.n1 = axis;
        p.p2 = v3((x0 * scale + x1 * scale) / 2.f, (y0 * scale + y1 * scale) / 2.f, (z0 * scale + z1 * scale) / 2.f);
        p.size = v3(w * scale, w * scale, w * scale);
        break;
    }
    case B200_PT_XYPLANE: p.n0 = v3(0.f, 0.f, 1.f); p.n1 = p.n0; p.n2 = p.n0; break;
    case B200_PT_YZPLANE: p.n0 = v3(1.f, 0.f, 0.f); p.n1 = p.n0; p.n2 = p.n0; break;
    case B200_PT_XZPLANE:
    case B200_PT_CHECKBOARD: p.n0 = v3(0.f, 1.f, 0.f); p.n1 = p.n0; p.n2 = p.n0; break;
    case B200_PT_TRIANGLE:
    {
        b200_float3 v0 = v3(p.p1.x - p.p0.x, p.p1.y - p.p0.y, p.p1.z - p.p0.z);
        normalizeVector(v0);
        b200_float3 v1 = v3(p.p2.x - p.p0.x, p.p2.y - p.p0.y, p.p2.z - p.p0.z);
        normalizeVector(v1);
        p.n0 = crossProduct(v0, v1);
        normalizeVector(p.n0);
        p.n1 = p.n0; p.n2 = p.n0;
        break;
    }
    }
    // scene bounds grow with p0 only (GPUKernel.cpp:670-678)
    m_minPos = v3(std::min(x0 * scale, m_minPos.x), std::min(y0 * scale, m_minPos.y), std::min(z0 * scale, m_minPos.z));
    m_maxPos = v3(std::max(x0 * scale, m_maxPos.x), std::max(y0 * scale, m_maxPos.y), std::max(z0 * scale, m_maxPos.z));
}

void SceneHost::setPrimitiveNormals(unsigned index, b200_float3 n0, b200_float3 n1, b200_float3 n2) // GPUKernel.cpp:714-727
{
    if (index < m_primitives.size())
    {
        HostPrimitive& p = m_primitives[index];
        normalizeVector(n0); p.n0 = n0;
        normalizeVector(n1); p.n1 = n1;
        normalizeVector(n2); p.n2 = n2;
        m_primitivesTransfered = false;
    }
}

void SceneHost::setPrimitiveTextureCoordinates(unsigned index, b200_float2 vt0, b200_float2 vt1, b200_float2 vt2) // :703-712
{
    if (index < m_primitives.size())
    {
        HostPrimitive& p = m_primitives[index];
        p.vt0 = vt0; p.vt1 = vt1; p.vt2 = vt2;
        m_primitivesTransfered = false;
    }
}

int SceneHost::addMaterial() { return ++m_nbActiveMaterials; } // GPUKernel.cpp:1761-1767

void SceneHost::setMaterial(unsigned index, const b200_Material& material) // GPUKernel.cpp:1769-1776
{
    if (index < (unsigned)B200_NB_MAX_MATERIALS) { m_hMaterials[index] = material; m_materialsTransfered = false; }
}

// GPUKernel.cpp:1778-1909
void SceneHost::setMaterial(unsigned index, float r, float g, float b, float noise, float reflection, float refraction,
                            bool procedural, bool wireframe, int wireframeWidth, float transparency, float opacity,
                            int diffuseTextureId, int normalTextureId, int bumpTextureId, int specularTextureId,
                            int reflectionTextureId, int transparentTextureId, int ambientOcclusionTextureId, float specValue,
                            float specPower, float specCoef, float innerIllumination, float illuminationDiffusion,
                            float illuminationPropagation, bool fastTransparency)
{
    if (index >= (unsigned)B200_NB_MAX_MATERIALS)
    {
        fprintf(stderr, "[solr_b200] SceneHost::setMaterial: out of bounds (%u)\n", index);
        return;
    }
    b200_Material& m = m_hMaterials[index];
    m.color.x = r; m.color.y = g; m.color.z = b; m.color.w = 0.f;
    m.specular.x = specValue; m.specular.y = specPower; m.specular.z = 0.f; m.specular.w = specCoef;
    m.innerIllumination.x = innerIllumination; m.innerIllumination.y = illuminationDiffusion;
    m.innerIllumination.z = illuminationPropagation; m.innerIllumination.w = noise;
    m.reflection = reflection; m.refraction = refraction; m.transparency = transparency; m.opacity = opacity;
    m.attributes.x = fastTransparency ? 1 : 0;
    m.attributes.y = procedural ? 1 : 0;
    m.attributes.z = wireframe ? ((wireframeWidth == 0) ? 1 : 2) : 0;
    m.attributes.w = wireframeWidth;
    m.textureMapping.x = 1; m.textureMapping.y = 1; m.textureMapping.z = B200_TEXTURE_NONE; m.textureMapping.w = 0;
    m.textureIds.x = diffuseTextureId; m.textureIds.y = normalTextureId; m.textureIds.z = bumpTextureId; m.textureIds.w = specularTextureId;
    m.advancedTextureIds.x = reflectionTextureId; m.advancedTextureIds.y = transparentTextureId;
    m.advancedTextureIds.z = ambientOcclusionTextureId; m.advancedTextureIds.w = B200_TEXTURE_NONE;
    m.advancedTextureOffset.x = m.advancedTextureOffset.y = m.advancedTextureOffset.z = m.advancedTextureOffset.w = 0;
    m.mappingOffset.x = 1.f; m.mappingOffset.y = 1.f;
    auto off = [&](int id) { return (id == B200_TEXTURE_NONE) ? 0 : m_textures[id].offset; };
    if (diffuseTextureId >= 0 && diffuseTextureId < m_nbActiveTextures)
    {
        m.textureMapping.x = m_textures[diffuseTextureId].size.x;
        m.textureMapping.y = m_textures[diffuseTextureId].size.y;
        m.textureMapping.w = m_textures[diffuseTextureId].size.z;
        m.textureMapping.z = B200_TEXTURE_NONE;
        m.textureOffset.x = m_textures[diffuseTextureId].offset;
        m.textureOffset.y = off(normalTextureId);
        m.textureOffset.z = off(bumpTextureId);
        m.textureOffset.w = off(specularTextureId);
        m.advancedTextureOffset.x = off(reflectionTextureId);
        m.advancedTextureOffset.y = off(transparentTextureId);
        m.advancedTextureOffset.z = off(ambientOcclusionTextureId);
    }
    else
    {
        // "computed textures" branch — also taken by every untextured material (:1886-1900)
        m.textureMapping.x = 40000; m.textureMapping.y = 40000; m.textureMapping.z = B200_TEXTURE_NONE; m.textureMapping.w = 3;
        m.textureIds.x = diffuseTextureId; m.textureIds.y = B200_TEXTURE_NONE; m.textureIds.z = B200_TEXTURE_NONE; m.textureIds.w = B200_TEXTURE_NONE;
        m.textureOffset.x = m.textureOffset.y = m.textureOffset.z = m.textureOffset.w = 0;
    }
    m_materialsTransfered = false;
}

// GPUKernel.cpp:2017-2033 + processTextureOffsets :2691-2705
void SceneHost::setTexture(int index, const unsigned char* texels, int width, int height, int depth)
{
    if (index < 0 || index >= B200_NB_MAX_TEXTURES) return;
    if (index >= m_nbActiveTextures) ++m_nbActiveTextures;
    Texture& t = m_textures[index];
    t.texels.assign(texels, texels + (size_t)width * height * depth);
    t.size.x = width; t.size.y = height; t.size.z = depth;
    realignTexturesAndMaterials();
    m_texturesTransfered = false;
}

// GPUKernel.cpp:2238-2340.  The reference also runs this over untextured materials, where it indexes
// m_hTextures[-1] (:2284-2289, out of bounds); those materials never read the fields it writes, so only
// textured and fractal materials are touched here.
void SceneHost::realignTexturesAndMaterials()
{
    int totalSize = 0;
    for (auto& t : m_textures)
    {
        if (!t.texels.empty()) { t.offset = totalSize; totalSize += t.size.x * t.size.y * t.size.z; }
        else t.offset = 0;
    }
    auto off = [&](int id) { return (id == B200_TEXTURE_NONE || id < 0) ? 0 : m_textures[id].offset; };
    for (int i = 0; i < m_nbActiveMaterials; ++i)
    {
        b200_Material& m = m_hMaterials[i];
        const int diffuse = m.textureIds.x;
        if (diffuse == B200_TEXTURE_MANDELBROT || diffuse == B200_TEXTURE_JULIA)
        {
            m.textureMapping.x = 40000; m.textureMapping.y = 40000; m.textureMapping.z = B200_TEXTURE_NONE; m.textureMapping.w = 3;
            m.textureIds.y = m.textureIds.z = m.textureIds.w = B200_TEXTURE_NONE;
            m.textureOffset.x = m.textureOffset.y = m.textureOffset.z = m.textureOffset.w = 0;
            m.advancedTextureIds.x = m.advancedTextureIds.y = m.advancedTextureIds.z = m.advancedTextureIds.w = B200_TEXTURE_NONE;
            m.advancedTextureOffset.x = m.advancedTextureOffset.y = m.advancedTextureOffset.z = m.advancedTextureOffset.w = 0;
        }
        else if (diffuse >= 0 && diffuse < m_nbActiveTextures)
        {
            m.textureMapping.x = m_textures[diffuse].size.x;
            m.textureMapping.y = m_textures[diffuse].size.y;
            m.textureMapping.z = B200_TEXTURE_NONE;
            m.textureMapping.w = m_textures[diffuse].size.z;
            m.textureOffset.x = m_textures[diffuse].offset;
            m.textureOffset.y = off(m.textureIds.y);
            m.textureOffset.z = off(m.textureIds.z);
            m.textureOffset.w = off(m.textureIds.w);
            m.advancedTextureOffset.x = off(m.advancedTextureIds.x);
            m.advancedTextureOffset.y = off(m.advancedTextureIds.y);
            m.mappingOffset.x = 1.f; m.mappingOffset.y = 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// grid hierarchy (GPUKernel.cpp:741-1083)
// ---------------------------------------------------------------------------------------------------
// the box one primitive contributes to its cell (GPUKernel.cpp:752-826)
static void primitiveExtent(const HostPrimitive& primitive, b200_float3& p0, b200_float3& p1)
{
    b200_float3 corner0, corner1;
    switch (primitive.type)
    {
    case B200_PT_TRIANGLE: corner0 = min3(primitive.p0, primitive.p1, primitive.p2); corner1 = max3(primitive.p0, primitive.p1, primitive.p2); break;
    case B200_PT_CYLINDER: corner0 = min2(primitive.p0, primitive.p1); corner1 = max2(primitive.p0, primitive.p1); break;
    default: corner0 = primitive.p0; corner1 = primitive.p0; break;
    }
    p0.x = (corner0.x <= corner1.x) ? corner0.x : corner1.x;
    p0.y = (corner0.y <= corner1.y) ? corner0.y : corner1.y;
    p0.z = (corner0.z <= corner1.z) ? corner0.z : corner1.z;
    p1.x = (corner0.x > corner1.x) ? corner0.x : corner1.x;
    p1.y = (corner0.y > corner1.y) ? corner0.y : corner1.y;
    p1.z = (corner0.z > corner1.z) ? corner0.z : corner1.z;
    switch (primitive.type)
    {
    case B200_PT_CYLINDER:
    case B200_PT_SPHERE:
    case B200_PT_CONE:
        p0.x -= primitive.size.x; p0.y -= primitive.size.x; p0.z -= primitive.size.x;
        p1.x += primitive.size.x; p1.y += primitive.size.x; p1.z += primitive.size.x;
        break;
    default:
        p0.x -= primitive.size.x; p0.y -= primitive.size.y; p0.z -= primitive.size.z;
        p1.x += primitive.size.x; p1.y += primitive.size.y; p1.z += primitive.size.z;
        break;
    }
}

// m_primitives[id] without the tree descent: the hierarchy build asks once per primitive and level-0 box, the flatten once more.
HostPrimitive& SceneHost::primitiveById(unsigned int id)
{
    if (id < m_primitiveTable.size() && m_primitiveTable[id]) return *m_primitiveTable[id];
    return m_primitives[id]; // a missing key is created empty, as in the reference
}

bool SceneHost::updateBoundingBox(HostBox& box) // :741-839
{
    bool result = false;
    box.parameters[0] = v3(1000000, 1000000, 1000000);
    box.parameters[1] = v3(-1000000, -1000000, -1000000);
    for (const auto& p : box.primitives)
    {
        HostPrimitive& primitive = primitiveById((unsigned)p);
        result = (m_hMaterials[primitive.materialId].innerIllumination.x != 0.f);
        b200_float3 p0, p1;
        primitiveExtent(primitive, p0, p1);
        if (p0.x < box.parameters[0].x) box.parameters[0].x = p0.x;
        if (p0.y < box.parameters[0].y) box.parameters[0].y = p0.y;
        if (p0.z < box.parameters[0].z) box.parameters[0].z = p0.z;
        if (p1.x > box.parameters[1].x) box.parameters[1].x = p1.x;
        if (p1.y > box.parameters[1].y) box.parameters[1].y = p1.y;
        if (p1.z > box.parameters[1].z) box.parameters[1].z = p1.z;
    }
    box.center.x = (box.parameters[0].x + box.parameters[1].x) / 2.f;
    box.center.y = (box.parameters[0].y + box.parameters[1].y) / 2.f;
    box.center.z = (box.parameters[0].z + box.parameters[1].z) / 2.f;
    return result;
}

void SceneHost::updateOutterBoundingBox(HostBox& outterBox, int depth) // :841-892
{
    const float vd = m_sceneInfo.viewDistance;
    outterBox.parameters[0] = v3(vd, vd, vd);
    outterBox.parameters[1] = v3(-vd, -vd, -vd);
    const bool linked = outterBox.children.size() == outterBox.primitives.size();
    for (size_t c = 0; c < outterBox.primitives.size(); ++c)
    {
        // operator[]: a missing key is created empty, as in the reference
        HostBox& box = linked ? *outterBox.children[c] : m_boundingBoxes[depth][(unsigned)outterBox.primitives[c]];
        if (outterBox.parameters[0].x > box.parameters[0].x) outterBox.parameters[0].x = box.parameters[0].x;
        if (outterBox.parameters[0].y > box.parameters[0].y) outterBox.parameters[0].y = box.parameters[0].y;
        if (outterBox.parameters[0].z > box.parameters[0].z) outterBox.parameters[0].z = box.parameters[0].z;
        if (outterBox.parameters[1].x < box.parameters[1].x) outterBox.parameters[1].x = box.parameters[1].x;
        if (outterBox.parameters[1].y < box.parameters[1].y) outterBox.parameters[1].y = box.parameters[1].y;
        if (outterBox.parameters[1].z < box.parameters[1].z) outterBox.parameters[1].z = box.parameters[1].z;
    }
    outterBox.center.x = (outterBox.parameters[0].x + outterBox.parameters[1].x) / 2.f;
    outterBox.center.y = (outterBox.parameters[0].y + outterBox.parameters[1].y) / 2.f;
    outterBox.center.z = (outterBox.parameters[0].z + outterBox.parameters[1].z) / 2.f;
}

void SceneHost::resetBoxes(bool resetPrimitives) // :894-901
{
    materialiseBoxes();
    if (resetPrimitives)
        for (size_t i = 0; i < m_boundingBoxes[0].size(); ++i) resetBox(m_boundingBoxes[0][(unsigned)i], resetPrimitives);
    else
        m_boundingBoxes[0].clear();
}

void SceneHost::resetBox(HostBox& box, bool resetPrimitives) // :903-917
{
    if (resetPrimitives) { box.primitives.clear(); box.children.clear(); box.indexForNextBox = 1; }
    const float vd = m_sceneInfo.viewDistance;
    box.parameters[0] = v3(vd, vd, vd);
    box.parameters[1] = v3(-vd, -vd, -vd);
}

void SceneHost::processBoxes(const int boxSize) // :919-992 (simulate == false)
{
    b200_float3 boxSteps;
    boxSteps.x = (m_maxPos.x - m_minPos.x) / boxSize;
    boxSteps.y = (m_maxPos.y - m_minPos.y) / boxSize;
    boxSteps.z = (m_maxPos.z - m_minPos.z) / boxSize;
    boxSteps.x = (boxSteps.x == 0.f) ? 1 : boxSteps.x;
    boxSteps.y = (boxSteps.y == 0.f) ? 1 : boxSteps.y;
    boxSteps.z = (boxSteps.z == 0.f) ? 1 : boxSteps.z;
    const float vd = m_sceneInfo.viewDistance;
    // The reference inserts into a std::map per primitive (find, insert, operator[]).  The same map results from taking the
    // cell keys in iteration order, ordering them by key with a STABLE sort (a box lists its primitives in iteration order)
    // and merging them into the map in one ascending sweep; lights go to the first box of the top level as they come.
    struct Entry { unsigned int key; unsigned int p; bool light; };
    std::vector<Entry> entries;
    entries.reserve(m_primitives.size());
    unsigned int p = 0;
    for (const auto& prim : m_primitives)
    {
        const HostPrimitive& primitive = prim.second;
        const b200_float3& center = primitive.p0;
        // unsigned arithmetic: the key wraps modulo 2^32 for X >= 105 (:938-941)
        unsigned int X = static_cast<int>((center.x - m_minPos.x) / boxSteps.x);
        unsigned int Y = static_cast<int>((center.y - m_minPos.y) / boxSteps.y);
        unsigned int Z = static_cast<int>((center.z - m_minPos.z) / boxSteps.z);
        unsigned int B = 1 + 1000 * (X * boxSize * boxSize + Y * boxSize + Z);
        const bool light = m_hMaterials[primitive.materialId].innerIllumination.x != 0.f;
        if (light) m_boundingBoxes[m_treeDepth][0].primitives.push_back(p); // lights: first box of the top level
        entries.push_back({B, p, light});
        ++p;
    }
    std::stable_sort(entries.begin(), entries.end(), [](const Entry& a, const Entry& b) { return a.key < b.key; });
    auto& level0 = m_boundingBoxes[0];
    auto it = level0.begin();
    for (const Entry& e : entries)
    {
        while (it != level0.end() && it->first < e.key) ++it;
        if (it == level0.end() || it->first != e.key)
        {
            HostBox box;
            box.parameters[0] = v3(vd, vd, vd);
            box.parameters[1] = v3(-vd, -vd, -vd);
            box.center = v3(0.f, 0.f, 0.f);
            box.indexForNextBox = 1;
            it = level0.emplace_hint(it, e.key, std::move(box)); // the cell exists even if all it received was a light
        }
        if (!e.light) it->second.primitives.push_back(e.p);
    }
    for (auto& box : m_boundingBoxes[0]) updateBoundingBox(box.second);
}

void SceneHost::processOutterBoxes(const int boxSize, const int depth) // :994-1039
{
    b200_float3 boxSteps;
    boxSteps.x = (m_maxPos.x - m_minPos.x) / boxSize;
    boxSteps.y = (m_maxPos.y - m_minPos.y) / boxSize;
    boxSteps.z = (m_maxPos.z - m_minPos.z) / boxSize;
    boxSteps.x = (boxSteps.x == 0.f) ? 1 : boxSteps.x;
    boxSteps.y = (boxSteps.y == 0.f) ? 1 : boxSteps.y;
    boxSteps.z = (boxSteps.z == 0.f) ? 1 : boxSteps.z;
    const float vd = m_sceneInfo.viewDistance;
    // as in processBoxes: keys in iteration order, stable sort, one ascending merge into the level's map
    struct Entry { unsigned int key; unsigned int child; HostBox* box; };
    std::vector<Entry> entries;
    entries.reserve(m_boundingBoxes[depth - 1].size());
    for (auto& box : m_boundingBoxes[depth - 1])
    {
        const b200_float3& center = box.second.center;
        int X = static_cast<int>((center.x - m_minPos.x) / boxSteps.x);
        int Y = static_cast<int>((center.y - m_minPos.y) / boxSteps.y);
        int Z = static_cast<int>((center.z - m_minPos.z) / boxSteps.z);
        // int arithmetic in the reference (:1012-1014); it overflows for large grids, so wrap explicitly
        const uint32_t bs = (uint32_t)boxSize;
        uint32_t Bu = (uint32_t)X * bs * bs + (uint32_t)Y * bs + (uint32_t)Z;
        Bu += 1; // key 0 holds the lights
        entries.push_back({Bu, box.first, &box.second});
    }
    std::stable_sort(entries.begin(), entries.end(), [](const Entry& a, const Entry& b) { return a.key < b.key; });
    auto& level = m_boundingBoxes[depth];
    auto it = level.begin();
    for (const Entry& e : entries)
    {
        while (it != level.end() && it->first < e.key) ++it;
        if (it == level.end() || it->first != e.key) it = level.emplace_hint(it, e.key, HostBox());
        HostBox& ob = it->second;
        ob.parameters[0] = v3(vd, vd, vd);
        ob.parameters[1] = v3(-vd, -vd, -vd);
        ob.primitives.push_back(e.child);
        ob.children.push_back(e.box);
    }
    for (auto& box : m_boundingBoxes[depth]) updateOutterBoundingBox(box.second, depth - 1);
}


// ---------------------------------------------------------------------------------------------------
// flat build: the first compaction of a fresh container without node-based containers
// ---------------------------------------------------------------------------------------------------
// GPUKernel::compactBoxes builds one std::map of boxes per level and flattens them by recursion with a map lookup per box
// (GPUKernel.cpp:919-1149).  On a fresh container everything it computes is a function of sorted sequences: a level's boxes
// are the distinct cell keys of the level below in ascending order, a box lists its children in ascending key order, its
// bounds are the min / max over them.  So each level is one stable sort of (cell key, child) pairs and one pass over the runs,
// the flattening a recursion over index ranges — the same arrays, byte for byte (tests/test_scene_host.py compares both paths
// and the reference's own builder), in a form that is also the outline of a GPU build (sort by key, segmented reduce, scan).
// What the reference leaves behind in its maps for later calls — including the empty box its first resetBox() creates at
// level 2, and the empty boxes its look-ups of light ids as box keys create one level below the top — is materialised from
// the flat arrays if and when a later call needs it.
struct SceneHost::FlatHierarchy
{
    struct Level
    {
        struct Box { b200_float3 lo, hi; unsigned int first, count; }; // children of the box: items[first .. first + count)
        std::vector<unsigned int> keys;         // ascending
        std::vector<Box> boxes;                 // one 32-byte record per box: the flattening touches one cache line per box
        std::vector<b200_float3> center;
        std::vector<long> items;                // level 0: primitive ids; level d: keys of boxes of level d - 1
        std::vector<unsigned int> itemIndex;    // level d >= 1: where that box is in level d - 1
        void reserve(size_t nBoxes, size_t nItems, bool index)
        {
            keys.reserve(nBoxes); boxes.reserve(nBoxes); center.reserve(nBoxes); items.reserve(nItems);
            if (index) itemIndex.reserve(nItems);
        }
        void add(unsigned int key, const b200_float3& lo, const b200_float3& hi, unsigned int first)
        {
            keys.push_back(key);
            boxes.push_back({lo, hi, first, (unsigned int)items.size() - first});
            center.push_back(b200_float3{(lo.x + hi.x) / 2.f, (lo.y + hi.y) / 2.f, (lo.z + hi.z) / 2.f});
        }
    };
    std::vector<Level> levels;                  // 0 .. depth
    std::vector<long> lights;                   // primitives of the first box of the top level
    std::vector<unsigned int> lightKeyIndex;    // per light id: its index in level depth - 1 when read as a box key, or ~0u
    b200_float3 lightsLo, lightsHi, lightsCenter;
    unsigned int oldDepth = 0;                  // level of the empty box resetBox() created before the build
    unsigned int depth = 0;
};

void SceneHost::dropFlat()
{
    delete m_flat;
    m_flat = nullptr;
}

bool SceneHost::flatBuildApplies() const
{
    if (m_flat || !m_useFlatBuild) return false;
    for (const auto& level : m_boundingBoxes)
        if (!level.empty()) return false;
    unsigned int depth = 0;
    for (int nb = static_cast<int>(m_primitives.size()); nb > 2; nb /= 4) ++depth;
    if (m_treeDepth < 1 || m_treeDepth >= depth) return false;        // the empty box of the old top level lies inside the new tree
    if (m_primitiveTable.size() != m_primitives.size()) return false; // ids 0 .. n-1
    bool light = false;
    for (const auto& prim : m_primitives)
        if (m_hMaterials[prim.second.materialId].innerIllumination.x != 0.f) { light = true; break; }
    return light; // without a light the reference flattens its first ordinary box as if it held the lights
}

bool SceneHost::flatBuild()
{
    FlatHierarchy* F = new FlatHierarchy();
    m_flat = F;
    const float vd = m_sceneInfo.viewDistance;
    const size_t n = m_primitives.size();
    F->oldDepth = m_treeDepth;
    const unsigned int F_oldDepth = m_treeDepth;
    const int gridGranularity = 2, gridDivider = 4;
    unsigned int depth = 0;
    for (int nb = static_cast<int>(n); nb > gridGranularity; nb /= gridDivider) ++depth;
    F->depth = depth;
    F->levels.resize(depth + 1);

    struct Entry { unsigned int key; unsigned int child; };
    std::vector<Entry> entries;
    // ---- level 0: primitives into cells (processBoxes) ----
    {
        const int boxSize = AABB_MAGIC_NUMBER;
        b200_float3 boxSteps;
        boxSteps.x = (m_maxPos.x - m_minPos.x) / boxSize;
        boxSteps.y = (m_maxPos.y - m_minPos.y) / boxSize;
        boxSteps.z = (m_maxPos.z - m_minPos.z) / boxSize;
        boxSteps.x = (boxSteps.x == 0.f) ? 1 : boxSteps.x;
        boxSteps.y = (boxSteps.y == 0.f) ? 1 : boxSteps.y;
        boxSteps.z = (boxSteps.z == 0.f) ? 1 : boxSteps.z;
        entries.resize(n);
        std::vector<unsigned char> isLight(n, 0);
        std::vector<b200_float3> extentLo(n), extentHi(n);
        parallelFor(n, [&](size_t begin, size_t end) {
            for (size_t p = begin; p < end; ++p)
            {
                const HostPrimitive& primitive = *m_primitiveTable[p];
                const b200_float3& center = primitive.p0;
                // unsigned arithmetic: the key wraps modulo 2^32 for X >= 105 (:938-941)
                unsigned int X = static_cast<int>((center.x - m_minPos.x) / boxSteps.x);
                unsigned int Y = static_cast<int>((center.y - m_minPos.y) / boxSteps.y);
                unsigned int Z = static_cast<int>((center.z - m_minPos.z) / boxSteps.z);
                unsigned int B = 1 + 1000 * (X * boxSize * boxSize + Y * boxSize + Z);
                isLight[p] = m_hMaterials[primitive.materialId].innerIllumination.x != 0.f;
                entries[p] = {B, (unsigned int)p};
                primitiveExtent(primitive, extentLo[p], extentHi[p]);
            }
        });
        for (unsigned int p = 0; p < n; ++p)
            if (isLight[p]) F->lights.push_back(p);
        parallelStableSortByKey(entries);
        FlatHierarchy::Level& L = F->levels[0];
        L.reserve(entries.size(), entries.size(), false);
        for (size_t i = 0; i < entries.size();)
        {
            size_t j = i;
            const unsigned int first = (unsigned int)L.items.size();
            b200_float3 lo = v3(1000000, 1000000, 1000000), hi = v3(-1000000, -1000000, -1000000);
            for (; j < entries.size() && entries[j].key == entries[i].key; ++j)
            {
                if (isLight[entries[j].child]) continue; // the cell exists, the light itself goes to the top level
                L.items.push_back(entries[j].child);
                const b200_float3& p0 = extentLo[entries[j].child];
                const b200_float3& p1 = extentHi[entries[j].child];
                if (p0.x < lo.x) lo.x = p0.x;
                if (p0.y < lo.y) lo.y = p0.y;
                if (p0.z < lo.z) lo.z = p0.z;
                if (p1.x > hi.x) hi.x = p1.x;
                if (p1.y > hi.y) hi.y = p1.y;
                if (p1.z > hi.z) hi.z = p1.z;
            }
            L.add(entries[i].key, lo, hi, first);
            i = j;
        }
    }
    // ---- levels 1 .. depth: boxes of the level below into coarser cells (processOutterBoxes) ----
    int boxSize = static_cast<int>(n);
    for (unsigned int d = 1; d <= depth; ++d, boxSize /= gridDivider)
    {
        const FlatHierarchy::Level& below = F->levels[d - 1];
        FlatHierarchy::Level& L = F->levels[d];
        b200_float3 boxSteps;
        boxSteps.x = (m_maxPos.x - m_minPos.x) / boxSize;
        boxSteps.y = (m_maxPos.y - m_minPos.y) / boxSize;
        boxSteps.z = (m_maxPos.z - m_minPos.z) / boxSize;
        boxSteps.x = (boxSteps.x == 0.f) ? 1 : boxSteps.x;
        boxSteps.y = (boxSteps.y == 0.f) ? 1 : boxSteps.y;
        boxSteps.z = (boxSteps.z == 0.f) ? 1 : boxSteps.z;
        entries.resize(below.keys.size());
        parallelFor(below.keys.size(), [&](size_t begin, size_t end) {
            for (size_t c = begin; c < end; ++c)
            {
                const b200_float3& center = below.center[c];
                int X = static_cast<int>((center.x - m_minPos.x) / boxSteps.x);
                int Y = static_cast<int>((center.y - m_minPos.y) / boxSteps.y);
                int Z = static_cast<int>((center.z - m_minPos.z) / boxSteps.z);
                // int arithmetic in the reference (:1012-1014); it overflows for large grids, so wrap explicitly
                const uint32_t bs = (uint32_t)boxSize;
                uint32_t Bu = (uint32_t)X * bs * bs + (uint32_t)Y * bs + (uint32_t)Z;
                Bu += 1; // key 0 holds the lights
                entries[c] = {Bu, (unsigned int)c};
            }
        });
        parallelStableSortByKey(entries);
        if (!entries.empty() && entries[0].key == 0u)
        {
            // a cell key wrapped onto the key of the lights box: leave this scene to the literal path
            delete F;
            m_flat = nullptr;
            m_treeDepth = F_oldDepth;
            return false;
        }
        // the empty box resetBox() left at key 0 of the old top level is there before the level's own boxes
        L.reserve(entries.size() + 1, entries.size(), true);
        if (d == F->oldDepth) L.add(0u, v3(vd, vd, vd), v3(-vd, -vd, -vd), 0u);
        for (size_t i = 0; i < entries.size();)
        {
            size_t j = i;
            const unsigned int first = (unsigned int)L.items.size();
            b200_float3 lo = v3(vd, vd, vd), hi = v3(-vd, -vd, -vd);
            for (; j < entries.size() && entries[j].key == entries[i].key; ++j)
            {
                const unsigned int c = entries[j].child;
                L.items.push_back((long)below.keys[c]);
                L.itemIndex.push_back(c);
                const FlatHierarchy::Level::Box& cb = below.boxes[c];
                if (lo.x > cb.lo.x) lo.x = cb.lo.x;
                if (lo.y > cb.lo.y) lo.y = cb.lo.y;
                if (lo.z > cb.lo.z) lo.z = cb.lo.z;
                if (hi.x < cb.hi.x) hi.x = cb.hi.x;
                if (hi.y < cb.hi.y) hi.y = cb.hi.y;
                if (hi.z < cb.hi.z) hi.z = cb.hi.z;
            }
            L.add(entries[i].key, lo, hi, first);
            i = j;
        }
    }
    // ---- the lights box (key 0 of the top level): its "children" are light ids read as box keys one level down ----
    {
        const FlatHierarchy::Level& below = F->levels[depth - 1];
        b200_float3 lo = v3(vd, vd, vd), hi = v3(-vd, -vd, -vd);
        for (long id : F->lights)
        {
            const auto it = std::lower_bound(below.keys.begin(), below.keys.end(), (unsigned int)id);
            const bool found = it != below.keys.end() && *it == (unsigned int)id;
            const unsigned int idx = found ? (unsigned int)(it - below.keys.begin()) : ~0u;
            F->lightKeyIndex.push_back(idx);
            const b200_float3 blo = found ? below.boxes[idx].lo : v3(0.f, 0.f, 0.f), bhi = found ? below.boxes[idx].hi : v3(0.f, 0.f, 0.f);
            if (lo.x > blo.x) lo.x = blo.x;
            if (lo.y > blo.y) lo.y = blo.y;
            if (lo.z > blo.z) lo.z = blo.z;
            if (hi.x < bhi.x) hi.x = bhi.x;
            if (hi.y < bhi.y) hi.y = bhi.y;
            if (hi.z < bhi.z) hi.z = bhi.z;
        }
        F->lightsLo = lo; F->lightsHi = hi;
        F->lightsCenter = v3((lo.x + hi.x) / 2.f, (lo.y + hi.y) / 2.f, (lo.z + hi.z) / 2.f);
    }
    m_treeDepth = depth;
    return true;
}

void SceneHost::flatRecurse(const int depth, const unsigned int j) // recursiveDataStreamToGPU for one box
{
    const FlatHierarchy::Level& L = m_flat->levels[depth];
    const FlatHierarchy::Level::Box& B = L.boxes[j];
    const unsigned int begin = B.first, end = B.first + B.count;
    if (begin == end || (size_t)m_nbActiveBoxes >= m_maxBoxes) return;
    const int boxIndex = m_nbActiveBoxes;
    b200_BoundingBox out;
    memset(&out, 0, sizeof(out));
    out.parameters[0] = B.lo;
    out.parameters[1] = B.hi;
    out.nbPrimitives = (depth == 0) ? static_cast<int>(end - begin) : 0;
    out.startIndex = (depth == 0) ? m_nbActivePrimitives : depth;
    m_hBoundingBoxes.push_back(out);
    ++m_nbActiveBoxes;
    if (depth == 0)
    {
        for (unsigned int i = begin; i < end; ++i)
        {
            const long id = L.items[i];
            if ((size_t)id < m_maxPrimitives && (size_t)m_nbActivePrimitives < m_maxPrimitives) emitPrimitive(id);
        }
    }
    else
        for (unsigned int i = begin; i < end; ++i) flatRecurse(depth - 1, L.itemIndex[i]);
    m_hBoundingBoxes[boxIndex].indexForNextBox.x = (depth == 0) ? 1 : m_nbActiveBoxes - boxIndex;
}

void SceneHost::flatStream() // streamDataToGPU over the flat levels
{
    const FlatHierarchy& F = *m_flat;
    m_primitivesTransfered = false;
    m_nbActiveBoxes = 0; m_nbActivePrimitives = 0; m_nbActiveLamps = 0;
    m_hBoundingBoxes.clear(); m_hPrimitives.clear(); m_hLamps.clear();
    {
        size_t boxes = 1;
        for (const auto& L : F.levels) boxes += L.keys.size();
        m_hBoundingBoxes.reserve(std::min(boxes, m_maxBoxes + 1));
        m_hPrimitives.reserve(std::min(m_primitives.size(), m_maxPrimitives));
    }
    const float vd = m_sceneInfo.viewDistance;
    const int maxDepth = (int)F.depth;
    // box 0 of the top level holds the lights, with bounds +-viewDistance (:1177-1190)
    {
        const int boxIndex = m_nbActiveBoxes;
        b200_BoundingBox out;
        memset(&out, 0, sizeof(out));
        m_lightInformationSize = 0;
        m_lightInformation.clear();
        out.parameters[0] = v3(-vd, -vd, -vd);
        out.parameters[1] = v3(vd, vd, vd);
        out.nbPrimitives = static_cast<int>(F.lights.size());
        out.startIndex = 0;
        m_hBoundingBoxes.push_back(out);
        for (long id : F.lights)
        {
            emitPrimitive(id);
            const HostPrimitive& primitive = primitiveById((unsigned)id);
            const b200_Material& material = m_hMaterials[primitive.materialId];
            b200_LightInformation li;
            memset(&li, 0, sizeof(li));
            li.primitiveId = (int)id;
            li.materialId = primitive.materialId;
            li.location = primitive.p0;
            li.color.x = material.color.x; li.color.y = material.color.y; li.color.z = material.color.z;
            li.color.w = material.innerIllumination.x;
            if (m_lightInformationSize < B200_NB_MAX_LIGHTINFORMATIONS) m_lightInformation.push_back(li);
            if (m_nbActiveLamps < NB_MAX_LAMPS) m_hLamps.push_back((int)id);
            ++m_nbActiveLamps;
            ++m_lightInformationSize;
        }
        ++m_nbActiveBoxes;
        // the reference recurses into the lights box too, reading the light ids as keys of the level below
        for (unsigned int idx : F.lightKeyIndex)
            if (idx != ~0u) flatRecurse(maxDepth - 1, idx);
        m_hBoundingBoxes[boxIndex].indexForNextBox.x = m_nbActiveBoxes - boxIndex;
    }
    const FlatHierarchy::Level& top = F.levels[maxDepth];
    if (m_levelOrderFlatten && flatStreamByLevels())
    {
        if ((size_t)m_nbActivePrimitives != m_primitives.size())
            fprintf(stderr, "[solr_b200] compactBoxes: lost primitives on the way... %d != %zu\n", m_nbActivePrimitives, m_primitives.size());
        return;
    }
    for (unsigned int j = 0; j < top.keys.size(); ++j)
    {
        const int boxIndex = m_nbActiveBoxes;
        b200_BoundingBox out;
        memset(&out, 0, sizeof(out));
        out.parameters[0] = top.boxes[j].lo;
        out.parameters[1] = top.boxes[j].hi;
        out.nbPrimitives = 0;
        out.startIndex = maxDepth;
        m_hBoundingBoxes.push_back(out);
        ++m_nbActiveBoxes;
        for (unsigned int i = top.boxes[j].first; i < top.boxes[j].first + top.boxes[j].count; ++i) flatRecurse(maxDepth - 1, top.itemIndex[i]);
        m_hBoundingBoxes[boxIndex].indexForNextBox.x = m_nbActiveBoxes - boxIndex;
    }
    if ((size_t)m_nbActivePrimitives != m_primitives.size())
        fprintf(stderr, "[solr_b200] compactBoxes: lost primitives on the way... %d != %zu\n", m_nbActivePrimitives, m_primitives.size());
}

// The depth-first flatten above is a chain of dependent, scattered reads (a box, then its children, then theirs).  The same arrays
// follow from three passes per level whose accesses are independent of each other: bottom-up the number of boxes and primitives
// each subtree emits, top-down every box's slot (its parent's slot + 1 + the sizes of the siblings before it) and its first
// primitive, then every record written straight to its slot.  Applies when nothing gets truncated by the capacities and no
// light id doubles as a box key (the lights box is already out when this runs); otherwise the recursion does it.  Sizes, scan,
// scatter: this is also the shape of a GPU flatten.
bool SceneHost::flatStreamByLevels()
{
    const FlatHierarchy& F = *m_flat;
    const int D = (int)F.depth;
    for (unsigned int idx : F.lightKeyIndex)
        if (idx != ~0u) return false;
    if (m_primitives.size() > m_maxPrimitives) return false;
    std::vector<std::vector<unsigned int>> size(D + 1), prims(D + 1), slot(D + 1), start(D + 1);
    // ---- bottom-up: what every subtree emits ----
    for (int d = 0; d <= D; ++d)
    {
        const FlatHierarchy::Level& L = F.levels[d];
        const size_t nb = L.boxes.size();
        size[d].resize(nb); prims[d].resize(nb);
        parallelFor(nb, [&](size_t begin, size_t end) {
            for (size_t j = begin; j < end; ++j)
            {
                const FlatHierarchy::Level::Box& B = L.boxes[j];
                if (d == 0) { size[0][j] = B.count ? 1u : 0u; prims[0][j] = B.count; continue; }
                unsigned int sz = 0, pr = 0;
                if (B.count || d == D) // an inner box without children emits nothing; a top-level box always does
                {
                    sz = 1;
                    for (unsigned int i = B.first; i < B.first + B.count; ++i) { sz += size[d - 1][L.itemIndex[i]]; pr += prims[d - 1][L.itemIndex[i]]; }
                }
                size[d][j] = sz; prims[d][j] = pr;
            }
        });
    }
    // ---- the top level in order, after the lights box ----
    const FlatHierarchy::Level& top = F.levels[D];
    slot[D].resize(top.boxes.size()); start[D].resize(top.boxes.size());
    size_t boxes = (size_t)m_nbActiveBoxes, primitives = (size_t)m_nbActivePrimitives;
    for (size_t j = 0; j < top.boxes.size(); ++j)
    {
        slot[D][j] = (unsigned int)boxes; start[D][j] = (unsigned int)primitives;
        boxes += size[D][j]; primitives += prims[D][j];
    }
    if (boxes > m_maxBoxes || primitives > m_maxPrimitives || primitives != m_primitives.size()) return false;
    b200_BoundingBox zeroBox;
    memset(&zeroBox, 0, sizeof(zeroBox));
    m_hBoundingBoxes.resize(boxes, zeroBox);
    b200_Primitive zeroPrimitive;
    memset(&zeroPrimitive, 0, sizeof(zeroPrimitive));
    m_hPrimitives.resize(primitives, zeroPrimitive);
    // ---- top-down: slots of the children, records of the level ----
    for (int d = D; d >= 0; --d)
    {
        const FlatHierarchy::Level& L = F.levels[d];
        if (d > 0) { slot[d - 1].assign(F.levels[d - 1].boxes.size(), 0u); start[d - 1].assign(F.levels[d - 1].boxes.size(), 0u); }
        // every box has one parent: the slots written below, and the records written at them, are all distinct
        parallelFor(L.boxes.size(), [&](size_t begin, size_t end) {
            for (size_t j = begin; j < end; ++j)
            {
                if (size[d][j] == 0) continue;
                const FlatHierarchy::Level::Box& B = L.boxes[j];
                b200_BoundingBox& out = m_hBoundingBoxes[slot[d][j]];
                out.parameters[0] = B.lo;
                out.parameters[1] = B.hi;
                if (d == 0)
                {
                    out.nbPrimitives = (int)B.count;
                    out.startIndex = (int)start[0][j];
                    out.indexForNextBox.x = 1;
                    for (unsigned int i = 0; i < B.count; ++i) writePrimitive(start[0][j] + i, L.items[B.first + i]);
                    continue;
                }
                out.nbPrimitives = 0;
                out.startIndex = d;
                out.indexForNextBox.x = (int)size[d][j];
                unsigned int s = slot[d][j] + 1, p = start[d][j];
                for (unsigned int i = B.first; i < B.first + B.count; ++i)
                {
                    const unsigned int c = L.itemIndex[i];
                    slot[d - 1][c] = s; start[d - 1][c] = p;
                    s += size[d - 1][c]; p += prims[d - 1][c];
                }
            }
        });
        size[d].clear(); size[d].shrink_to_fit();
    }
    m_nbActiveBoxes = (int)boxes;
    m_nbActivePrimitives = (int)primitives;
    return true;
}

// The maps as the reference's own build would have left them, for whatever is called next.
void SceneHost::materialiseBoxes()
{
    if (!m_flat) return;
    FlatHierarchy* F = m_flat;
    m_flat = nullptr;
    std::vector<HostBox*> belowPtr, ptr;
    for (unsigned int d = 0; d <= F->depth; ++d)
    {
        const FlatHierarchy::Level& L = F->levels[d];
        auto& level = m_boundingBoxes[d];
        ptr.assign(L.keys.size(), nullptr);
        if (d == F->depth)
        {
            // the lights box: created by operator[] (all zero), bounds from the look-ups of its light ids
            HostBox box;
            memset(box.parameters, 0, sizeof(box.parameters));
            box.parameters[0] = F->lightsLo; box.parameters[1] = F->lightsHi; box.center = F->lightsCenter;
            box.indexForNextBox = 0;
            box.primitives = F->lights;
            level.emplace_hint(level.end(), 0u, std::move(box));
        }
        for (unsigned int j = 0; j < L.keys.size(); ++j)
        {
            HostBox box;
            box.parameters[0] = L.boxes[j].lo; box.parameters[1] = L.boxes[j].hi; box.center = L.center[j];
            box.indexForNextBox = (d == 0 || (d == F->oldDepth && L.keys[j] == 0u)) ? 1 : 0;
            const unsigned int begin = L.boxes[j].first, end = begin + L.boxes[j].count;
            box.primitives.assign(L.items.begin() + begin, L.items.begin() + end);
            if (d > 0)
                for (unsigned int i = begin; i < end; ++i) box.children.push_back(belowPtr[L.itemIndex[i]]);
            auto it = level.emplace_hint(level.end(), L.keys[j], std::move(box));
            ptr[j] = &it->second;
        }
        belowPtr.swap(ptr);
    }
    // light ids looked up as box keys one level below the top: a missing key was created empty by operator[]
    for (size_t i = 0; i < F->lights.size(); ++i)
        if (F->lightKeyIndex[i] == ~0u) m_boundingBoxes[F->depth - 1][(unsigned int)F->lights[i]];
    delete F;
}

int SceneHost::compactBoxes(bool reconstructBoxes) // :1041-1083
{
    if (m_hostStale && !reconstructBoxes) return m_nbActiveBoxes; // the step was applied on the device, where the arrays already are
    syncFromDevice();
    m_primitivesTransfered = false;
    m_primitiveTable.clear();
    if (!m_primitives.empty() && m_primitives.rbegin()->first < 4u * m_primitives.size() + 1024u)
    {
        m_primitiveTable.assign((size_t)m_primitives.rbegin()->first + 1, nullptr);
        for (auto& prim : m_primitives) m_primitiveTable[prim.first] = &prim.second;
    }
    if (reconstructBoxes && flatBuildApplies() && flatBuild())
    {
        flatStream();
        m_primitiveTable.clear();
        return m_nbActiveBoxes;
    }
    materialiseBoxes();
    if (reconstructBoxes)
    {
        resetBox(m_boundingBoxes[m_treeDepth][0], true);
        const int gridGranularity = 2, gridDivider = 4;
        m_treeDepth = 0;
        int nbBoxes = static_cast<int>(m_primitives.size());
        while (nbBoxes > gridGranularity) { ++m_treeDepth; nbBoxes /= gridDivider; }
        processBoxes(AABB_MAGIC_NUMBER);
        m_treeDepth = 0;
        nbBoxes = static_cast<int>(m_primitives.size());
        do
        {
            ++m_treeDepth;
            processOutterBoxes(nbBoxes, m_treeDepth);
            nbBoxes /= gridDivider;
        } while (nbBoxes > gridGranularity);
    }
    streamDataToGPU();
    m_primitiveTable.clear();
    return m_nbActiveBoxes;
}

void SceneHost::writePrimitive(size_t slot, long id) // the record emitPrimitive appends, written in place
{
    const HostPrimitive& primitive = primitiveById((unsigned)id);
    b200_Primitive& out = m_hPrimitives[slot]; // zero-filled by the caller
    out.index = (int)id;
    out.type = primitive.type;
    out.p0 = primitive.p0; out.p1 = primitive.p1; out.p2 = primitive.p2;
    out.n0 = primitive.n0; out.n1 = primitive.n1; out.n2 = primitive.n2;
    out.size = primitive.size;
    out.materialId = primitive.materialId;
    out.vt0 = primitive.vt0; out.vt1 = primitive.vt1; out.vt2 = primitive.vt2;
}

void SceneHost::emitPrimitive(long id) // :1116-1135, :1199-1212
{
    HostPrimitive& primitive = primitiveById((unsigned)id);
    b200_Primitive out;
    memset(&out, 0, sizeof(out));
    out.index = (int)id;
    out.type = primitive.type;
    out.p0 = primitive.p0; out.p1 = primitive.p1; out.p2 = primitive.p2;
    out.n0 = primitive.n0; out.n1 = primitive.n1; out.n2 = primitive.n2;
    out.size = primitive.size;
    out.materialId = primitive.materialId;
    out.vt0 = primitive.vt0; out.vt1 = primitive.vt1; out.vt2 = primitive.vt2;
    m_hPrimitives.push_back(out);
    ++m_nbActivePrimitives;
}

void SceneHost::recursiveDataStreamToGPU(const int depth, std::vector<long>& elements, const std::vector<HostBox*>* linked) // :1085-1149
{
    for (size_t c = 0; c < elements.size(); ++c)
    {
        HostBox& box = linked ? *(*linked)[c] : m_boundingBoxes[depth][(unsigned)elements[c]];
        if (box.primitives.size() != 0 && (size_t)m_nbActiveBoxes < m_maxBoxes)
        {
            const int boxIndex = m_nbActiveBoxes;
            b200_BoundingBox out;
            memset(&out, 0, sizeof(out));
            out.parameters[0] = box.parameters[0];
            out.parameters[1] = box.parameters[1];
            out.nbPrimitives = (depth == 0) ? static_cast<int>(box.primitives.size()) : 0;
            out.startIndex = (depth == 0) ? m_nbActivePrimitives : depth;
            m_hBoundingBoxes.push_back(out);
            ++m_nbActiveBoxes;
            if (depth == 0)
            {
                for (long id : box.primitives)
                    if ((size_t)id < m_maxPrimitives && (size_t)m_nbActivePrimitives < m_maxPrimitives) emitPrimitive(id);
            }
            else
                recursiveDataStreamToGPU(depth - 1, box.primitives, box.children.size() == box.primitives.size() ? &box.children : nullptr);
            m_hBoundingBoxes[boxIndex].indexForNextBox.x = (depth == 0) ? 1 : m_nbActiveBoxes - boxIndex;
        }
    }
}

void SceneHost::streamDataToGPU() // :1151-1281
{
    m_primitivesTransfered = false;
    m_nbActiveBoxes = 0; m_nbActivePrimitives = 0; m_nbActiveLamps = 0;
    m_hBoundingBoxes.clear(); m_hPrimitives.clear(); m_hLamps.clear();
    const float vd = m_sceneInfo.viewDistance;
    const int maxDepth = (int)m_treeDepth;
    bool first = true;
    for (auto& entry : m_boundingBoxes[maxDepth])
    {
        HostBox& box = entry.second;
        const int boxIndex = m_nbActiveBoxes;
        b200_BoundingBox out;
        memset(&out, 0, sizeof(out));
        out.parameters[0] = box.parameters[0];
        out.parameters[1] = box.parameters[1];
        out.nbPrimitives = 0;
        out.startIndex = maxDepth;
        if (first)
        {
            // box 0 of the top level holds the lights, with bounds +-viewDistance (:1177-1190)
            m_lightInformationSize = 0;
            m_lightInformation.clear();
            out.parameters[0] = v3(-vd, -vd, -vd);
            out.parameters[1] = v3(vd, vd, vd);
            out.nbPrimitives = static_cast<int>(box.primitives.size());
            out.startIndex = 0;
            m_hBoundingBoxes.push_back(out);
            for (long id : box.primitives)
            {
                emitPrimitive(id);
                const HostPrimitive& primitive = primitiveById((unsigned)id);
                const b200_Material& material = m_hMaterials[primitive.materialId];
                b200_LightInformation li;
                memset(&li, 0, sizeof(li));
                li.primitiveId = (int)id;
                li.materialId = primitive.materialId;
                li.location = primitive.p0;
                li.color.x = material.color.x; li.color.y = material.color.y; li.color.z = material.color.z;
                li.color.w = material.innerIllumination.x;
                if (m_lightInformationSize < B200_NB_MAX_LIGHTINFORMATIONS) m_lightInformation.push_back(li);
                if (m_nbActiveLamps < NB_MAX_LAMPS) m_hLamps.push_back((int)id);
                ++m_nbActiveLamps;
                ++m_lightInformationSize;
            }
        }
        else
            m_hBoundingBoxes.push_back(out);
        first = false;
        ++m_nbActiveBoxes;
        if (maxDepth > 0)
            recursiveDataStreamToGPU(maxDepth - 1, box.primitives, box.children.size() == box.primitives.size() ? &box.children : nullptr);
        m_hBoundingBoxes[boxIndex].indexForNextBox.x = m_nbActiveBoxes - boxIndex;
    }
    if ((size_t)m_nbActivePrimitives != m_primitives.size())
        fprintf(stderr, "[solr_b200] compactBoxes: lost primitives on the way... %d != %zu\n", m_nbActivePrimitives, m_primitives.size());
}

void SceneHost::sceneBounds(float* o) const
{
    o[0] = m_minPos.x; o[1] = m_minPos.y; o[2] = m_minPos.z; o[3] = m_maxPos.x; o[4] = m_maxPos.y; o[5] = m_maxPos.z;
}

// ---------------------------------------------------------------------------------------------------
// frame protocol (cuda/CudaKernel.cpp:116-145, 174-313)
// ---------------------------------------------------------------------------------------------------
void SceneHost::setLimits(int w, int h) { m_maxWidth = w; m_maxHeight = h; }
void SceneHost::setPartition(int rank, int world) { m_rank = rank; m_world = world; }
void SceneHost::setDevice(int device) { m_device = device; }

void SceneHost::setRandoms(const float* randoms, size_t n, int timestamp)
{
    m_hRandoms.assign((size_t)m_maxWidth * m_maxHeight, 0.f);
    memcpy(m_hRandoms.data(), randoms, std::min(n, m_hRandoms.size()) * sizeof(float));
    m_sceneInfo.timestamp = timestamp;
    m_randomsTransfered = false;
}

void SceneHost::initBuffers() // CudaKernel.cpp:116-145 + GPUKernel.cpp:299-360
{
    const size_t px = (size_t)m_maxWidth * m_maxHeight;
    dropSharedFrame();
    unpinBuffers(); // a second initBuffers may move them
    m_bitmap.assign(px * B200_COLOR_DEPTH, 0);
    b200_PrimitiveXYIdBuffer zero = {0, 0, 0, 0};
    m_primitivesXYIds.assign(px, zero);
    m_bitmapPtr = m_bitmap.data(); m_idsPtr = m_primitivesXYIds.data();
    if (m_hRandoms.size() != px)
    {
        // GPUKernel::render_begin fills the table with 0.000005f * (rand() % 2000 - 1000), rand() seeded from
        // time(0) (GPUKernel.cpp:2719-2727).  Same distribution, fixed seed: frames are reproducible.
        std::mt19937 gen(20261017u);
        m_hRandoms.resize(px);
        for (size_t i = 0; i < px; ++i) m_hRandoms[i] = 0.000005f * ((int)(gen() % 2000u) - 1000);
    }
    b200_int2 occ = {1, 1};
    if (m_device >= 0) b200_set_device(m_device);
    b200_set_limits(m_maxWidth, m_maxHeight);
    b200_set_partition(m_rank, m_world);
    b200_initialize_scene(occ, m_sceneInfo, (int)m_maxPrimitives, NB_MAX_LAMPS, B200_NB_MAX_MATERIALS);
    b200_reshape_scene(occ, m_sceneInfo);
    m_deviceInitialised = true;
    // this object owns the frame and id buffers for its lifetime: pinned in place, read-backs are direct DMAs
    if (b200_register_host(m_bitmap.data(), m_bitmap.size()) == 0) m_pinnedBitmap = m_bitmap.data();
    if (b200_register_host(m_primitivesXYIds.data(), m_primitivesXYIds.size() * sizeof(b200_PrimitiveXYIdBuffer)) == 0) m_pinnedIds = m_primitivesXYIds.data();
    m_primitivesTransfered = m_materialsTransfered = m_texturesTransfered = m_randomsTransfered = false;
}

void SceneHost::render_begin(const float) // CudaKernel.cpp:174-302
{
    if (!m_deviceInitialised) initBuffers();
    b200_int2 occ = {1, 1};
    if (m_refresh)
    {
        const int nbBoxes = m_nbActiveBoxes, nbPrimitives = m_nbActivePrimitives, nbLamps = m_nbActiveLamps;
        const int nbMaterials = m_nbActiveMaterials + 1;
        if (!m_primitivesTransfered)
        {
            b200_h2d_scene(occ, m_hBoundingBoxes.data(), nbBoxes, m_hPrimitives.data(), nbPrimitives, m_hLamps.data(), nbLamps);
            b200_h2d_lightInformation(occ, m_lightInformation.data(), (int)m_lightInformation.size());
            m_primitivesTransfered = true;
        }
        if (!m_randomsTransfered)
        {
            b200_h2d_randoms(occ, m_hRandoms.data());
            m_randomsTransfered = true;
        }
        if (!m_materialsTransfered)
        {
            realignTexturesAndMaterials();
            b200_h2d_materials(occ, m_hMaterials.data(), nbMaterials);
            m_materialsTransfered = true;
        }
        if (!m_texturesTransfered)
        {
            std::vector<b200_TextureInfo> infos(m_textures.size());
            for (size_t i = 0; i < m_textures.size(); ++i)
            {
                memset(&infos[i], 0, sizeof(b200_TextureInfo));
                infos[i].buffer = m_textures[i].texels.empty() ? nullptr : m_textures[i].texels.data();
                infos[i].offset = m_textures[i].offset;
                infos[i].size = m_textures[i].size;
            }
            b200_h2d_textures(occ, (int)infos.size(), infos.data());
            m_texturesTransfered = true;
        }
        b200_int4 objects = {nbBoxes, nbPrimitives, nbLamps, m_lightInformationSize};
        b200_SceneInfo sceneInfo = m_sceneInfo;
        if (m_sceneInfo.draftMode && m_sceneInfo.pathTracingIteration == 0) sceneInfo.graphicsLevel = B200_GL_NO_SHADING;
        if (m_sceneInfo.draftMode && m_sceneInfo.pathTracingIteration == m_sceneInfo.maxPathTracingIterations)
            sceneInfo.cameraType = B200_CT_ANTIALIASED;
        b200_int4 blockSize = {8, 4, 1, 0};
        b200_render(occ, blockSize, sceneInfo, objects, m_postProcessingInfo, m_viewPos, m_viewDir, m_angles);
    }
    m_refresh = (m_sceneInfo.pathTracingIteration < m_sceneInfo.maxPathTracingIterations);
}

void SceneHost::render_end() // CudaKernel.cpp:304-313 (the GL blit that follows there is the viewer's business)
{
    // The reference copies the id buffer back with every frame (33 MB at 1080p beside 6 MB of pixels) although only
    // getPrimitiveAt reads it.  Here it stays on the device until somebody asks (setLazyIds(false) = the reference's protocol).
    b200_int2 occ = {1, 1};
    // (a shared frame receives both from every GPU's kernels: nothing to copy, nothing left on the device)
    const bool lazy = m_lazyIds && !m_shared;
    b200_d2h_bitmap(occ, m_sceneInfo, m_bitmapPtr, lazy ? nullptr : m_idsPtr);
    m_idsOnDevice = lazy;
}

b200_PrimitiveXYIdBuffer* SceneHost::getPrimitiveIds()
{
    if (m_idsOnDevice && m_deviceInitialised)
    {
        b200_int2 occ = {1, 1};
        b200_d2h_bitmap(occ, m_sceneInfo, nullptr, m_idsPtr);
        m_idsOnDevice = false;
    }
    return m_idsPtr;
}

unsigned int SceneHost::getPrimitiveAt(int x, int y) // GPUKernel.cpp:729-739
{
    unsigned int returnValue = (unsigned int)-1;
    const unsigned int index = y * m_sceneInfo.size.x + x;
    if (index < static_cast<unsigned int>(m_sceneInfo.size.x * m_sceneInfo.size.y))
    {
        if (m_idsOnDevice && m_deviceInitialised && x >= 0 && x < m_sceneInfo.size.x)
            b200_d2h_primitive_id(m_sceneInfo, x, y, &m_idsPtr[index]);
        returnValue = m_idsPtr[index].x;
    }
    return returnValue;
}
} // namespace solr_b200


// ---------------------------------------------------------------------------------------------------
// animation step (GPUKernel.cpp:1378-1513, :1574-1721)
// ---------------------------------------------------------------------------------------------------
namespace solr_b200
{
namespace
{
void rotateVector(b200_float3& v, const b200_float3& rotationCenter, const b200_float3& cosAngles, const b200_float3& sinAngles) // :1602-1632
{
    b200_float3 vector = v3(v.x - rotationCenter.x, v.y - rotationCenter.y, v.z - rotationCenter.z);
    b200_float3 result = vector;
    /* X axis */
    result.y = vector.y * cosAngles.x - vector.z * sinAngles.x;
    result.z = vector.y * sinAngles.x + vector.z * cosAngles.x;
    vector = result;
    /* Y axis */
    result.z = vector.z * cosAngles.y - vector.x * sinAngles.y;
    result.x = vector.z * sinAngles.y + vector.x * cosAngles.y;
    vector = result;
    /* Z axis */
    result.x = vector.x * cosAngles.z - vector.y * sinAngles.z;
    result.y = vector.x * sinAngles.z + vector.y * cosAngles.z;
    v = v3(result.x + rotationCenter.x, result.y + rotationCenter.y, result.z + rotationCenter.z);
}
void rotatePrimitive(HostPrimitive& primitive, const b200_float3& rotationCenter, const b200_float3& cosAngles,
                     const b200_float3& sinAngles) // :1641-1674
{
    rotateVector(primitive.p0, rotationCenter, cosAngles, sinAngles);
    if (primitive.type == B200_PT_CYLINDER || primitive.type == B200_PT_TRIANGLE)
    {
        rotateVector(primitive.p1, rotationCenter, cosAngles, sinAngles);
        rotateVector(primitive.p2, rotationCenter, cosAngles, sinAngles);
        const b200_float3 zeroCenter = v3(0.f, 0.f, 0.f);
        rotateVector(primitive.n0, zeroCenter, cosAngles, sinAngles);
        rotateVector(primitive.n1, zeroCenter, cosAngles, sinAngles);
        rotateVector(primitive.n2, zeroCenter, cosAngles, sinAngles);
        if (primitive.type == B200_PT_CYLINDER)
        {
            b200_float3 axis = v3(primitive.p1.x - primitive.p0.x, primitive.p1.y - primitive.p0.y, primitive.p1.z - primitive.p0.z);
            const float len = sqrtf(axis.x * axis.x + axis.y * axis.y + axis.z * axis.z);
            if (len != 0) { axis.x /= len; axis.y /= len; axis.z /= len; }
            primitive.n1 = axis;
        }
    }
}
} // namespace

// every level-0 box is reset and re-fitted to its primitives (an empty one keeps the reset bounds and its old centre: the
// reference calls updateBoundingBox inside the loop over the box's primitives), then every box of every upper level to its children
void SceneHost::refreshBoxesAfterMove()
{
    for (int b = 1; b < 64; ++b)
        for (auto& box : m_boundingBoxes[b]) updateOutterBoundingBox(box.second, b - 1);
}

// The reference arrays come back from the device (where the last animation steps were applied), the primitives go back into the
// container by their ids, and the boxes are re-fitted the way a host-side step would have left them.
void SceneHost::syncFromDevice()
{
    if (!m_hostStale) return;
    m_hostStale = false;
    if (b200_d2h_scene(m_hBoundingBoxes.data(), m_hPrimitives.data()) != 0) return;
    materialiseBoxes();
    for (const b200_Primitive& p : m_hPrimitives)
    {
        auto it = m_primitives.find((unsigned)p.index);
        if (it == m_primitives.end()) continue;
        HostPrimitive& hp = it->second;
        hp.p0 = p.p0; hp.p1 = p.p1; hp.p2 = p.p2; hp.n0 = p.n0; hp.n1 = p.n1; hp.n2 = p.n2; hp.size = p.size;
    }
    for (auto& entry : m_boundingBoxes[0])
    {
        HostBox& box = entry.second;
        resetBox(box, false);
        if (!box.primitives.empty()) updateBoundingBox(box);
    }
    refreshBoxesAfterMove();
}

void SceneHost::rotatePrimitives(const b200_float3& rotationCenter, const b200_float3& angles) // :1378-1460
{
    if (m_deviceAnimation && m_deviceInitialised && m_primitivesTransfered && b200_rotate_primitives(rotationCenter, angles) == 0)
    {
        m_hostStale = true;
        return;
    }
    syncFromDevice();
    materialiseBoxes();
    m_primitivesTransfered = false;
    const b200_float3 cosAngles = v3(cos(angles.x), cos(angles.y), cos(angles.z));
    const b200_float3 sinAngles = v3(sin(angles.x), sin(angles.y), sin(angles.z));
    for (auto& entry : m_boundingBoxes[0])
    {
        HostBox& box = entry.second;
        resetBox(box, false);
        for (long id : box.primitives)
        {
            HostPrimitive& primitive = m_primitives[(unsigned)id];
            if (primitive.type != B200_PT_CAMERA) rotatePrimitive(primitive, rotationCenter, cosAngles, sinAngles);
        }
        if (!box.primitives.empty()) updateBoundingBox(box);
    }
    refreshBoxesAfterMove();
}

void SceneHost::translatePrimitives(const b200_float3& translation) // :1462-1513
{
    if (m_deviceAnimation && m_deviceInitialised && m_primitivesTransfered && b200_translate_primitives(translation) == 0)
    {
        m_hostStale = true;
        return;
    }
    syncFromDevice();
    materialiseBoxes();
    m_primitivesTransfered = false;
    for (auto& entry : m_boundingBoxes[0])
    {
        HostBox& box = entry.second;
        resetBox(box, false);
        for (long id : box.primitives)
        {
            HostPrimitive& primitive = m_primitives[(unsigned)id];
            if (primitive.type != B200_PT_CAMERA)
            {
                primitive.p0.x += translation.x; primitive.p0.y += translation.y; primitive.p0.z += translation.z;
                primitive.p1.x += translation.x; primitive.p1.y += translation.y; primitive.p1.z += translation.z;
                primitive.p2.x += translation.x; primitive.p2.y += translation.y; primitive.p2.z += translation.z;
            }
        }
        if (!box.primitives.empty()) updateBoundingBox(box);
    }
    refreshBoxesAfterMove();
}

void SceneHost::scalePrimitives(const float scale) // :1574-1600 (the reference ignores its from / to arguments)
{
    syncFromDevice();
    m_primitivesTransfered = false;
    for (auto& entry : m_primitives)
    {
        HostPrimitive& primitive = entry.second;
        primitive.p0.x *= scale; primitive.p0.y *= scale; primitive.p0.z *= scale;
        primitive.p1.x *= scale; primitive.p1.y *= scale; primitive.p1.z *= scale;
        primitive.p2.x *= scale; primitive.p2.y *= scale; primitive.p2.z *= scale;
        primitive.size.x *= scale; primitive.size.y *= scale; primitive.size.z *= scale;
    }
}
} // namespace solr_b200

// ---------------------------------------------------------------------------------------------------
// flat C API (ctypes)
// ---------------------------------------------------------------------------------------------------
using solr_b200::SceneHost;

struct b200h_Scene
{
    const void* boxes; int nbBoxes;
    const void* primitives; int nbPrimitives;
    const void* materials; int nbMaterials;
    const void* lightInformation; int lightInformationSize;
    const int* lamps; int nbLamps;
    float bounds[6];
    int treeDepth;
};

extern "C" {
void* b200h_create(const b200_SceneInfo* si) { return new SceneHost(*si); }
void b200h_destroy(void* h) { delete static_cast<SceneHost*>(h); }
void b200h_set_scene_info(void* h, const b200_SceneInfo* si) { static_cast<SceneHost*>(h)->setSceneInfo(*si); }
void b200h_set_post_processing_info(void* h, const b200_PostProcessingInfo* pp) { static_cast<SceneHost*>(h)->setPostProcessingInfo(*pp); }
void b200h_set_camera(void* h, const float* eye, const float* dir, const float* angles)
{
    b200_float3 e = {eye[0], eye[1], eye[2]}, d = {dir[0], dir[1], dir[2]};
    b200_float4 a;
    a.x = angles[0]; a.y = angles[1]; a.z = angles[2]; a.w = angles[3];
    static_cast<SceneHost*>(h)->setCamera(e, d, a);
}
int b200h_add_primitive(void* h, int type) { return static_cast<SceneHost*>(h)->addPrimitive(type); }
void b200h_set_primitive(void* h, int index, const float* v, int materialId)
{
    static_cast<SceneHost*>(h)->setPrimitive(index, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], materialId);
}
void b200h_add_primitives(void* h, int n, const int* types, const float* v, const int* materialIds)
{
    SceneHost* s = static_cast<SceneHost*>(h);
    for (int p = 0; p < n; ++p)
    {
        const int id = s->addPrimitive(types[p]);
        const float* V = v + 12 * p;
        s->setPrimitive(id, V[0], V[1], V[2], V[3], V[4], V[5], V[6], V[7], V[8], V[9], V[10], V[11], materialIds[p]);
    }
}
void b200h_set_normals_bulk(void* h, int first, int n, const float* normals)
{
    SceneHost* s = static_cast<SceneHost*>(h);
    for (int p = 0; p < n; ++p)
    {
        const float* N = normals + 9 * p;
        b200_float3 n0 = {N[0], N[1], N[2]}, n1 = {N[3], N[4], N[5]}, n2 = {N[6], N[7], N[8]};
        s->setPrimitiveNormals(first + p, n0, n1, n2);
    }
}
void b200h_set_texcoords(void* h, int index, const float* t)
{
    b200_float2 a = {t[0], t[1]}, b = {t[2], t[3]}, c = {t[4], t[5]};
    static_cast<SceneHost*>(h)->setPrimitiveTextureCoordinates(index, a, b, c);
}
int b200h_add_material(void* h) { return static_cast<SceneHost*>(h)->addMaterial(); }
void b200h_add_materials(void* h, int n, const float* f, const int* i)
{
    SceneHost* s = static_cast<SceneHost*>(h);
    for (int m = 0; m < n; ++m)
    {
        const int id = s->addMaterial();
        const float* F = f + 14 * m;
        const int* I = i + 11 * m;
        s->setMaterial(id, F[0], F[1], F[2], F[3], F[4], F[5], I[0] != 0, I[1] != 0, I[2], F[6], F[7], I[3], I[4], I[5], I[6],
                       I[7], I[8], I[9], F[8], F[9], F[10], F[11], F[12], F[13], I[10] != 0);
    }
}
void b200h_set_material_raw(void* h, int index, const b200_Material* m) { static_cast<SceneHost*>(h)->setMaterial(index, *m); }
void b200h_set_texture(void* h, int index, const unsigned char* texels, int w, int hh, int d) { static_cast<SceneHost*>(h)->setTexture(index, texels, w, hh, d); }
int b200h_compact_boxes(void* h, int reconstruct) { return static_cast<SceneHost*>(h)->compactBoxes(reconstruct != 0); }
void b200h_get_scene(void* h, b200h_Scene* out)
{
    SceneHost* s = static_cast<SceneHost*>(h);
    s->syncFromDevice();
    out->boxes = s->boxes(); out->nbBoxes = s->nbActiveBoxes();
    out->primitives = s->primitives(); out->nbPrimitives = s->nbActivePrimitives();
    out->materials = s->materials(); out->nbMaterials = s->nbMaterials();
    out->lightInformation = s->lightInformation(); out->lightInformationSize = s->lightInformationSize();
    out->lamps = s->lamps(); out->nbLamps = s->nbActiveLamps();
    s->sceneBounds(out->bounds);
    out->treeDepth = s->treeDepth();
}
void b200h_set_randoms(void* h, const float* r, long n, int timestamp) { static_cast<SceneHost*>(h)->setRandoms(r, (size_t)n, timestamp); }
void b200h_set_capacity(void* h, long maxBoxes, long maxPrimitives) { static_cast<SceneHost*>(h)->setCapacity((size_t)maxBoxes, (size_t)maxPrimitives); }
void b200h_set_limits(void* h, int w, int hh) { static_cast<SceneHost*>(h)->setLimits(w, hh); }
void b200h_rotate_primitives(void* h, const float* c, const float* a)
{
    b200_float3 center = {c[0], c[1], c[2]}, angles = {a[0], a[1], a[2]};
    static_cast<SceneHost*>(h)->rotatePrimitives(center, angles);
}
void b200h_translate_primitives(void* h, const float* t)
{
    b200_float3 translation = {t[0], t[1], t[2]};
    static_cast<SceneHost*>(h)->translatePrimitives(translation);
}
void b200h_scale_primitives(void* h, float scale) { static_cast<SceneHost*>(h)->scalePrimitives(scale); }
void b200h_set_device_animation(void* h, int on) { static_cast<SceneHost*>(h)->setDeviceAnimation(on != 0); }
void b200h_sync_from_device(void* h) { static_cast<SceneHost*>(h)->syncFromDevice(); }
void b200h_set_flat_build(void* h, int mode) { static_cast<SceneHost*>(h)->setFlatBuild(mode); }
void b200h_set_lazy_ids(void* h, int lazy) { static_cast<SceneHost*>(h)->setLazyIds(lazy != 0); }
void b200h_set_partition(void* h, int rank, int world) { static_cast<SceneHost*>(h)->setPartition(rank, world); }
void b200h_set_device(void* h, int device) { static_cast<SceneHost*>(h)->setDevice(device); }
void b200h_init_buffers(void* h) { static_cast<SceneHost*>(h)->initBuffers(); }
int b200h_share_frame(void* h, const char* name, int create) { return static_cast<SceneHost*>(h)->shareFrame(name, create != 0); }
void b200h_render_begin(void* h, float timer) { static_cast<SceneHost*>(h)->render_begin(timer); }
void b200h_render_end(void* h) { static_cast<SceneHost*>(h)->render_end(); }
unsigned char* b200h_get_bitmap(void* h) { return static_cast<SceneHost*>(h)->getBitmap(); }
void* b200h_get_primitive_ids(void* h) { return static_cast<SceneHost*>(h)->getPrimitiveIds(); }
unsigned int b200h_get_primitive_at(void* h, int x, int y) { return static_cast<SceneHost*>(h)->getPrimitiveAt(x, y); }
}
