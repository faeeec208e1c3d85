// scene_host.h — host side of the engine: the scene container and frame protocol the reference keeps in
// solr::GPUKernel / solr::CudaKernel, restated for the hot path so that scenes can be built and rendered
// without the reference tree (bench, tests, smoke) through the same calls a Sol-R application makes:
//
//   addPrimitive / setPrimitive / setPrimitiveNormals / setPrimitiveTextureCoordinates
//                                   /root/reference/solr/engines/GPUKernel.cpp:495-516, 528-684, 686-727
//   addMaterial / setMaterial       GPUKernel.cpp:1761-1909
//   setTexture                      GPUKernel.cpp:2017-2033, 2238-2340, 2691-2705
//   setSceneInfo / setCamera / setPostProcessingInfo       GPUKernel.cpp:481-493, 2071-2092
//   compactBoxes (grid hierarchy -> flattened BoundingBox[] / Primitive[] with skip counts)
//                                   GPUKernel.cpp:741-1039 (boxes), 1041-1083, 1085-1281 (flatten)
//   initBuffers / render_begin / render_end / getBitmap / getPrimitiveAt
//                                   cuda/CudaKernel.cpp:116-145, 174-302, 304-313; GPUKernel.cpp:729-739
//
// The flattened arrays are byte-identical to the reference's for the same setter calls (array order decides
// tie-breaking in the walk, so the builder is restated literally: same containers, same integer wraps;
// tests/test_scene_host.py compares against the reference built from source and against golden hashes).
// Rendering goes through the engine's C ABI only (include/solr_b200.h); nothing here computes pixels.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/solr_b200.h"

namespace solr_b200
{
struct HostPrimitive // GPUKernel.h:44-63 (CPUPrimitive), hot-path fields
{
    b200_float3 p0, p1, p2, n0, n1, n2, size;
    int type;
    int materialId;
    b200_float2 vt0, vt1, vt2;
};

struct HostBox // GPUKernel.h:65-71 (CPUBoundingBox)
{
    b200_float3 parameters[2];
    b200_float3 center;
    std::vector<long> primitives; // level 0: primitive ids; level k: keys of level k-1 boxes
    long indexForNextBox;
    // level k >= 1: the boxes those keys name, appended in step with `primitives` (std::map nodes do not move).  Where the two
    // lists have the same length the builder follows the pointers instead of looking the keys up again.
    std::vector<HostBox*> children;
};

class SceneHost
{
public:
    explicit SceneHost(const b200_SceneInfo& sceneInfo);
    ~SceneHost();

    // ---- scene (GPUKernel interface) ----
    void setSceneInfo(const b200_SceneInfo& si) { m_sceneInfo = si; }
    b200_SceneInfo& getSceneInfo() { return m_sceneInfo; }
    void setPostProcessingInfo(const b200_PostProcessingInfo& pp) { m_postProcessingInfo = pp; }
    void setCamera(const b200_float3& eye, const b200_float3& dir, const b200_float4& angles);
    int addPrimitive(int type);
    void setPrimitive(int index, float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2,
                      float w, float h, float d, int materialId);
    void setPrimitiveNormals(unsigned index, b200_float3 n0, b200_float3 n1, b200_float3 n2);
    void setPrimitiveTextureCoordinates(unsigned index, b200_float2 vt0, b200_float2 vt1, b200_float2 vt2);
    int addMaterial();
    void setMaterial(unsigned index, float r, float g, float b, float noise, float reflection, float refraction,
                     bool procedural, bool wireframe, int wireframeWidth, float transparency, float opacity,
                     int diffuseTextureId, int normalTextureId, int bumpTextureId, int specularTextureId,
                     int reflectionTextureId, int transparentTextureId, int ambientOcclusionTextureId, float specValue,
                     float specPower, float specCoef, float innerIllumination, float illuminationDiffusion,
                     float illuminationPropagation, bool fastTransparency);
    void setMaterial(unsigned index, const b200_Material& material);
    void setTexture(int index, const unsigned char* texels, int width, int height, int depth);
    int compactBoxes(bool reconstructBoxes);
    void resetBoxes(bool resetPrimitives);
    void setRandoms(const float* randoms, size_t n, int timestamp); // fixes what GPUKernel::render_begin draws from rand()

    // ---- frame (CudaKernel interface) ----
    void setLimits(int maxWidth, int maxHeight);
    // box / primitive capacity of the flattened arrays; default = the reference's NB_MAX_BOXES / NB_MAX_PRIMITIVES (2.5 M,
    // Consts.h:32-33), at which compactBoxes silently drops boxes for ~1 M-primitive scenes (SURVEY finding 4)
    void setCapacity(size_t maxBoxes, size_t maxPrimitives) { m_maxBoxes = maxBoxes; m_maxPrimitives = maxPrimitives; }
    // animation step (GPUKernel.cpp:1378-1513, :1574-1600): move the primitives, refresh the bounds of the existing boxes; the
    // caller flattens again with compactBoxes(false)
    void rotatePrimitives(const b200_float3& rotationCenter, const b200_float3& angles);
    void translatePrimitives(const b200_float3& translation);
    void scalePrimitives(float scale);
    void setPartition(int rank, int worldSize);
    void setDevice(int device);
    void initBuffers();
    void render_begin(float timer);
    void render_end();
    unsigned char* getBitmap() { return m_bitmapPtr; }
    // One host frame for several processes (one per GPU, each rendering its own tiles: setPartition): the frame and id buffers move
    // into POSIX shared memory `name` (created by the process that passes create = true, opened by the others afterwards), pinned
    // by each process and named to its engine as the destination of every frame (b200_stream_target): every GPU writes its own tiles
    // there over its own PCIe link, and once every rank's stream is idle getBitmap() / getPrimitiveIds() of any of them is the whole
    // frame.  Call after initBuffers(), with the same frame limits everywhere.  0, or a negative code.
    int shareFrame(const char* name, bool create);
    b200_PrimitiveXYIdBuffer* getPrimitiveIds(); // fetches the buffer from the device first when render_end left it there
    void setLazyIds(bool lazy) { m_lazyIds = lazy; }
    // Animation on the device (b200_rotate_primitives / b200_translate_primitives): once the scene is on the device, rotatePrimitives
    // and translatePrimitives move it THERE and this container's own copy goes stale until somebody needs it (syncFromDevice: any
    // setter, compactBoxes(true), the array accessors after a step) — the per-frame loop of MoleculeScene.cpp:75-81 never does.
    void setDeviceAnimation(bool on) { if (!on) syncFromDevice(); m_deviceAnimation = on; }
    void syncFromDevice();
    bool hostIsStale() const { return m_hostStale; }
    // 0: always the literal per-level maps; 1: flat build with the depth-first flatten; 2 (default): flat build with the flatten by
    // levels (tests compare the three)
    void setFlatBuild(int mode) { m_useFlatBuild = mode != 0; m_levelOrderFlatten = mode == 2; }
    unsigned int getPrimitiveAt(int x, int y);

    // ---- flattened arrays, as the engine seam receives them ----
    const b200_BoundingBox* boxes() const { return m_hBoundingBoxes.data(); }
    const b200_Primitive* primitives() const { return m_hPrimitives.data(); }
    const b200_Material* materials() const { return m_hMaterials.data(); }
    const b200_LightInformation* lightInformation() const { return m_lightInformation.data(); }
    const int* lamps() const { return m_hLamps.data(); }
    int nbActiveBoxes() const { return m_nbActiveBoxes; }
    int nbActivePrimitives() const { return m_nbActivePrimitives; }
    int nbActiveLamps() const { return m_nbActiveLamps; }
    int nbMaterials() const { return m_nbActiveMaterials + 1; }
    int lightInformationSize() const { return m_lightInformationSize; }
    int treeDepth() const { return (int)m_treeDepth; }
    void sceneBounds(float* out6) const;
    size_t nbPrimitivesAdded() const { return m_primitives.size(); }

private:
    bool updateBoundingBox(HostBox& box);
    void updateOutterBoundingBox(HostBox& outterBox, int depth);
    void resetBox(HostBox& box, bool resetPrimitives);
    void processBoxes(int boxSize);
    void processOutterBoxes(int boxSize, int depth);
    void recursiveDataStreamToGPU(int depth, std::vector<long>& elements, const std::vector<HostBox*>* linked);
    void streamDataToGPU();
    void emitPrimitive(long id);
    void refreshBoxesAfterMove();
    HostPrimitive& primitiveById(unsigned int id);
    // first compaction of a fresh container: the same hierarchy and the same flattened arrays from sorted flat arrays per level
    // (scene_host.cpp "flat build"); the per-level maps are only materialised if a later call needs them
    bool flatBuildApplies() const;
    bool flatBuild();
    void flatStream();
    bool flatStreamByLevels();
    void writePrimitive(size_t slot, long id);
    void flatRecurse(int depth, unsigned int box);
    void materialiseBoxes();
    void dropFlat();
    struct FlatHierarchy;
    FlatHierarchy* m_flat = nullptr;
    bool m_useFlatBuild = true;
    bool m_levelOrderFlatten = true;
    std::vector<HostPrimitive*> m_primitiveTable; // id -> record, valid during one compactBoxes()
    void realignTexturesAndMaterials();

    b200_SceneInfo m_sceneInfo;
    b200_PostProcessingInfo m_postProcessingInfo;
    b200_float3 m_viewPos, m_viewDir;
    b200_float4 m_angles;

    std::map<unsigned int, HostPrimitive> m_primitives;            // GPUKernel.h:74 (PrimitiveContainer)
    std::map<unsigned int, HostBox> m_boundingBoxes[64];           // per level; GPUKernel.h:73,414
    b200_float3 m_minPos, m_maxPos;
    unsigned int m_treeDepth;

    std::vector<b200_BoundingBox> m_hBoundingBoxes;
    std::vector<b200_Primitive> m_hPrimitives;
    std::vector<b200_Material> m_hMaterials;
    std::vector<b200_LightInformation> m_lightInformation;
    std::vector<int> m_hLamps;
    std::vector<float> m_hRandoms;
    struct Texture { std::vector<unsigned char> texels; int offset; b200_int3 size; };
    std::vector<Texture> m_textures; // NB_MAX_TEXTURES slots
    int m_nbActiveTextures;
    int m_nbActiveBoxes, m_nbActivePrimitives, m_nbActiveLamps, m_nbActiveMaterials, m_lightInformationSize;
    size_t m_maxBoxes, m_maxPrimitives;

    std::vector<unsigned char> m_bitmap;
    std::vector<b200_PrimitiveXYIdBuffer> m_primitivesXYIds;
    unsigned char* m_bitmapPtr = nullptr;           // the vectors' storage, or the shared frame's (shareFrame)
    b200_PrimitiveXYIdBuffer* m_idsPtr = nullptr;
    void* m_shared = nullptr; size_t m_sharedBytes = 0; std::string m_sharedName; bool m_sharedOwner = false;
    void dropSharedFrame();
    bool m_lazyIds = true;      // render_end copies the pixels only; ids are fetched when asked for
    bool m_deviceAnimation = false, m_hostStale = false;
    bool m_idsOnDevice = false; // the host copy is older than the last frame
    bool m_primitivesTransfered, m_materialsTransfered, m_texturesTransfered, m_randomsTransfered, m_refresh;
    bool m_deviceInitialised;
    void* m_pinnedBitmap = nullptr; // buffers registered with the engine (b200_register_host), unpinned before they are freed
    void* m_pinnedIds = nullptr;
    void unpinBuffers();
    int m_maxWidth, m_maxHeight, m_rank, m_world, m_device;
};
} // namespace solr_b200
