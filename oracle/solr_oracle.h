/* oracle/solr_oracle.h — TEST INFRASTRUCTURE (see solr_oracle.cpp).  C ABI of the CPU restatement. */
#ifndef SOLR_ORACLE_H
#define SOLR_ORACLE_H
#include <stdint.h>
#include "../include/solr_b200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* the arrays the engine seam receives (CudaRayTracer.h:25-58), already flattened by compactBoxes */
typedef struct {
    const b200_BoundingBox* boxes; int nbBoxes;
    const b200_Primitive* primitives; int nbPrimitives;
    const b200_Material* materials; int nbMaterials;
    const b200_LightInformation* lightInformation; int lightInformationSize; int nbLamps;
    const unsigned char* textures;
    const float* randoms; int randomTableSize; /* MAX_BITMAP_SIZE of the build restated (1920*1080 in the reference) */
} oracle_Scene;

/* work counted in reference traversal order; all uint64 */
typedef struct {
    uint64_t pixels, rays, primary_rays, shadow_rays, box_tests, sphere_tests, cylinder_tests, cone_tests,
        triangle_tests, plane_tests, ellipsoid_tests, accepted_hits, shade_calls;
} oracle_Counters;

int oracle_abi_version(void);

/* Renders rows rowBegin, rowBegin+rowStride, ... < rowEnd of one frame (ray generation, bounce loop,
 * accumulation into post/ids exactly as k_standardRenderer / k_anaglyphRenderer, then k_default into
 * bitmap).  post/ids carry the progressive state between frames, as the device buffers do. */
void oracle_render(const oracle_Scene* scene, const b200_SceneInfo* sceneInfo, const b200_PostProcessingInfo* postInfo,
                   const float* eye, const float* target, const float* angles, b200_PostProcessingBuffer* post,
                   b200_int4* ids, unsigned char* bitmap, int rowBegin, int rowEnd, int rowStride, int nThreads,
                   oracle_Counters* counters);

/* Post-processing effects of cudaRender's second pass (CudaRayTracer.cu:1081-1358) over a whole frame: reads post / ids,
 * rewrites bitmap.  No-op for ppe_none. */
void oracle_post_process(const oracle_Scene* scene, const b200_SceneInfo* sceneInfo, const b200_PostProcessingInfo* postInfo,
                         const b200_PostProcessingBuffer* post, const b200_int4* ids, unsigned char* bitmap);

double oracle_algorithmic_flops(const oracle_Counters* counters);

#ifdef __cplusplus
}
#endif
#endif
