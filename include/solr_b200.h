/* solr_b200.h — C ABI of the B200-native engine for Sol-R's ray-propagation hot path.
 *
 * The ten seam functions below are what the reference's engine host class binds
 * (/root/reference/solr/engines/cuda/CudaRayTracer.h:25-67, called from CudaKernel.cpp:116-145,174-302,
 * 304-313): same argument order, same by-value structs (include/solr_b200_types.h has the wire format),
 * same call protocol (init -> (h2d* -> render -> d2h)* -> finalize).  Names carry a b200_ prefix so the
 * library can sit in one process next to the reference's own engine.  All functions return void like the
 * reference's; failures are logged to stderr and latched — query with b200_last_error().  Unlike the
 * reference (CudaRayTracer.cu:1530) finalize does NOT cudaDeviceReset(): the process may share the device
 * with NCCL / PyTorch.
 *
 * One process drives one GPU (occupancyParameters.x must be 1; .y is ignored — persistent CTAs replace the
 * reference's stream split).  Multi-GPU = one process per GPU, the frame split by b200_set_partition().
 */
#ifndef SOLR_B200_H
#define SOLR_B200_H

#include "solr_b200_types.h"

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- the seam (1:1 with CudaRayTracer.h) ------------------------------------------------------------ */

/* CudaRayTracer.h:25 / CudaRayTracer.cu:1408-1491.  Creates the stream and scene-independent state. */
void b200_initialize_scene(b200_int2 occupancyParameters, b200_SceneInfo sceneInfo, int nbPrimitives, int nbLamps,
                           int nbMaterials);
/* CudaRayTracer.h:33 / .cu:1499-1532.  Frees every device buffer; no device reset. */
void b200_finalize_scene(b200_int2 occupancyParameters);
/* CudaRayTracer.h:40 / .cu:1360-1400.  (Re)allocates the per-pixel buffers for the frame-size limits
 * (b200_set_limits; default = the reference's 1920x1080) and zeroes them. */
void b200_reshape_scene(b200_int2 occupancyParameters, b200_SceneInfo sceneInfo);
/* CudaRayTracer.h:42 / .cu:1540-1555.  Takes the flattened AoS arrays produced by compactBoxes and
 * re-lays them out for the device (DESIGN.md "Data layout"). */
void b200_h2d_scene(b200_int2 occupancyParameters, const b200_BoundingBox* boundingBoxes, int nbActiveBoxes,
                    const b200_Primitive* primitives, int nbPrimitives, const int* lamps, int nbLamps);
/* CudaRayTracer.h:45 / .cu:1557-1565 */
void b200_h2d_materials(b200_int2 occupancyParameters, const b200_Material* materials, int nbActiveMaterials);
/* CudaRayTracer.h:47 / .cu:1567-1576.  Copies exactly maxWidth*maxHeight floats (the limits). */
void b200_h2d_randoms(b200_int2 occupancyParameters, const float* randoms);
/* CudaRayTracer.h:49 / .cu:1578-1613 */
void b200_h2d_textures(b200_int2 occupancyParameters, int activeTextures, const b200_TextureInfo* textureInfos);
/* CudaRayTracer.h:51 / .cu:1615-1625 */
void b200_h2d_lightInformation(b200_int2 occupancyParameters, const b200_LightInformation* lightInformation,
                               int lightInformationSize);
/* CudaRayTracer.h:54 / .cu:1647-1672.  Synchronises the render stream, then copies RGB8 (W*H*3) and ids
 * (W*H int4) to the caller's host buffers. Returns when the copies are complete. */
void b200_d2h_bitmap(b200_int2 occupancyParameters, b200_SceneInfo sceneInfo, b200_BitmapBuffer* bitmap,
                     b200_PrimitiveXYIdBuffer* primitivesXYIds);
/* CudaRayTracer.h:57 (cudaRender) / .cu:1680-1908.  objects = {boxes, primitives, lamps, lightInformation}.
 * blockSize is accepted for signature parity and ignored (persistent CTAs). Asynchronous on the stream. */
void b200_render(b200_int2 occupancyParameters, b200_int4 blockSize, b200_SceneInfo sceneInfo, b200_int4 objects,
                 b200_PostProcessingInfo postProcessingInfo, b200_float3 origin, b200_float3 direction,
                 b200_float4 angles);

/* ---- extensions (not in the reference seam) ------------------------------------------------------------ */

/* 0 = ok; otherwise the first latched error (cudaError_t value or negative engine code); msg may be NULL. */
int b200_last_error(char* msg, int msgCapacity);
void b200_clear_error(void);
/* CUDA device ordinal for this process (default: current device). Call before initialize_scene. */
void b200_set_device(int device);
/* Run on a caller-owned CUDA stream (cudaStream_t as void*), e.g. PyTorch's current stream; NULL = own stream. */
void b200_set_stream(void* cudaStream);
/* Frame-size limits (the reference hard-codes 1920x1080, Consts.h:39-41; 4K configs need more). Also the
 * length of the random table and the modulus of its index wraps. Call before reshape_scene. */
void b200_set_limits(int maxBitmapWidth, int maxBitmapHeight);
/* Engine options.  key 1 = device box layout: 0 auto (ordered BVH over the reference's leaves when provably
 * equivalent, else literal), 1 literal (the reference's hierarchy, single-child chains collapsed), 2 ordered BVH.
 * key 2 = which walks run warp-synchronously as packets (bit0 primary rays, bit1 secondary rays, bit2 shadow rays
 * of primary hits, bit3 other shadow rays); packets are only used with the ordered BVH.
 * key 3 = per-lane walks over the 4-wide form of the ordered BVH (1, default) or over the binary list (0).
 * key 4 = order-independent walks over the unordered SAH BVH where they are exact (1, default; a bounce ray whose candidates form a
 *         chain up to the edge of its gather window, or overflow its list, takes the ordered walk) or ordered walks only (0).
 * key 5 = order-independent walks also look up hits BEHIND the ray origin, which the reference's cylinder/cone test registers
 *         (GeometryIntersections.cuh:316-325) (1, default); 0 drops them (not reference-exact; for measurement).
 * key 6 = staged rendering: one launch per bounce pass over a compacted queue of the paths still alive (1, default, used for
 *         the one-ray-tree-per-pixel cameras and the anaglyph camera), the single persistent kernel for every camera (0), or
 *         fused stages (2: every pass of the frame in ONE persistent launch, a warp that runs out of work of one pass taking
 *         work of another; exact, measured 5 % slower on one GPU — engine.cu k_stage_fused, profiles/r02_history.md).
 * key 8 = bounce pass p whose queue holds at most p times this percentage of the GPU's resident lanes carries its paths to the
 *         end of their ray trees in registers instead of queueing them for one more launch per pass (default 300; 0 = never).
 * key 9 = order in which a GPU's own tiles are handed to its warps: 0 row-major (default), 1 along a Z-order curve.
 * key 10 = where b200_h2d_scene builds the trees of the order-independent walks: 0 on host threads (binned SAH, default: the better
 *         tree for a scene that is uploaded once), 1 on the GPU (linear BVH, sol-r_b200/csrc/treebuild.cuh: milliseconds instead of
 *         tenths of a second per upload, for scenes that are re-uploaded every animation step — GPUKernel::rotatePrimitives +
 *         compactBoxes(false), MoleculeScene.cpp:75-81).  Same frames either way.
 * key 11 = what b200_rotate_primitives / b200_translate_primitives / b200_scale_primitives do with the main walk tree: 1 re-fit it in
 *         place (default: the tree keeps the shape its builder gave it), 0 rebuild it on the GPU.  Same frames either way.
 * key 12 = streamed output into registered host buffers (1, default; see b200_frames_streamed below), 0 = always copy.
 */
void b200_set_option(int key, int value);
/* Multi-GPU frame split: this process renders tiles t with t % worldSize == rank (interleaved 8x4-pixel
 * tiles). Non-owned pixels of the device bitmap/ids stay zero so frames merge by summation. */
void b200_set_partition(int rank, int worldSize);
/* The frame exchange fused into the ray kernels (replaces a reduce/gather of partial bitmaps after them; the reference's
 * dormant multi-GPU path copies every device's band back through the host, CudaRayTracer.cu:1647-1672).  The root process
 * exports a 64-byte inter-process handle of its device bitmap (after reshape_scene; a reshape invalidates it), every other
 * process opens it, and from then on the kernel that ends a path stores its RGB8 straight into the root's frame through
 * NVLink peer memory; the processes' own device bitmaps are no longer written.  The caller orders frames across processes
 * (one stream-ordered barrier when the kernels are done, one before the next frame starts; sol-r_b200/partition.py).
 * b200_peer_frame_open(NULL, 0) goes back to the local bitmap.  Both return 0 or the latched error code. */
int b200_peer_frame_export(void* handle64, int handleBytes);
int b200_peer_frame_open(const void* handle64, int handleBytes);
/* The reference's animation step on the device-resident scene: GPUKernel::rotatePrimitives / translatePrimitives / scalePrimitives
 * (GPUKernel.cpp:1378-1513, :1574-1600) followed by compactBoxes(false) and the re-upload (MoleculeScene.cpp:75-81 does this per frame).
 * The step keeps the hierarchy's shape, so the flattened arrays keep their order and only coordinates change: the primitives are
 * moved, the reference's boxes re-fitted and everything the engine derives from them rebuilt WITHOUT leaving the device
 * (sol-r_b200/csrc/animate.cuh), with the reference's arithmetic — b200_d2h_scene afterwards returns arrays byte-identical to the
 * host container's after the same step.  Needs a scene uploaded by b200_h2d_scene with the default layout; returns 0 or the latched
 * error (-13: this scene cannot be animated on the device).  angles: radians about x, y, z, as the reference takes them.
 * b200_d2h_scene copies the (animated) reference arrays back, e.g. to re-synchronise a host container; either pointer may be NULL. */
int b200_rotate_primitives(b200_float3 rotationCenter, b200_float3 angles);
int b200_translate_primitives(b200_float3 translation);
int b200_scale_primitives(float scale);
int b200_d2h_scene(b200_BoundingBox* boundingBoxes, b200_Primitive* primitives);
float b200_last_animation_ms(void);
/* Scene replication for the multi-GPU frame split (replaces the reference's per-device upload loop, CudaRayTracer.cu:1540-1613): the
 * root process uploads the scene with b200_h2d_scene, b200_scene_layout describes what it built (B200_SCENE_LAYOUT_ENTRIES
 * numbers), every other process passes those numbers to b200_scene_adopt_layout, which allocates its device arrays, and the
 * arrays listed by b200_scene_device_arrays (same order on every process; a NULL pointer is an empty array) are then filled from
 * the root's over NVLink — sol-r_b200/partition.py broadcast_scene does it with one NCCL broadcast per array.
 * b200_scene_adopt_finish completes the adopted scene (host copy of the primitives, packed material words).  Materials, lights,
 * textures and randoms go up per process as before.  All return a count or 0, or a negative / latched error code. */
#define B200_SCENE_LAYOUT_ENTRIES 16
int b200_scene_layout(long long* layout, int capacity);
int b200_scene_adopt_layout(const long long* layout, int entries);
int b200_scene_device_arrays(void** devicePointers, long long* bytes, int capacity);
int b200_scene_adopt_finish(void);
/* Device pointers of the per-pixel buffers for in-place collectives (NCCL) — valid until reshape/finalize. */
void b200_device_buffers(void** bitmap, void** primitivesXYIds, void** postProcessingBuffer);
/* One pixel's PrimitiveXYIdBuffer (16 bytes) from the device: what GPUKernel::getPrimitiveAt (GPUKernel.cpp:729-739) reads, the only
 * host-side consumer of the id buffer d2h_bitmap copies every frame (33 MB at 1080p, CudaKernel.cpp:307).  A host class that
 * passes primitivesXYIds = NULL to b200_d2h_bitmap and asks here on a pick moves 6 MB per frame instead of 39 MB. */
void b200_d2h_primitive_id(b200_SceneInfo sceneInfo, int x, int y, b200_PrimitiveXYIdBuffer* id);
/* Copies the float accumulation buffer (W*H PostProcessingBuffer) to the host — the state k_default packs
 * into RGB8; the reference keeps it device-only (CudaRayTracer.cu:38).  Used by parity tests. */
void b200_d2h_post(b200_SceneInfo sceneInfo, b200_PostProcessingBuffer* postProcessingBuffer);
/* Work done by b200_render calls since the last reset: rays = box-list walks (closest-hit + shadow),
 * pixels = pixels actually traced.  Synchronises the stream. */
void b200_get_counters(unsigned long long* rays, unsigned long long* pixels, int reset);
/* Device time of the last b200_render's kernels in ms (CUDA events on the render stream); synchronises. */
float b200_last_render_ms(void);
/* Number of kernels this library launched since initialize_scene. */
unsigned long long b200_kernel_launches(void);
/* Bytes b200_render copies host->device per frame (scene-info, camera and buffer pointers: the kernel's
 * constant-memory parameter block). */
int b200_frame_parameter_bytes(void);
/* Compacted scene statistics of the last h2d_scene: boxes kept after single-child chain collapse etc. */
void b200_scene_stats(int* nbBoxesIn, int* nbBoxesDevice, int* nbPrimitives, int* reserved);
/* The last b200_h2d_scene: host wall-clock milliseconds of the whole call, nodes of the main and of the point-query walk tree,
 * and whether they were built on the GPU (option key 10).  Any pointer may be NULL. */
void b200_scene_upload_stats(float* milliseconds, int* walkTreeNodes, int* pointQueryTreeNodes, int* builtOnGpu);
/* Host-only (no CUDA call): the box re-layout h2d_scene applies — single-child chains collapsed, skip counts
 * recomputed, 8 floats per box: (min.xyz, w0) (max.xyz, w1), w as int bits.  Returns the number of boxes kept;
 * writes them if capacityBoxes suffices.  Lets tests check the collapse without a GPU. */
int b200_debug_relayout_boxes(const b200_BoundingBox* boxes, int nbBoxes, float* outPacked, int capacityBoxes);
/* Host-only: the unordered SAH BVH built over the same leaves (binary depth-first list, same 8-float format). */
int b200_debug_build_unordered(const b200_BoundingBox* boxes, int nbBoxes, float* outPacked, int capacityBoxes);
/* Host-only: the trees h2d_scene builds for the order-independent walks: 128-byte records of four child boxes
 * (rows lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4], then refs[4] as int bits: >= 0 node, < 0 leaf = ~primitive with bit 30 set
 * in the point-query tree, INT_MIN empty); nbMain nodes of the tree over primitives, then nbExt nodes of the point-query tree.
 * Returns the number of float4 (writes them if capacityFloat4 suffices); primLeafOut[nbPrimitives] = reference leaf per primitive. */
int b200_debug_build_walk_trees(const b200_BoundingBox* boxes, int nbBoxes, const b200_Primitive* primitives, int nbPrimitives,
                                float* outNodes, int capacityFloat4, int* primLeafOut, int* nbMainOut, int* nbExtOut);
/* Host-only debug read of the raw work counters (8 values; [0] rays, [1] pixels, others only in instrumented builds). */
void b200_debug_counters(unsigned long long* out8);
/* Block until everything queued on the render stream is done. */
void b200_synchronize(void);
/* Measurement helper (bench.py): FP32 FMA microbenchmark on the current device — achieved TFLOP/s of dependent FFMA chains on
 * every resident lane, and the SM clock sustained under that load (MHz): the measured roofline denominator BASELINE.md 2 asks
 * for beside the nominal 148 SM x 128 lanes x 2 x f.  Returns 0 or the latched error code. */
int b200_measure_fp32_peak(float* tflops, float* smMhz);

/* Pins a host buffer of the caller in place (cudaHostRegister) so that b200_d2h_bitmap into it is a direct DMA; the caller
 * unregisters it before freeing it.  The engine never pins memory it does not own on its own initiative.  A host class registers
 * the frame and id buffers it allocates once (GPUKernel.cpp:344-360).  0, or a negative code (-12: could not pin; copies into the
 * buffer still work, staged). */
int b200_register_host(void* buffer, size_t bytes);
int b200_unregister_host(void* buffer);
/* Streamed output (option key 12, default 1).  The reference's render_end reads the frame and the id buffer back after every frame
 * (CudaKernel.cpp:304-313: 19 bytes per pixel behind the kernels).  When b200_d2h_bitmap has just filled REGISTERED buffers, the
 * next frame's ray kernels write those buffers themselves — each 8 x 4 tile as its last path ends, over PCIe while the rest of the
 * frame is still being traced — and the b200_d2h_bitmap that follows with the same pointers only waits for the stream.  Same bytes as
 * the copy.  Frames it does not apply to (post-processing effect, single-kernel cameras, fused stages, BGR frames, sizes that are not
 * whole tiles, other buffers, a reader that takes the pixels without the ids) are copied as before; a frame rendered without an
 * intervening b200_d2h_bitmap is not streamed either (nobody is reading every frame).  b200_frames_streamed counts the frames whose outputs were written this way. */
unsigned long long b200_frames_streamed(void);
/* Names the host buffers of the frame outright: from now on EVERY frame this process renders goes there — tile by tile from the ray
 * kernels where they count tiles, else in one launch after the frame's other kernels — for the tiles this GPU owns
 * (b200_set_partition), whatever else is read or not.  This is how several processes fill ONE host frame: the buffers lie in memory
 * they share (POSIX shared memory mapped and registered by each: SceneHost::shareFrame, partition.SharedHostFrame), every GPU
 * writes its own tiles over its own PCIe link, and once every rank's stream is idle the frame AND the id buffer are whole in host
 * memory — no GPU-to-GPU exchange, no read-back on the root.  It takes precedence over b200_peer_frame_open (the root GPU's
 * frame as the destination).  Both buffers must be registered (b200_register_host) and hold si.size.x * si.size.y pixels; either
 * may be NULL; both NULL ends it.  b200_d2h_bitmap with these pointers only waits for the stream.  0 or a negative code. */
int b200_stream_target(b200_SceneInfo si, b200_BitmapBuffer* bitmap, b200_PrimitiveXYIdBuffer* ids);
/* Sample-split accumulation over GPUs (the second split north_star names: "sample accumulation optionally split by GPU", frame
 * "reduced with NCCL over NVLink").  Past NB_MAX_ITERATIONS the reference only adds a frame's sample to the accumulation buffer
 * (CudaRayTracer.cu:550-562) and divides by the sample count when it packs (k_default, :1066-1070), so the samples of a
 * progressive sequence can be rendered by different GPUs, each over the WHOLE frame (set_partition(0, 1) on every process), and
 * summed: _clear zeroes the colour sums of this process (keeps depth and ids), _export writes them as one float4 per pixel to a
 * device buffer of the caller (which the caller sum-reduces onto the root, sol-r_b200/partition.py SampleSplit), and
 * _import_and_pack stores the reduced sums and packs the RGB8 frame as iteration `iteration` would (divide by
 * iteration - NB_MAX_ITERATIONS + 1).  All three run on the render stream; each returns 0 or an error code. */
int b200_accumulation_clear(void);
int b200_accumulation_export(void* dstFloat4Device);
int b200_accumulation_import_and_pack(const void* srcFloat4Device, int iteration);

#ifdef __cplusplus
}
#endif
#endif /* SOLR_B200_H */
