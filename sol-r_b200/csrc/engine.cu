// engine.cu — the C ABI of include/solr_b200.h: device-buffer ownership, scene re-layout at upload, the
// persistent render kernel and the readback.  One process drives one GPU.
//
// Replaces, behind the same seam, /root/reference/solr/engines/cuda/CudaRayTracer.cu:1360-1908
// (reshape/initialize/finalize_scene, h2d_*, d2h_bitmap, cudaRender) and the kernels it launches
// (k_standardRenderer :437-563, k_anaglyphRenderer :840-926, k_default :1057-1073).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <chrono>
#include <algorithm>
#include <thread>

#include <cuda_runtime.h>

#include "../../include/solr_b200.h"
#include "shade.cuh"

#define TILE_W 8
#define TILE_H 4
#define CTA_THREADS 128
// 8 CTAs x 128 threads = 32 warps/SM (64 registers): the walks are bound by dependent-load latency, and the extra
// resident warps hide it better than the extra registers help (measured: 4 -> 11.9 ms, 6 -> 9.5 ms, 8 -> 8.4 ms, 12 -> 8.4 ms on config 2)
#ifndef MIN_CTAS_PER_SM
#define MIN_CTAS_PER_SM 6
#endif

// GeometryShaders.cuh:132-165 (makeColor) fused with k_default's averaging (CudaRayTracer.cu:1068-1072)
SB_DEV void packPixelAt(float4 color, unsigned char* bitmap, const int index, const int iteration)
{
    if (iteration > B200_NB_MAX_ITERATIONS)
        color /= (float)(iteration - B200_NB_MAX_ITERATIONS + 1);
    color.x = (color.x > 1.f) ? 1.f : color.x; color.y = (color.y > 1.f) ? 1.f : color.y; color.z = (color.z > 1.f) ? 1.f : color.z;
    color.x = (color.x < 0.f) ? 0.f : color.x; color.y = (color.y < 0.f) ? 0.f : color.y; color.z = (color.z < 0.f) ? 0.f : color.z;
    if (cSI.frameBufferType == B200_FT_BGR)
    {
        const int y = index / cSI.size.y, x = index % cSI.size.x;
        const int i = ((y + 1) * cSI.size.y - x - 1) * B200_COLOR_DEPTH;
        bitmap[i] = (unsigned char)(color.z * 255.f);
        bitmap[i + 1] = (unsigned char)(color.y * 255.f);
        bitmap[i + 2] = (unsigned char)(color.x * 255.f);
    }
    else
    {
        const int i = index * B200_COLOR_DEPTH;
        bitmap[i] = (unsigned char)(color.x * 255.f);
        bitmap[i + 1] = (unsigned char)(color.y * 255.f);
        bitmap[i + 2] = (unsigned char)(color.z * 255.f);
    }
}
SB_DEV void packPixel(const float4 color, unsigned char* bitmap, const int index) { packPixelAt(color, bitmap, index, cSI.pathTracingIteration); }

// Primary ray of a pixel for the standard / orthographic / antialiased cameras (CudaRayTracer.cu:462-522): eye and look-at
// target, the target shifted per pixel, both rotated; depth-of-field and rotated-grid jitter of the accumulation passes.
SB_DEV void primaryRay(const Rotation& rot, const int x, const int y, const int index, const float storedDepth, float3& o, float3& t)
{
    const int W = cSI.size.x, H = cSI.size.y;
    const int iter = cSI.pathTracingIteration;
    const int camera = cSI.cameraType;
    const float3 rotationCenter = (camera == B200_CT_VR) ? cP.eye : f3(0.f, 0.f, 0.f);
    const float ratio = (float)W / (float)H;
    const float stepx = ratio * cP.angles.w / (float)W, stepy = cP.angles.w / (float)H;
    o = cP.eye; t = cP.target;
    // NATURAL_DEPTHOFFIELD (Consts.h:54; CudaRayTracer.cu:470-479), precedence as written there
    if (cP.pp.type != B200_PPE_DEPTH_OF_FIELD && iter >= B200_NB_MAX_ITERATIONS)
    {
        const float a = (cP.pp.param1 / 20000.f);
        const int rindex = index + cSI.timestamp % (cS.randomTableSize - 2);
        const bool in = rindex + 1 < cS.randomTableSize + 4;
        o.x += (in ? rnd(rindex) : 0.f) * storedDepth * a;
        o.y += (in ? rnd(rindex + 1) : 0.f) * storedDepth * a;
    }
    if (camera == B200_CT_ORTHOGRAPHIC)
    {
        t.x = o.z * 0.001f * (x - (W / 2));
        t.y = -o.z * 0.001f * (y - (H / 2));
        o.x = t.x;
        o.y = t.y;
    }
    else
    {
        // fused, as in the reference's build (see mulAdd2 in shade.cuh)
        t.x = __fmaf_rn(-stepx, (float)(x - (W / 2)), t.x);
        t.y = __fmaf_rn(stepy, (float)(y - (H / 2)), t.y);
    }
    vectorRotation(o, rotationCenter, rot);
    vectorRotation(t, rotationCenter, rot);
    if (camera != B200_CT_ANTIALIASED && iter >= B200_NB_MAX_ITERATIONS)
    {
        // rotated-grid jitter of the accumulation passes (:515-522), applied after the rotation
        const int k = iter % 4;
        t.x += (k == 0) ? 3.f : (k == 1) ? 5.f : (k == 2) ? -3.f : -5.f;
        t.y += (k == 0) ? 5.f : (k == 1) ? -3.f : (k == 2) ? -5.f : 3.f;
    }
}

// k_anaglyphRenderer's rays (CudaRayTracer.cu:866-901): eyes at origin.x -/+ eyeSeparation, both origin and target rotated
SB_DEV void anaglyphRay(const Rotation& rot, const int x, const int y, const int eye, float3& o, float3& t)
{
    const int W = cSI.size.x, H = cSI.size.y;
    const float ratio = (float)W / (float)H;
    const float stepx = ratio * cP.angles.w / (float)W, stepy = cP.angles.w / (float)H;
    const float3 rotationCenter = f3(0.f, 0.f, 0.f);
    o = f3(eye == 0 ? cP.eye.x - cSI.eyeSeparation : cP.eye.x + cSI.eyeSeparation, cP.eye.y, cP.eye.z);
    t.x = cP.target.x - stepx * (float)(x - (W / 2));
    t.y = cP.target.y + stepy * (float)(y - (H / 2));
    t.z = cP.target.z;
    vectorRotation(o, rotationCenter, rot);
    vectorRotation(t, rotationCenter, rot);
}

// left eye -> luma (CudaRayTracer.cu:905): x 0.299 + y 0.587 + z 0.114 with the contraction nvcc gives that expression (first product
// fused into the sum, second rounded on its own, third fused), spelled out so that the two drivers that use it agree to the bit
SB_DEV float anaglyphLuma(const float4 c) { return __fmaf_rn(c.z, 0.114f, __fmaf_rn(c.x, 0.299f, __fmul_rn(c.y, 0.587f))); }

// What a pixel keeps of its ray tree(s) (:537-562, anaglyph :903-925), followed by k_default for that pixel.
SB_DEV void resolvePixel(const int index, float4 color, const float4 left, const int4 id, const float dof, float4 stored)
{
    const int iter = cSI.pathTracingIteration;
    const int camera = cSI.cameraType;
    float4 sinfo = *reinterpret_cast<float4*>(&cP.post[index].sceneInfo);
    if (iter == 0) stored.w = dof;
    if (camera == B200_CT_ANAGLYPH)
    {
        // left eye -> luma in red, right eye -> green/blue (:903-925); sceneInfo is not written
        const float r1 = anaglyphLuma(left);
        const float g2 = color.y, b2 = color.z;
        if (iter <= B200_NB_MAX_ITERATIONS) { stored.x = r1 + 0.f; stored.y = 0.f + g2; stored.z = 0.f + b2; }
        else { stored.x += r1 + 0.f; stored.y += 0.f + g2; stored.z += 0.f + b2; }
    }
    else if (camera == B200_CT_VR || camera == B200_CT_PANORAMIC)
    {
        // :1016-1042, :797-812: plain accumulation, sceneInfo is not written; only the stereo camera takes random illumination
        if (camera == B200_CT_VR && cSI.advancedIllumination == B200_AI_RANDOM)
        {
            const int rindex = (index + cSI.timestamp) % cS.randomTableSize;
            color += f4(cSI.backgroundColor.x, cSI.backgroundColor.y, cSI.backgroundColor.z, cSI.backgroundColor.w) * rnd(rindex) * 5.f;
        }
        if (iter <= B200_NB_MAX_ITERATIONS) { stored.x = color.x; stored.y = color.y; stored.z = color.z; }
        else { stored.x += color.x; stored.y += color.y; stored.z += color.z; }
    }
    else
    {
        if (cSI.advancedIllumination == B200_AI_RANDOM)
        {
            const int rindex = (index + cSI.timestamp) % cS.randomTableSize;
            color += f4(cSI.backgroundColor.x, cSI.backgroundColor.y, cSI.backgroundColor.z, cSI.backgroundColor.w) * rnd(rindex) * 5.f;
        }
        if (camera == B200_CT_ANTIALIASED) color /= 5.f;
        if (iter <= B200_NB_MAX_ITERATIONS)
        {
            stored.x = color.x; stored.y = color.y; stored.z = color.z;
            sinfo.x = color.x; sinfo.y = color.y; sinfo.z = color.z;
        }
        else
        {
            // accumulation passes (:550-562)
            sinfo.x = (id.z > 0) ? fmaxf(sinfo.x, color.x) : color.x;
            sinfo.y = (id.z > 0) ? fmaxf(sinfo.y, color.y) : color.y;
            sinfo.z = (id.z > 0) ? fmaxf(sinfo.z, color.z) : color.z;
            stored.x += sinfo.x; stored.y += sinfo.y; stored.z += sinfo.z;
        }
#if STREAM_HINTS >= 2
        stOnce(reinterpret_cast<float4*>(&cP.post[index].sceneInfo), sinfo);
#else
        *reinterpret_cast<float4*>(&cP.post[index].sceneInfo) = sinfo;
#endif
    }
#if STREAM_HINTS >= 2
    stOnce(reinterpret_cast<float4*>(&cP.post[index].colorInfo), stored);
    stOnce(reinterpret_cast<float4*>(cP.ids + index), make_float4(__int_as_float(id.x), __int_as_float(id.y), __int_as_float(id.z), __int_as_float(id.w)));
#else
    *reinterpret_cast<float4*>(&cP.post[index].colorInfo) = stored;
    cP.ids[index] = id;
#endif
    packPixel(stored, cP.bitmap, index);
}

// One channel of packPixelAt (c = 0 red, 1 green, 2 blue): the three channels are divided, clamped and converted independently.
SB_DEV void packChannel(float v, unsigned char* bitmap, const int index, const int c, const int iteration)
{
    if (iteration > B200_NB_MAX_ITERATIONS) v /= (float)(iteration - B200_NB_MAX_ITERATIONS + 1);
    v = (v > 1.f) ? 1.f : v;
    v = (v < 0.f) ? 0.f : v;
    int i;
    if (cSI.frameBufferType == B200_FT_BGR)
    {
        const int y = index / cSI.size.y, x = index % cSI.size.x;
        i = ((y + 1) * cSI.size.y - x - 1) * B200_COLOR_DEPTH + (2 - c);
    }
    else
        i = index * B200_COLOR_DEPTH + c;
    bitmap[i] = (unsigned char)(v * 255.f);
}

// Staged anaglyph frames: the two eyes of a pixel are two paths that end in different stages, and what resolvePixel does with them
// separates by component — the left eye's luma is the red channel, the right eye supplies green, blue, the first-hit depth and
// the ids (it is traced second in k_anaglyphRenderer, CudaRayTracer.cu:866-925, so its values are the ones that stay) — so each
// eye writes its own words of colorInfo and its own bytes of the frame.  Path tags carry the eye in bit 30 of the pixel index.
#define PATH_EYE_BIT 30
SB_DEV void resolveAnaglyphEye(const int index, const int eye, const float4 color, const int4 id, const float dof)
{
    const int iter = cSI.pathTracingIteration;
    float* ci = &cP.post[index].colorInfo.x;
    if (eye == 0)
    {
        const float r1 = anaglyphLuma(color);
        const float v = (iter <= B200_NB_MAX_ITERATIONS) ? r1 + 0.f : ci[0] + (r1 + 0.f);
        ci[0] = v;
        packChannel(v, cP.bitmap, index, 0, iter);
    }
    else
    {
        const float g = (iter <= B200_NB_MAX_ITERATIONS) ? 0.f + color.y : ci[1] + (0.f + color.y);
        const float b = (iter <= B200_NB_MAX_ITERATIONS) ? 0.f + color.z : ci[2] + (0.f + color.z);
        ci[1] = g; ci[2] = b;
        if (iter == 0) ci[3] = dof;
        cP.ids[index] = id;
        packChannel(g, cP.bitmap, index, 1, iter);
        packChannel(b, cP.bitmap, index, 2, iter);
    }
}

// ----------------------------------------------------------------------------------------------------
// Streamed output.  The reference's host protocol reads the frame AND the id buffer back after every frame
// (CudaKernel.cpp:304-313): 19 bytes per pixel, 39 MB per 1080p frame, 0.8 ms of PCIe time behind a 4.5 ms frame — and behind a
// 1.2 ms frame when eight GPUs share it.  When the caller's buffers are pinned (b200_register_host) and it reads every frame, the
// staged kernels write them directly instead: every tile (8 x 4 pixels) counts the paths it still has to end, the lane that ends
// the last one says so, and the warp copies the tile's finished ids (four full 128-byte lines) and RGB bytes (four runs of 24) from
// the device buffers into the mapped host buffers — posted writes over PCIe while the other tiles are still being traced.  A pixel
// the frame does not touch (pixelNeedsWork) keeps the value both sides already hold.  b200_d2h_bitmap then only waits for the stream.
// ----------------------------------------------------------------------------------------------------
// paths the warp's tile will end in this frame (stage 0, all lanes; before any of them can end)
SB_DEV void streamExpect(const int tile, const bool valid, const int eyes)
{
    if (cP.tileRemaining == nullptr) return;
    const unsigned int m = __ballot_sync(FULL_MASK, valid);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(cP.tileRemaining + tile, __popc(m) * eyes);
    __syncwarp();
}
// All lanes, where the warp is convergent again after a batch of paths (nothing of the paths is live any more): `ended` = this lane's
// path ended in the batch and its pixel `index` is written.  Every such lane counts its tile down, and the tiles this completes go
// to the host buffers: ids as four full 128-byte lines, RGB as four runs of 24 bytes.
SB_DEV void streamBatch(const bool ended, const int index)
{
    if (cP.tileRemaining == nullptr) return;
    if (!__any_sync(FULL_MASK, ended)) return;
    const int lane = threadIdx.x & 31;
    const int W = cSI.size.x;
    int doneTile = -1;
    if (ended)
    {
        const int x = index % W, y = index / W;
        const int tile = (y / TILE_H) * cP.tilesX + x / TILE_W;
        // The count is decremented with RELEASE semantics by the lane that wrote the pixel (its words are performed before the count
        // moves), not behind a __threadfence(): that is MEMBAR + CCTL.IVALL — it throws the SM's whole L1 away, 87 k times per frame
        // in pass 0 alone, and the walks of every warp on the SM pay for it (ncu, profiles/r02_ncu_streamed_frame.json: +0.17 ms in
        // k_stage_primary, long-scoreboard 2.8 -> 3.6 per issue).  atom.release is MEMBAR + ATOMG.
        int old;
        asm volatile("atom.release.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(cP.tileRemaining + tile), "r"(-1) : "memory");
        if (old == 1) doneTile = tile;
    }
    unsigned int d = __ballot_sync(FULL_MASK, doneTile >= 0);
    if (d == 0) return;
    // the finished tiles are read past L1 (ld.cg): the other warps' (other SMs') pixels were performed before their counts, and the count
    // that reached zero was seen after them — the reader's side of a fenced reduction needs no L1 invalidation of its own
    while (d)
    {
        const int src = __ffs(d) - 1;
        d &= d - 1;
        const int tile = __shfl_sync(FULL_MASK, doneTile, src);
        const int x0 = (tile % cP.tilesX) * TILE_W, y0 = (tile / cP.tilesX) * TILE_H;
        if (cP.hostIds)
        {
            const size_t i = (size_t)(y0 + lane / TILE_W) * W + x0 + (lane & (TILE_W - 1));
            cP.hostIds[i] = __ldcg(cP.ids + i);
        }
        if (cP.hostBitmap && lane < TILE_H * (TILE_W * B200_COLOR_DEPTH / 4))
        {
            // a row of the tile is TILE_W x 3 = 24 bytes at a multiple of 8 (the frame's width is a multiple of TILE_W): six words
            const int row = lane / (TILE_W * B200_COLOR_DEPTH / 4), w = lane % (TILE_W * B200_COLOR_DEPTH / 4);
            const size_t b = ((size_t)(y0 + row) * W + x0) * B200_COLOR_DEPTH + 4 * (size_t)w;
            *reinterpret_cast<unsigned int*>(cP.hostBitmap + b) = __ldcg(reinterpret_cast<const unsigned int*>(cP.bitmap + b));
        }
    }
}

// a staged path has ended: its pixel (or its eye's share of the pixel) is resolved
SB_DEV void endPath(const int tag, const float4 color, const int4 id, const float dof)
{
    const int index = tag & ((1 << PATH_EYE_BIT) - 1);
    if (cSI.cameraType == B200_CT_ANAGLYPH) resolveAnaglyphEye(index, (tag >> PATH_EYE_BIT) & 1, color, id, dof);
    else resolvePixel(index, color, f4(0.f, 0.f, 0.f, 0.f), id, dof, *reinterpret_cast<float4*>(&cP.post[index].colorInfo));
}

// pixels whose ray tree ended before this deepening pass need no work (:454-458)
SB_DEV bool pixelNeedsWork(const int4 id)
{
    const int iter = cSI.pathTracingIteration;
    return !(iter > id.y && id.w == 0 && iter > 0 && iter <= B200_NB_MAX_ITERATIONS);
}

// One pixel: CudaRayTracer.cu:437-563 (standard / orthographic / antialiased cameras) and :840-926
// (anaglyph), followed by k_default for that pixel.  The cameras differ only in how many ray trees a
// pixel owns and how they are combined, so they share one loop around a single launchRayTracing site:
//   standard/orthographic: 1 sample; antialiased: 4 offset samples + 1; anaglyph: left eye, right eye.
SB_DEV void renderPixel(const Rotation& rot, const bool inFrame, const int xIn, const int yIn, Counters& cnt, unsigned int& pixelsTraced)
{
    const int W = cSI.size.x, H = cSI.size.y;
    const int x = inFrame ? xIn : 0, y = inFrame ? yIn : 0;
    const int index = y * W + x;
    int4 id = cP.ids[index];
    // a lane without work still walks along with its warp (valid == false) because the walks are warp-synchronous
    const bool valid = inFrame && pixelNeedsWork(id);
    if (!__any_sync(FULL_MASK, valid)) return;
    if (valid) pixelsTraced++;
    const int camera = cSI.cameraType;
    const int iter = cSI.pathTracingIteration;
    const float3 rotationCenter = (camera == B200_CT_VR) ? cP.eye : f3(0.f, 0.f, 0.f);
    float dof = 0.f;
    const float4 stored = *reinterpret_cast<float4*>(&cP.post[index].colorInfo);
    const float ratio = (float)W / (float)H;
    const float stepx = ratio * cP.angles.w / (float)W, stepy = cP.angles.w / (float)H;
    const bool anaglyph = camera == B200_CT_ANAGLYPH, antialiased = camera == B200_CT_ANTIALIASED;

    float3 o = cP.eye, t = cP.target;
    if (camera == B200_CT_VR)
    {
        // k_3DVisionRenderer (CudaRayTracer.cu:953-1043): left half = left eye, right half = right eye.  The focus depth is
        // read from the accumulation buffer (a pixel this launch may be writing at iteration 0, as in the reference).
        const float focus = fabsf(cP.post[W / 2 * H / 2].colorInfo.w - cP.eye.z);
        const float eyeSeparation = cSI.eyeSeparation * (cP.target.z / focus);
        const int halfWidth = W / 2;
        const bool leftEye = x < halfWidth;
        o.x = leftEye ? cP.eye.x + eyeSeparation : cP.eye.x - eyeSeparation;
        const float xf = leftEye ? (float)(x - (W / 2) + halfWidth / 2) : (float)(x - (W / 2) - halfWidth / 2);
        t.x = leftEye ? cP.target.x - stepx * xf + cSI.eyeSeparation : cP.target.x - stepx * xf - cSI.eyeSeparation;
        t.y = cP.target.y + stepy * (float)(y - (H / 2));
        vectorRotation(o, rotationCenter, rot, true);
        vectorRotation(t, rotationCenter, rot, true);
    }
    else if (camera == B200_CT_PANORAMIC)
    {
        // k_fishEyeRenderer (:757-813): the image width spans 360 degrees about the vertical axis through the eye
        if (iter >= B200_NB_MAX_ITERATIONS)
        {
            const int rindex = (index + cSI.timestamp) % (cS.randomTableSize - 3);
            const float a = float(iter) / float(cSI.maxPathTracingIterations);
            t.x += rnd(rindex) * stored.w * cP.pp.param2 * a;
            t.y += rnd(rindex + 1) * stored.w * cP.pp.param2 * a;
            t.z += rnd(rindex + 2) * stored.w * cP.pp.param2 * a;
        }
        t.y = t.y + stepy * (float)(y - (H / 2));
        const float ay = cP.angles.y + (2.f * 3.14159265358979323846f / W) * (float)x;
        Rotation fish;
        fish.cx = cosf(0.f); fish.cy = cosf(ay); fish.cz = cosf(0.f);
        fish.sx = sinf(0.f); fish.sy = sinf(ay); fish.sz = sinf(0.f);
        vectorRotation(t, o, fish);
    }
    else if (!anaglyph) primaryRay(rot, x, y, index, stored.w, o, t);

    const int nSamples = anaglyph ? 2 : (antialiased ? 5 : 1);
    float4 color = f4(0.f, 0.f, 0.f, 0.f);
    float4 left = f4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int s = 0; s < nSamples; ++s)
    {
        if (anaglyph)
        {
            anaglyphRay(rot, x, y, s, o, t);
        }
        else if (antialiased && s < 4)
        {
            // the reference offsets the ORIGIN cumulatively (:504-514); the 5th tree uses the sum (= 0,0)
            o.x += (s == 0) ? 3.f : (s == 1) ? 5.f : (s == 2) ? -3.f : -5.f;
            o.y += (s == 0) ? 5.f : (s == 1) ? -3.f : (s == 2) ? -5.f : 3.f;
        }
        const float4 c = launchRayTracing(valid, index, o, t, dof, id, cnt);
        if (anaglyph && s == 0) left = c;
        else color += c;
    }

    if (!valid) return;
    resolvePixel(index, color, left, id, dof, stored);
}

// Persistent CTAs: every warp pulls 8x4-pixel tiles from one atomic queue until the frame is drained, so
// a warp stuck on deep bounce chains does not hold back the rest of its CTA or its SM.
// What a ray kernel does before its first walk: the frame's first kernel asks the L2 for the scene arrays the walks read (a slice per
// CTA), and every CTA stages the top of the tree in shared memory (trace.cuh "Bulk copies").
#ifdef DUMMY_SMEM_KB
__shared__ int s_dummy[DUMMY_SMEM_KB * 256]; // experiment: shared-memory footprint without any use of it (instruction-fetch stalls vs carve-out)
#endif
SB_DEV void framePrologue(const bool firstKernelOfFrame)
{
#ifdef DUMMY_SMEM_KB
    if (cP.tilesX < 0) s_dummy[threadIdx.x] = 1; // never true; keeps the array allocated
#endif
#ifdef SCENE_L2_PREFETCH
    if (firstKernelOfFrame && threadIdx.x == 0 && cS.nbUWide > 0)
    {
        scenePrefetchSlice(cS.uwnodes, (size_t)(cS.nbUWide + cS.nbUX) * 128);
        scenePrefetchSlice(cS.primRecs, (size_t)cS.nbPrimitives * 96);
    }
#endif
    topStage();
}

__global__ void __launch_bounds__(CTA_THREADS, MIN_CTAS_PER_SM) k_render()
{
    const int lane = threadIdx.x & 31;
    framePrologue(true);
    Counters cnt;
    cnt.rays = 0;
    unsigned int pixelsTraced = 0;
    // VectorUtils.cuh:108-114 evaluates these six per pixel and per rotated vector (fast-math sinf/cosf);
    // same intrinsics, once per thread.
    Rotation rot;
    rot.cx = cosf(cP.angles.x); rot.cy = cosf(cP.angles.y); rot.cz = cosf(cP.angles.z);
    rot.sx = sinf(cP.angles.x); rot.sy = sinf(cP.angles.y); rot.sz = sinf(cP.angles.z);
    while (true)
    {
        unsigned int k = 0;
        if (lane == 0) k = atomicAdd(cP.tileCounter, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= (unsigned int)cP.nbLocalTiles) break;
        const int tile = cP.tileOrder ? cP.tileOrder[k] : k * cP.worldSize + cP.rank; // interleaved tile ownership across GPUs
        const int tx = tile % cP.tilesX, ty = tile / cP.tilesX;
        const int x = tx * TILE_W + (lane & (TILE_W - 1));
        const int y = ty * TILE_H + (lane / TILE_W);
        renderPixel(rot, x < cSI.size.x && y < cSI.size.y, x, y, cnt, pixelsTraced);
        __syncwarp();
    }
    // one atomic per warp for the work counters
    unsigned int rays = cnt.rays, px = pixelsTraced;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        rays += __shfl_xor_sync(0xffffffffu, rays, o);
        px += __shfl_xor_sync(0xffffffffu, px, o);
    }
    if (lane == 0)
    {
        atomicAdd(cP.workCounters, (unsigned long long)rays);
        atomicAdd(cP.workCounters + 1, (unsigned long long)px);
    }
}

// ----------------------------------------------------------------------------------------------------
// Staged kernels.  In the kernel above a lane owns its pixel to the end of the ray tree, so after the first pass
// only the lanes whose hit reflects or refracts (a third of them on the molecule scene, fewer each bounce) still carry
// a ray while the registers of the others sit idle — and the walks are latency-bound, so useful resident lanes are
// what sets the ray rate.  Here each pass is its own launch over a compacted queue of the paths that are still alive:
//   k_stage_primary    pass 0 of every pixel of the owned tiles (tile queue, as above)
//   k_stage_pass(p)    pass p of the paths queued by pass p - 1
//   k_stage_reflected  the extra reflected ray of transparent + reflective first hits (CudaRayTracer.cu:296-315)
// A path that ends is folded and written (resolvePixel) by the stage that ends it.  Between stages a path is ~40 words
// in pathWords[word][slot] plus its colors[pass] / colorContributions[pass], slot = tile-major pixel number, so a
// warp's loads and stores are contiguous in stage 0 and sector-contiguous afterwards (queues keep tile order within
// a warp's push).  Same device functions as the megakernel (pathPass / pathReflectedRay / pathFinish), same results.
// Used for the one-ray-tree-per-pixel cameras without box-debug / global-illumination rays; the rest take k_render.
// ----------------------------------------------------------------------------------------------------
#define PATH_WORDS 38
#define QUEUE_COUNTERS (4 * (B200_NB_MAX_ITERATIONS + 2)) // per queue: pushed, handed out; per pass: warps at work, entries available (fused driver)
#define FUSED_CTR (2 * (B200_NB_MAX_ITERATIONS + 2))
SB_DEV void storePath(const size_t slot, const PathState& s, const int index)
{
    float* w = cP.pathWords + slot;
    const size_t n = cP.pathStride;
    int k = 0;
#define PUT(v) stOnce(w + (size_t)(k++) * n, (v)) // vec.cuh: written once, read once
#define PUTI(v) stOnce(w + (size_t)(k++) * n, __int_as_float(v))
    PUT(s.curO.x); PUT(s.curO.y); PUT(s.curO.z); PUT(s.curT.x); PUT(s.curT.y); PUT(s.curT.z);
    PUT(s.initialRefraction); PUTI(s.currentMaterialId);
    PUT(s.closestColor.x); PUT(s.closestColor.y); PUT(s.closestColor.z); PUT(s.closestColor.w);
    PUT(s.shadowIntensity);
    PUT(s.rBlinn.x); PUT(s.rBlinn.y); PUT(s.rBlinn.z); PUT(s.rBlinn.w);
    PUT(s.recursiveBlinn.x); PUT(s.recursiveBlinn.y); PUT(s.recursiveBlinn.z);
    PUT(s.latestIntersection.x); PUT(s.latestIntersection.y); PUT(s.latestIntersection.z);
    PUT(s.rayLength); PUT(s.depthOfField);
    PUTI(s.reflectedRays);
    PUT(s.reflO.x); PUT(s.reflO.y); PUT(s.reflO.z); PUT(s.reflT.x); PUT(s.reflT.y); PUT(s.reflT.z);
    PUT(s.reflectedRatio);
    PUTI(s.idx); PUTI(s.idz); PUTI(s.idw); PUTI(s.iteration); PUTI(index);
#undef PUT
#undef PUTI
}

SB_DEV void loadPath(const size_t slot, PathState& s, int& index)
{
    const float* w = cP.pathWords + slot;
    const size_t n = cP.pathStride;
    int k = 0;
#define GET() ldOnce(w + (size_t)(k++) * n) // past L1: see GlobalColors
#define GETI() __float_as_int(ldOnce(w + (size_t)(k++) * n))
    s.curO.x = GET(); s.curO.y = GET(); s.curO.z = GET(); s.curT.x = GET(); s.curT.y = GET(); s.curT.z = GET();
    s.initialRefraction = GET(); s.currentMaterialId = GETI();
    s.closestColor.x = GET(); s.closestColor.y = GET(); s.closestColor.z = GET(); s.closestColor.w = GET();
    s.shadowIntensity = GET();
    s.rBlinn.x = GET(); s.rBlinn.y = GET(); s.rBlinn.z = GET(); s.rBlinn.w = GET();
    s.recursiveBlinn.x = GET(); s.recursiveBlinn.y = GET(); s.recursiveBlinn.z = GET(); s.recursiveBlinn.w = 0.f;
    s.latestIntersection.x = GET(); s.latestIntersection.y = GET(); s.latestIntersection.z = GET();
    s.rayLength = GET(); s.depthOfField = GET();
    s.reflectedRays = GETI();
    s.reflO.x = GET(); s.reflO.y = GET(); s.reflO.z = GET(); s.reflT.x = GET(); s.reflT.y = GET(); s.reflT.z = GET();
    s.reflectedRatio = GET();
    s.idx = GETI(); s.idz = GETI(); s.idw = GETI(); s.iteration = GETI(); index = GETI();
#undef GET
#undef GETI
    s.carryon = true;
    s.giO = f3(0.f, 0.f, 0.f); s.giT = f3(0.f, 0.f, 0.f); s.pathTracingRatio = 0.f; s.useGlobalIllumination = false;
    s.colorBox = f4(0.f, 0.f, 0.f, 0.f);
}

// warp-aggregated push of the lanes with `want` onto queue q.  The fused driver's consumers run in the same launch: there an entry is
// slot + 1 written into a queue that was zeroed before the frame, AFTER a device-wide fence behind everything the lane stored for
// the path (storePath), so whoever finds the entry non-zero finds the parked path too; the launch-per-pass drivers store the slot.
SB_DEV void pushPaths(const int q, const bool want, const size_t slot)
{
    const unsigned int m = __ballot_sync(FULL_MASK, want);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(cP.queueCounters + 2 * q, (unsigned int)__popc(m));
    base = __shfl_sync(FULL_MASK, base, __ffs(m) - 1);
    if (cP.fusedQueues)
    {
        // published with RELEASE semantics by the lane that parked the path (MEMBAR + store), not behind a __threadfence(), whose
        // CCTL.IVALL throws the SM's L1 away with every push (see streamBatch)
        if (want)
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(cP.pathQueues + (size_t)q * cP.pathStride + base + __popc(m & ((1u << lane) - 1u))), "r"((int)slot + 1) : "memory");
        // the consumers' semaphore (a consumer that is handed an entry before it is written waits for it to turn non-zero)
        if (lane == __ffs(m) - 1) atomicAdd(cP.queueCounters + FUSED_CTR + (B200_NB_MAX_ITERATIONS + 2) + q, (unsigned int)__popc(m));
        return;
    }
    if (want) cP.pathQueues[(size_t)q * cP.pathStride + base + __popc(m & ((1u << lane) - 1u))] = (int)slot;
}


// after pass `pass` of a path (all 32 lanes call; `has` = this lane carries one): queue it for the next stage, or end it
// returns whether the lane's path ended here (its pixel is written)
SB_DEV bool routePath(const bool has, const PathState& s, const GlobalColors& C, const int pass, const size_t slot, const int tag)
{
    const bool cont = has && s.carryon && s.rayLength < cSI.viewDistance && pass + 1 < cP.maxIteration;
    const bool refl = has && !cont && cSI.graphicsLevel >= B200_GL_REFLECTIONS && s.reflectedRays != -1;
    if (cont || refl) storePath(slot, s, tag);
    pushPaths(passQueue(pass + 1), cont, slot);
    pushPaths(reflectedQueue(), refl, slot);
    const bool ends = has && !cont && !refl;
    if (ends)
    {
        const float4 color = pathFinish(s, C, true);
        const int4 id = make_int4(s.idx, s.iteration, s.idz, s.idw);
        endPath(tag, color, id, s.depthOfField);
    }
    return ends;
}

SB_DEV void flushCounters(const unsigned int raysIn, const unsigned int pxIn)
{
    unsigned int rays = raysIn, px = pxIn;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        rays += __shfl_xor_sync(0xffffffffu, rays, o);
        px += __shfl_xor_sync(0xffffffffu, px, o);
    }
    if ((threadIdx.x & 31) == 0)
    {
        if (rays) atomicAdd(cP.workCounters, (unsigned long long)rays);
        if (px) atomicAdd(cP.workCounters + 1, (unsigned long long)px);
    }
}

#ifndef MIN_CTAS_PRIMARY
#define MIN_CTAS_PRIMARY MIN_CTAS_PER_SM
#endif
#ifndef MIN_CTAS_PASS
#define MIN_CTAS_PASS MIN_CTAS_PER_SM
#endif
template <bool STREAM>
__global__ void __launch_bounds__(CTA_THREADS, MIN_CTAS_PRIMARY) k_stage_primary()
{
    const int lane = threadIdx.x & 31;
    framePrologue(true);
    Counters cnt;
    cnt.rays = 0;
    unsigned int pixelsTraced = 0;
    Rotation rot;
    rot.cx = cosf(cP.angles.x); rot.cy = cosf(cP.angles.y); rot.cz = cosf(cP.angles.z);
    rot.sx = sinf(cP.angles.x); rot.sy = sinf(cP.angles.y); rot.sz = sinf(cP.angles.z);
    while (true)
    {
        unsigned int k = 0;
        if (lane == 0) k = atomicAdd(cP.tileCounter, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= (unsigned int)cP.nbLocalTiles) break;
        const int tile = cP.tileOrder ? cP.tileOrder[k] : k * cP.worldSize + cP.rank;
        const int tx = tile % cP.tilesX, ty = tile / cP.tilesX;
        const int xIn = tx * TILE_W + (lane & (TILE_W - 1));
        const int yIn = ty * TILE_H + (lane / TILE_W);
        const bool inFrame = xIn < cSI.size.x && yIn < cSI.size.y;
        const int x = inFrame ? xIn : 0, y = inFrame ? yIn : 0;
        const int index = y * cSI.size.x + x;
        const int4 id = cP.ids[index];
        const bool valid = inFrame && pixelNeedsWork(id);
        if (!__any_sync(FULL_MASK, valid)) continue;
        if (valid) pixelsTraced++;
        const bool anaglyph = cSI.cameraType == B200_CT_ANAGLYPH;
        if (STREAM) streamExpect(tile, valid, anaglyph ? 2 : 1);
        const float storedDepth = cP.post[index].colorInfo.w;
#pragma unroll 1
        for (int eye = 0; eye < (anaglyph ? 2 : 1); ++eye)
        {
            // the ids were read above, before either eye's path can end and rewrite them
            float3 o, t;
            if (anaglyph) anaglyphRay(rot, x, y, eye, o, t);
            else primaryRay(rot, x, y, index, storedDepth, o, t);
            const size_t slot = (size_t)k * 32 + lane + (size_t)eye * cP.eyeStride;
            PathState s;
            pathInit(s, o, t);
            GlobalColors C;
            C.c = cP.pathColors; C.k = cP.pathContributions; C.slot = slot; C.stride = cP.pathStride;
            pathPass(s, C, 0, valid, index, o, cP.packetMask, cnt);
            const bool ended = routePath(valid, s, C, 0, slot, index | (eye << PATH_EYE_BIT));
            __syncwarp();
            // STREAM: the kernel instance that counts tiles and writes the host buffers (streamed output), at the one place of a
            // batch where nothing of its paths is live; the other instance carries none of it (hooks inside routePath cost the
            // frame 3 % even when idle: 4.62 -> 4.76 ms on config 2)
            if (STREAM) streamBatch(ended, index);
        }
    }
    flushCounters(cnt.rays, pixelsTraced);
}

// pass >= 1: 32 queue entries per warp at a time
template <bool STREAM>
__global__ void __launch_bounds__(CTA_THREADS, MIN_CTAS_PASS) k_stage_pass(const int pass)
{
    const int lane = threadIdx.x & 31;
    framePrologue(false);
    Counters cnt;
    cnt.rays = 0;
    const int q = passQueue(pass);
    const unsigned int count = cP.queueCounters[2 * q];
    // the deeper the pass, the fewer passes remain to carry finished lanes through: the bound grows with the pass number
    const bool fuseTail = (unsigned long long)count * 100ull <=
                          (unsigned long long)gridDim.x * blockDim.x * (unsigned long long)cP.fuseTailPercent * (unsigned long long)pass;
    while (true)
    {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(cP.queueCounters + 2 * q + 1, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
        const bool has = base + lane < count;
        const size_t slot = has ? (size_t)cP.pathQueues[(size_t)q * cP.pathStride + base + lane] : 0;
        // the walk first, with only the ray live: the rest of the path is loaded after it, so the call into the walk has
        // next to nothing to save (call-boundary spills were most of the kernel's local-memory traffic)
        Hit hit;
        hit.prim = -1; hit.p = f3(0.f, 0.f, 0.f); hit.flags = 0;
        const bool separateWalk = cS.nbUWide > 0 && cSI.renderBoxes == 0;
#if UW_GROUP
        if (separateWalk)
        {
            const float* w = cP.pathWords + slot;
            const size_t n = cP.pathStride;
            hit = closestHitGroup(f3(w[0], w[n], w[2 * n]), f3(w[3 * n], w[4 * n], w[5 * n]), pass, __float_as_int(w[7 * n]), has);
        }
#else
        if (separateWalk && has)
        {
            const float* w = cP.pathWords + slot;
            const size_t n = cP.pathStride;
            hit = closestHitOrderIndependent(f3(w[0], w[n], w[2 * n]), f3(w[3 * n], w[4 * n], w[5 * n]), pass, __float_as_int(w[7 * n]));
        }
#endif
        PathState s;
        int tag = 0;
        loadPath(slot, s, tag);
        if (!has) tag = 0;
        const int index = tag & ((1 << PATH_EYE_BIT) - 1);
        GlobalColors C;
        C.c = cP.pathColors; C.k = cP.pathContributions; C.slot = slot; C.stride = cP.pathStride;
        pathPass(s, C, pass, has, index, f3(0.f, 0.f, 0.f), 0, cnt, separateWalk ? &hit : nullptr);
        if (!fuseTail)
        {
            const bool ended = routePath(has, s, C, pass, slot, tag);
            __syncwarp();
            if (STREAM) streamBatch(ended, index);
            continue;
        }
        // Few paths left (at most one batch per resident warp: deep passes, small frames, a 1/8 share of a frame): another
        // launch per pass would cost its tail and a round trip of the path through memory for nothing, since there are no
        // other paths to fill the lanes with.  The warp keeps its paths in registers to the end of their ray trees, the way
        // k_render does; the launches of the remaining passes find empty queues.
        bool live = has, ended = false;
        for (int p = pass;; ++p)
        {
            const bool cont = live && s.carryon && s.rayLength < cSI.viewDistance && p + 1 < cP.maxIteration;
            ended |= routePath(live && !cont, s, C, p, slot, tag); // ends here: reflected-ray stage or pixel
            live = cont;
            if (!__any_sync(FULL_MASK, live)) break;
#if UW_GROUP
            if (separateWalk) hit = closestHitGroup(s.curO, s.curT, p + 1, s.currentMaterialId, live);
#else
            if (separateWalk && live) hit = closestHitOrderIndependent(s.curO, s.curT, p + 1, s.currentMaterialId);
#endif
            pathPass(s, C, p + 1, live, index, f3(0.f, 0.f, 0.f), 0, cnt, separateWalk ? &hit : nullptr);
        }
        __syncwarp();
        if (STREAM) streamBatch(ended, index);
    }
    flushCounters(cnt.rays, 0);
}

// ----------------------------------------------------------------------------------------------------
// Fused stages (b200_set_option(6, 2)): the staged kernels as ONE persistent launch per frame.  A launch per pass ends with a tail
// — the pass is as long as its slowest warp while the next pass's paths sit in their queue — and a frame of p passes pays p tails:
// a tenth of a 1080p frame on one GPU, a third of a 1/8 share of it (8 GPUs), and ten tails per 4K frame of config 4.  Here a warp
// that runs out of work of one pass takes work of another:
//   1. 32 entries of the DEEPEST pass queue that has that many (depth first keeps the parked paths short-lived);
//   2. else a tile of primary rays (pass 0);
//   3. else the remainder (< 32 entries) of a queue that can get no more entries — every earlier pass is done;
//   4. else it sleeps a little and looks again, until the last pass is done.
// "Pass p is done" = its queue is closed (pass p - 1 is done; the tile queue is closed from the start), every entry is claimed, and
// no warp is at work on it (counter raised BEFORE a claim, lowered after the warp's pushes are reserved).  Compaction stays what it
// was: batches are full except one per pass.  Entries and parked paths cross SMs inside the launch: a producer fences and then writes
// slot + 1 into its reserved places of a zeroed queue, a consumer waits for its entry to turn non-zero and reads the path past L1
// (loadPath, GlobalColors).
// Same device functions, same rays, same results as the launch-per-pass drivers; one-ray-tree-per-pixel cameras with the
// order-independent walks only (the others keep the launches).
// ----------------------------------------------------------------------------------------------------
SB_DEV unsigned int ctrLoad(const unsigned int* p) { return *reinterpret_cast<const volatile unsigned int*>(p); }

__global__ void __launch_bounds__(CTA_THREADS, MIN_CTAS_PASS) k_stage_fused()
{
    const int lane = threadIdx.x & 31;
    framePrologue(true);
    Counters cnt;
    cnt.rays = 0;
    unsigned int pixelsTraced = 0;
    Rotation rot;
    rot.cx = cosf(cP.angles.x); rot.cy = cosf(cP.angles.y); rot.cz = cosf(cP.angles.z);
    rot.sx = sinf(cP.angles.x); rot.sy = sinf(cP.angles.y); rot.sz = sinf(cP.angles.z);
    const int maxIt = cP.maxIteration;
    unsigned int* const Q = cP.queueCounters;
    unsigned int* const atWork = Q + FUSED_CTR;
    unsigned int* const avail = Q + FUSED_CTR + (B200_NB_MAX_ITERATIONS + 2);
    bool tilesLeft = true;
    while (true)
    {
        // ---- what to do next (lane 0 decides)
        int pass = -2;            // -2: nothing right now, -3: the frame is done, 0: a tile, >= 1: a batch of that pass
        unsigned int base = 0, n = 0;
        if (lane == 0)
        {
            // 1. a full batch, deepest pass first.  avail[p] is a semaphore (entries reserved by producers minus entries taken):
            //    take 32 and give them back if there were not that many — no retry loops on a contended counter
            for (int p = maxIt - 1; p >= 1 && pass == -2; --p)
            {
                if ((int)ctrLoad(avail + p) < 32) continue;
                atomicAdd(atWork + p, 1u);
                if ((int)atomicSub(avail + p, 32u) >= 32) { pass = p; n = 32u; base = atomicAdd(Q + 2 * p + 1, 32u); }
                else { atomicAdd(avail + p, 32u); atomicSub(atWork + p, 1u); }
            }
            // 2. a tile
            if (pass == -2 && tilesLeft)
            {
                atomicAdd(atWork, 1u);
                const unsigned int k = atomicAdd(cP.tileCounter, 1u);
                if (k < (unsigned int)cP.nbLocalTiles) { pass = 0; base = k; }
                else { atomicSub(atWork, 1u); tilesLeft = false; }
            }
            // 2b. nothing full and no tile: rather than idle until a queue closes, a partial batch of the shallowest queue that has
            //     FUSED_MIN_PARTIAL entries (the idle warp costs nothing, and what it produces lets the deeper queues fill earlier)
#ifndef FUSED_MIN_PARTIAL
#define FUSED_MIN_PARTIAL 1
#endif
            for (int p = 1; p < maxIt && pass == -2 && FUSED_MIN_PARTIAL < 32; ++p)
            {
                if ((int)ctrLoad(avail + p) < FUSED_MIN_PARTIAL) continue;
                atomicAdd(atWork + p, 1u);
                const int had = (int)atomicSub(avail + p, 32u);
                if (had >= FUSED_MIN_PARTIAL)
                {
                    const unsigned int take = had < 32 ? (unsigned int)had : 32u;
                    if (take < 32u) atomicAdd(avail + p, 32u - take);
                    pass = p; n = take; base = atomicAdd(Q + 2 * p + 1, take);
                }
                else { atomicAdd(avail + p, 32u); atomicSub(atWork + p, 1u); }
            }
            // 3. / 4. remainders of closed queues, or the end
            if (pass == -2)
            {
                bool closed = ctrLoad(cP.tileCounter) >= (unsigned int)cP.nbLocalTiles && ctrLoad(atWork) == 0u; // pass 0 done
                for (int p = 1; p < maxIt && closed; ++p)
                {
                    __threadfence();
                    // queue p is closed: what is reserved is all there will be
                    if ((int)ctrLoad(avail + p) > 0)
                    {
                        atomicAdd(atWork + p, 1u);
                        const int had = (int)atomicSub(avail + p, 32u);
                        if (had > 0)
                        {
                            const unsigned int take = had < 32 ? (unsigned int)had : 32u;
                            if (take < 32u) atomicAdd(avail + p, 32u - take);
                            pass = p; n = take; base = atomicAdd(Q + 2 * p + 1, take);
                        }
                        else { atomicAdd(avail + p, 32u); atomicSub(atWork + p, 1u); }
                        closed = false; // taken, or somebody else was faster: look again next time
                        break;
                    }
                    closed = ctrLoad(Q + 2 * p + 1) == ctrLoad(Q + 2 * p) && ctrLoad(atWork + p) == 0u; // pass p done?
                }
                if (pass == -2 && closed) pass = -3;
            }
        }
        pass = __shfl_sync(FULL_MASK, pass, 0);
        if (pass == -3) break;
        if (pass == -2) { __nanosleep(1000); continue; }
        base = __shfl_sync(FULL_MASK, base, 0);
        n = __shfl_sync(FULL_MASK, n, 0);
        tilesLeft = __shfl_sync(FULL_MASK, tilesLeft ? 1 : 0, 0) != 0;

        // ---- one batch: 32 pixels of a tile (pass 0) or up to 32 parked paths (pass >= 1)
        bool has;
        size_t slot;
        int tag = 0;
        float3 rayO = f3(0.f, 0.f, 0.f), o = f3(0.f, 0.f, 0.f), t = f3(0.f, 0.f, 0.f);
        int material = -2;
        if (pass == 0)
        {
            const unsigned int k = base;
            const int tile = cP.tileOrder ? cP.tileOrder[k] : k * cP.worldSize + cP.rank;
            const int tx = tile % cP.tilesX, ty = tile / cP.tilesX;
            const int xIn = tx * TILE_W + (lane & (TILE_W - 1));
            const int yIn = ty * TILE_H + (lane / TILE_W);
            const bool inFrame = xIn < cSI.size.x && yIn < cSI.size.y;
            const int x = inFrame ? xIn : 0, y = inFrame ? yIn : 0;
            const int index = y * cSI.size.x + x;
            has = inFrame && pixelNeedsWork(cP.ids[index]);
            if (has) pixelsTraced++;
            primaryRay(rot, x, y, index, cP.post[index].colorInfo.w, o, t);
            rayO = o;
            slot = (size_t)k * 32 + lane;
            tag = index;
        }
        else
        {
            has = (unsigned int)lane < n;
            slot = 0;
            if (has)
            {
                // the entry is reserved; its producer writes it right after its fence
                const volatile int* entry = cP.pathQueues + (size_t)passQueue(pass) * cP.pathStride + base + lane;
                int e;
                while ((e = *entry) == 0) __nanosleep(20);
                // the parked path is read through addresses that depend on e, past L1 (ld.cg / no-allocate): no fence, no invalidation
                slot = (size_t)(e - 1);
                const float* w = cP.pathWords + slot;
                const size_t s = cP.pathStride;
                o = f3(__ldcg(w), __ldcg(w + s), __ldcg(w + 2 * s)); t = f3(__ldcg(w + 3 * s), __ldcg(w + 4 * s), __ldcg(w + 5 * s));
                material = __float_as_int(__ldcg(w + 7 * s));
            }
        }
        if (__any_sync(FULL_MASK, has))
        {
            // the walk first, with only the ray live (k_stage_pass has the why)
            Hit hit;
            hit.prim = -1; hit.p = f3(0.f, 0.f, 0.f); hit.flags = 0;
            if (has) hit = closestHitOrderIndependent(o, t, pass, material);
            PathState s;
            if (pass == 0) pathInit(s, o, t);
            else
            {
                loadPath(slot, s, tag);
                if (!has) tag = 0;
            }
            const int index = tag & ((1 << PATH_EYE_BIT) - 1);
            GlobalColors C;
            C.c = cP.pathColors; C.k = cP.pathContributions; C.slot = slot; C.stride = cP.pathStride;
            pathPass(s, C, pass, has, index, rayO, 0, cnt, &hit);
            routePath(has, s, C, pass, slot, tag);
        }
        __syncwarp();
        // this warp's pushes are reserved (pushPaths): it is no longer at work on the pass
        if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(atWork + pass), "r"(0xffffffffu) : "memory"); // behind the __syncwarp above
    }
    flushCounters(cnt.rays, pixelsTraced);
}



template <bool STREAM>
__global__ void __launch_bounds__(CTA_THREADS, MIN_CTAS_PER_SM) k_stage_reflected()
{
    const int lane = threadIdx.x & 31;
    framePrologue(false);
    Counters cnt;
    cnt.rays = 0;
    const int q = reflectedQueue();
    const unsigned int count = cP.queueCounters[2 * q];
    while (true)
    {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(cP.queueCounters + 2 * q + 1, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
        const bool has = base + lane < count;
        const size_t slot = has ? (size_t)(cP.pathQueues[(size_t)q * cP.pathStride + base + lane] - cP.fusedQueues) : 0;
        PathState s;
        int tag = 0;
        loadPath(slot, s, tag);
        if (!has) tag = 0;
        const int index = tag & ((1 << PATH_EYE_BIT) - 1);
        GlobalColors C;
        C.c = cP.pathColors; C.k = cP.pathContributions; C.slot = slot; C.stride = cP.pathStride;
        pathReflectedRay(s, C, has, index, 0, cnt);
        if (has)
        {
            const float4 color = pathFinish(s, C, true);
            const int4 id = make_int4(s.idx, s.iteration, s.idz, s.idw);
            endPath(tag, color, id, s.depthOfField);
        }
        __syncwarp();
        if (STREAM) streamBatch(has, index);
    }
    flushCounters(cnt.rays, 0);
}

// The unit walk's form of the walk trees (trace.cuh nodeKeysRegs): every child box of every node record as centre and half extent,
// the half extent rounded up so that the box contains the lo/hi box it came from; refs and the child count are copied.  One thread
// per record, after every build, re-fit or adoption of the lo/hi array (which stays what the builders and the re-fit work on).
__global__ void __launch_bounds__(256) k_nodes_centre_half(const float4* __restrict__ lohi, float4* __restrict__ ch, const int nbRecords)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbRecords) return;
    const float4* n = lohi + 8 * (size_t)i;
    float4* o = ch + 8 * (size_t)i;
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        const float4 lo = n[a], hi = n[3 + a];
        float4 c, h;
#define CH(C)                                                                                      \
    if (lo.C <= hi.C) { c.C = 0.5f * (lo.C + hi.C); h.C = fmaxf(__fsub_ru(hi.C, c.C), __fsub_ru(c.C, lo.C)); } \
    else { c.C = 0.f; h.C = -INFINITY; } // empty slot (lo = 3e38, hi = -3e38): never entered
        CH(x) CH(y) CH(z) CH(w)
#undef CH
        o[a] = c; o[3 + a] = h;
    }
    o[6] = n[6]; o[7] = n[7];
}

// Streamed output, whole-frame form: every pixel of the tiles this GPU owns goes from the device buffers to the host buffers in one
// launch after the frame's other kernels.  For the frames whose kernels do not count tiles (single-kernel cameras, an effect pass,
// sizes that are not whole tiles), and to bring a newly named target up to date (b200_stream_target).
__global__ void __launch_bounds__(256) k_stream_own_tiles(const int4* ids, const unsigned char* bitmap, int4* hostIds, unsigned char* hostBitmap,
                                                          const int W, const int H, const int tilesX, const int nbLocalTiles, const int rank, const int world, const int bgr)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < nbLocalTiles; k += warps)
    {
        const int tile = k * world + rank;
        const int x = (tile % tilesX) * TILE_W + (lane & (TILE_W - 1)), y = (tile / tilesX) * TILE_H + lane / TILE_W;
        if (x >= W || y >= H) continue;
        const int index = y * W + x;
        if (hostIds) hostIds[index] = ids[index];
        if (hostBitmap)
        {
            // where packPixelAt puts the pixel
            const int i = bgr ? (((index / H) + 1) * H - (index % W) - 1) * B200_COLOR_DEPTH : index * B200_COLOR_DEPTH;
            hostBitmap[i] = bitmap[i]; hostBitmap[i + 1] = bitmap[i + 1]; hostBitmap[i + 2] = bitmap[i + 2];
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// Post-processing effects: cudaRender's second pass (CudaRayTracer.cu:1853-1886) — k_depthOfField :1081-1119,
// k_ambiantOcclusion :1127-1180, k_radiosity :1188-1228, k_filter :1236-1330, k_cartoon :1338-1357.  Each output pixel
// gathers from the float accumulation buffer of the finished frame (and ids.z for radiosity) and rewrites its RGB8
// value, so the pass is one more launch on the render stream after the ray kernels, one thread per pixel.
// ----------------------------------------------------------------------------------------------------
__constant__ float cFilterKernels[6][5][5] = {
    {{-1, -1, 0, 0, 0}, {-1, 0, 1, 0, 0}, {0, 1, 1, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},                 // emboss
    {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {-1, -1, 2, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},                  // find edges
    {{-1, -1, -1, 0, 0}, {-1, 9, -1, 0, 0}, {-1, -1, -1, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},            // sharpen
    {{0, 0.2f, 0, 0, 0}, {0.2f, 0.2f, 0.2f, 0, 0}, {0, 0.2f, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},     // blur
    {{1, 0, 0, 0, 0}, {0, 1, 0, 0, 0}, {0, 0, 1, 0, 0}, {0, 0, 0, 1, 0}, {0, 0, 0, 0, 1}},                    // motion blur
    {{-1, -1, -1, -1, -1}, {-1, 2, 2, 2, -1}, {-1, 2, 8, 2, -1}, {-1, 2, 2, 2, -1}, {-1, -1, -1, -1, -1}}};  // subtle sharpen
__constant__ int cFilterSize[6] = {3, 5, 3, 3, 5, 5};
__constant__ float cFilterFactor[6] = {1.f, 1.f, 1.f, 1.f, 0.2f, 0.125f};
__constant__ float cFilterBias[6] = {128.f, 0.f, 0.f, 0.f, 0.f, 0.f};

SB_DEV float4 postColor(const int i) { return *reinterpret_cast<const float4*>(&cP.post[i].colorInfo); }

__global__ void __launch_bounds__(256) k_post_process()
{
    const int W = cSI.size.x, H = cSI.size.y, wh = W * H;
    const int iter = cSI.pathTracingIteration;
    const b200_PostProcessingInfo pp = cP.pp;
    for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < wh; index += gridDim.x * blockDim.x)
    {
        const int x = index % W, y = index / W;
        float4 out = f4(0.f, 0.f, 0.f, 0.f);
        switch (pp.type)
        {
        case B200_PPE_DEPTH_OF_FIELD:
        {
            const float depth = fabsf(cP.post[index].colorInfo.w - pp.param1) / cSI.viewDistance;
            for (int i = 0; i < pp.param3; ++i)
            {
                const int ix = i % wh, iy = (i + 1000) % wh;
                const int xx = x + depth * rnd(ix) * pp.param2;
                const int yy = y + depth * rnd(iy) * pp.param2;
                if (xx >= 0 && xx < W && yy >= 0 && yy < H)
                {
                    const int li = yy * W + xx;
                    if (li >= 0 && li < wh) out += postColor(li);
                }
                else
                    out += postColor(index);
            }
            out /= (float)pp.param3;
            if (iter > B200_NB_MAX_ITERATIONS) out /= (float)(iter - B200_NB_MAX_ITERATIONS + 1);
            break;
        }
        case B200_PPE_AMBIENT_OCCLUSION:
        {
            out = postColor(index);
            const float depth = out.w;
            float occ = 0.f, c = 0.f;
            int i = 0;
            for (int X = -16; X < 16; X += 2)
                for (int Y = -16; Y < 16; Y += 2)
                {
                    const int ix = i % wh, iy = (i + 100) % wh;
                    ++i;
                    c += 1.f;
                    const int xx = x + (X * pp.param2 * rnd(ix) / 10.f);
                    const int yy = y + (Y * pp.param2 * rnd(iy) / 10.f);
                    if (xx >= 0 && xx < W && yy >= 0 && yy < H)
                    {
                        if (cP.post[yy * W + xx].colorInfo.w >= depth) occ += 1.f;
                    }
                    else
                        occ += 1.f;
                }
            occ /= c;
            occ += 0.3f; // ambient light
            if (occ < 1.f) { out.x *= occ; out.y *= occ; out.z *= occ; }
            if (iter > B200_NB_MAX_ITERATIONS) out /= (float)(iter - B200_NB_MAX_ITERATIONS + 1);
            saturate4(out);
            break;
        }
        case B200_PPE_RADIOSITY:
        {
            const int div = (iter > B200_NB_MAX_ITERATIONS) ? (iter - B200_NB_MAX_ITERATIONS + 1) : 1;
            for (int i = 0; i < pp.param3; ++i)
            {
                const int ix = (i + iter) % wh, iy = (i + 100 + iter) % wh;
                const int xx = x + rnd(ix) * pp.param2;
                const int yy = y + rnd(iy) * pp.param2;
                out += postColor(index);
                if (xx >= 0 && xx < W && yy >= 0 && yy < H)
                {
                    const int li = yy * W + xx;
                    out += postColor(li) * (float)cP.ids[li].z / 256.f;
                }
            }
            out /= (float)pp.param3;
            out /= (float)div;
            saturate4(out);
            break;
        }
        case B200_PPE_FILTER:
        {
            if ((unsigned int)pp.param3 < 6u)
            {
                const int f = pp.param3, n = cFilterSize[f];
                float4 acc = f4(0.f, 0.f, 0.f, 0.f);
                for (int fx = 0; fx < n; ++fx)
                    for (int fy = 0; fy < n; ++fy)
                    {
                        const int imx = (x - n / 2 + fx + W) % W, imy = (y - n / 2 + fy + H) % H;
                        float4 c = postColor(imy * W + imx);
                        if (iter > B200_NB_MAX_ITERATIONS) c /= (float)(iter - B200_NB_MAX_ITERATIONS + 1);
                        const float k = cFilterKernels[f][fx][fy];
                        acc.x += c.x * k; acc.y += c.y * k; acc.z += c.z * k;
                    }
                out.x += fminf(fmaxf(cFilterFactor[f] * acc.x + cFilterBias[f] / 255.f, 0.f), 1.f);
                out.y += fminf(fmaxf(cFilterFactor[f] * acc.y + cFilterBias[f] / 255.f, 0.f), 1.f);
                out.z += fminf(fmaxf(cFilterFactor[f] * acc.z + cFilterBias[f] / 255.f, 0.f), 1.f);
            }
            saturate4(out);
            break;
        }
        case B200_PPE_CARTOON:
        {
            const float depth = cSI.viewDistance / fabsf(cP.post[index].colorInfo.w - pp.param1);
            out = f4(depth, depth, depth, 0.f);
            saturate4(out);
            break;
        }
        default:
            out = postColor(index);
            if (iter > B200_NB_MAX_ITERATIONS) out /= (float)(iter - B200_NB_MAX_ITERATIONS + 1);
            break;
        }
        // makeColor (GeometryShaders.cuh:132-165); the iteration average was applied above where the effect has one
        out.x = (out.x > 1.f) ? 1.f : out.x; out.y = (out.y > 1.f) ? 1.f : out.y; out.z = (out.z > 1.f) ? 1.f : out.z;
        out.x = (out.x < 0.f) ? 0.f : out.x; out.y = (out.y < 0.f) ? 0.f : out.y; out.z = (out.z < 0.f) ? 0.f : out.z;
        if (cSI.frameBufferType == B200_FT_BGR)
        {
            const int yy = index / cSI.size.y, xx = index % cSI.size.x;
            const int i = ((yy + 1) * cSI.size.y - xx - 1) * B200_COLOR_DEPTH;
            cP.bitmap[i] = (unsigned char)(out.z * 255.f);
            cP.bitmap[i + 1] = (unsigned char)(out.y * 255.f);
            cP.bitmap[i + 2] = (unsigned char)(out.x * 255.f);
        }
        else
        {
            const int i = index * B200_COLOR_DEPTH;
            cP.bitmap[i] = (unsigned char)(out.x * 255.f);
            cP.bitmap[i + 1] = (unsigned char)(out.y * 255.f);
            cP.bitmap[i + 2] = (unsigned char)(out.z * 255.f);
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// Sample-split accumulation (north_star: "sample accumulation optionally split by GPU ... reduced with NCCL").  Past
// NB_MAX_ITERATIONS a frame only ADDS its sample to colorInfo.xyz (CudaRayTracer.cu:550-562) and k_default divides the sum by
// the number of samples when it packs the pixel (:1066-1070), so samples are exchangeable: every GPU renders whole frames for
// its share of the iterations into its own accumulation buffer, the partial sums are reduced onto the root over NVLink, and the
// root stores the total and packs.  These three kernels are the device side: one float4 per pixel crosses the wire.
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_accum_clear(b200_PostProcessingBuffer* post, const int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        float4 c = *reinterpret_cast<float4*>(&post[i].colorInfo);
        c.x = 0.f; c.y = 0.f; c.z = 0.f; // .w is the first-hit depth of iteration 0: kept
        *reinterpret_cast<float4*>(&post[i].colorInfo) = c;
    }
}
__global__ void __launch_bounds__(256) k_accum_export(const b200_PostProcessingBuffer* post, float4* dst, const int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float4 c = *reinterpret_cast<const float4*>(&post[i].colorInfo);
        dst[i] = make_float4(c.x, c.y, c.z, 0.f);
    }
}
// the reduced sums come back: stored, divided by the sample count and packed (k_default fused)
__global__ void __launch_bounds__(256) k_accum_import_pack(b200_PostProcessingBuffer* post, const float4* src, unsigned char* bitmap, const int n,
                                                           const int iteration)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float4 t = src[i];
        float4 c = *reinterpret_cast<float4*>(&post[i].colorInfo);
        c.x = t.x; c.y = t.y; c.z = t.z;
        *reinterpret_cast<float4*>(&post[i].colorInfo) = c;
        packPixelAt(c, bitmap, i, iteration);
    }
}

// FP32 FMA throughput probe (b200_measure_fp32_peak): 8 independent dependent-FFMA chains per thread, 16 steps each per iteration
#define FP32_PEAK_FMAS_PER_ITER (8 * 16)
__global__ void __launch_bounds__(256) k_fp32_peak(float* out, const int iters, long long* cycles)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f + 1e-9f * blockIdx.x, c = 1e-3f;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
            a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ----------------------------------------------------------------------------------------------------
// host state
// ----------------------------------------------------------------------------------------------------
namespace
{
struct Engine
{
    bool initialised = false;
    int device = -1;
    cudaStream_t ownStream = nullptr;
    cudaStream_t stream = nullptr; // the stream in use (own or caller's)
    int maxW = B200_REF_MAX_BITMAP_WIDTH, maxH = B200_REF_MAX_BITMAP_HEIGHT;
    int frameW = 0, frameH = 0; // size of the last frame rendered
    int rank = 0, world = 1;
    int numSMs = 0, ctasPerSM = 0;
    // scene
    float4* dBoxes = nullptr; int nbBoxes = 0; int nbBoxesIn = 0; int boxLayoutUsed = 0;
    float4* dWide = nullptr; float4* dLeafRecs = nullptr; int nbWide = 0; size_t capWide = 0, capLeafRecs = 0;
    float4* dUWide = nullptr; int nbUWide = 0; size_t capUWide = 0; int opaqueShadows = 0;
    float4* dUWideCH = nullptr; size_t capUWideCH = 0; bool uwDirty = true; // the unit walk's centre / half-extent form of dUWide, refreshed before a frame when dUWide changed
    int nbUX = 0; // point-query tree for backward cylinder hits, appended to dUWide
    float4* dUGroup = nullptr; size_t capUGroup = 0; // the same trees child-major (group walk)
    float4* dGatherScratch = nullptr; size_t capGatherScratch = 0;
    int* dPrimLeaf = nullptr; size_t capPrimLeaf = 0;
    float4* dPrimRecs = nullptr; size_t capPrimRecs = 0; // 96-byte records of the unit walk (trace.cuh), 6 float4 per primitive
    // staged rendering
    float* dPathWords = nullptr; float4* dPathColors = nullptr; float* dPathContrib = nullptr; int* dPathQueues = nullptr;
    float uploadMs = 0.f; int treesOnGpu = 0; // the last b200_h2d_scene
    // device-side animation (animate.cuh): per reference leaf its raw box and its node in the ordered tree, the ordered tree's parent
    // links, the binary node behind every child slot of its 4-wide form, scratch flags; levels of the reference hierarchy
    int* dLeafRaw = nullptr; int* dLeafNode = nullptr; int* dPackedParent = nullptr; int4* dWideKid = nullptr; int* dFitFlags = nullptr;
    size_t capLeafMaps = 0, capPackedParent = 0, capWideKid = 0;
    int nbLeaves = 0, nbPacked = 0, maxBoxLevel = 0, nbLightPrims = 0; bool animatable = false; float animateMs = 0.f;
    int walkTreeLevels = 0;   // levels of the main walk tree (passes a re-fit in place needs)
    float viewDistance = 0.f; // SceneInfo.viewDistance as last seen (the reference resets inner boxes to +-viewDistance before re-fitting them)
    size_t nWideF4 = 0, nLeafRecsF4 = 0;      // float4 in dWide / dLeafRecs (scene replication)
    float4* dLeafBoxes = nullptr; size_t capLeafBoxes = 0; // boxes of the reference leaves, for the GPU tree build
    unsigned int* dQueueCounters = nullptr; size_t pathStride = 0; int pathIterations = 0;
    size_t pathFailedBytes = 0; // smallest path-state size that did not fit (not tried again)
    int ctasPerSMStage[6] = {0, 0, 0, 0, 0, 0}; // k_stage_primary, k_stage_pass, k_stage_reflected, -, -, k_stage_fused
    float4* dGeo = nullptr; int* dMeta = nullptr; b200_Primitive* dPrims = nullptr; int nbPrims = 0;
    b200_BoundingBox* dRawBoxes = nullptr;
    b200_Material* dMats = nullptr; int nbMats = 0;
    b200_LightInformation* dLights = nullptr; int nbLights = 0;
    unsigned char* dTex = nullptr; size_t texBytes = 0;
    float* dRandoms = nullptr;
    size_t capBoxes = 0, capPrims = 0, capMats = 0;
    // host copies needed to (re)build the packed per-primitive word when either side changes
    std::vector<b200_Primitive> hPrims;
    std::vector<b200_Material> hMats;
    // frame
    b200_PostProcessingBuffer* dPost = nullptr; int4* dIds = nullptr; unsigned char* dBitmap = nullptr;
    unsigned char* dPeerBitmap = nullptr; // the root GPU's bitmap mapped into this process (b200_peer_frame_open), else null
    unsigned int* dTileCounter = nullptr; unsigned long long* dWork = nullptr;
    int* dTileOrder = nullptr; size_t capTileOrder = 0; int tileOrderKey[4] = {0, 0, 0, 0}; // tilesX, tilesY, rank, world the table was made for
    size_t pixelsCap = 0;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    bool timed = false;
    unsigned long long launches = 0;
    // pinned registration cache for caller-owned readback buffers
    std::vector<std::pair<void*, size_t>> hostRegistered; // b200_register_host: pinned in place for their owners, never implicitly
    // streamed output: the registered host buffers the last b200_d2h_bitmap filled (host == device until the next frame), the same as
    // the device sees them, whether the frame in flight writes them itself, and the tiles' path counts
    void* mirrorBitmap = nullptr; void* mirrorIds = nullptr;
    unsigned char* mirrorDevBitmap = nullptr; int4* mirrorDevIds = nullptr;
    size_t mirrorPixels = 0;
    bool mirrorExplicit = false;                  // the buffers were named by b200_stream_target: every frame goes there, on every rank
    bool mirrorArmed = false;                     // a b200_d2h_bitmap since the last frame: the reader reads every frame
    bool validBitmap = false, validIds = false;   // the host buffer holds what the device buffer holds (once the stream is idle)
    bool streamedBitmap = false, streamedIds = false; // the frame in flight writes them
    int* dTileRemaining = nullptr; size_t capTileRemaining = 0;
    unsigned long long framesStreamed = 0;
    int err = 0;
    char errMsg[256] = {0};
};
Engine G;

void latch(int code, const char* what, const char* detail)
{
    fprintf(stderr, "[solr_b200] ERROR %s: %s (%d)\n", what, detail ? detail : "", code);
    if (G.err == 0)
    {
        G.err = code;
        snprintf(G.errMsg, sizeof(G.errMsg), "%s: %s", what, detail ? detail : "");
    }
}
#define CK(call)                                                                       \
    do                                                                                 \
    {                                                                                  \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) latch((int)e_, #call, cudaGetErrorString(e_));          \
    } while (0)

template <typename T>
void freeDev(T*& p)
{
    if (p) { CK(cudaFree(p)); p = nullptr; }
}

bool ensureDevice()
{
    if (G.device < 0)
    {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess) { latch(-1, "ensureDevice", "no CUDA device: this engine has no CPU fallback"); return false; }
        G.device = d;
    }
    cudaError_t e = cudaSetDevice(G.device);
    if (e != cudaSuccess) { latch((int)e, "cudaSetDevice", cudaGetErrorString(e)); return false; }
    return true;
}

float intBits(int v)
{
    float f;
    memcpy(&f, &v, sizeof(f));
    return f;
}

int packMeta(const b200_Primitive& p, const std::vector<b200_Material>& mats)
{
    int fast = 0, procedural = 0, transparent = 0;
    if (p.materialId >= 0 && (size_t)p.materialId < mats.size())
    {
        const b200_Material& m = mats[p.materialId];
        fast = (m.attributes.x == 0) ? 0 : (m.attributes.x == 1 ? 1 : 2);
        procedural = m.attributes.y != 0;
        transparent = m.transparency != 0.f;
    }
    return (p.type & 0xF) | (fast << 4) | (procedural << 6) | (transparent << 7) | (p.materialId << 8);
}

void uploadMeta()
{
    if (G.hPrims.empty() || !G.dMeta) return;
    std::vector<int> meta(G.hPrims.size());
    int opaque = 1;
    for (size_t i = 0; i < G.hPrims.size(); ++i)
    {
        meta[i] = packMeta(G.hPrims[i], G.hMats);
        // shadow casters that do not block fully: transparent materials (GeometryIntersections.cuh:280-281,880-892) and
        // planes whose opacity comes from a texel (:551-559)
        const b200_Primitive& p = G.hPrims[i];
        const bool planeClass = p.type == B200_PT_CHECKBOARD || p.type == B200_PT_CAMERA || p.type == B200_PT_XYPLANE ||
                                p.type == B200_PT_YZPLANE || p.type == B200_PT_XZPLANE || p.type == B200_PT_MAGICCARPET || p.type == B200_PT_QUAD;
        bool textured = false;
        if (p.materialId >= 0 && (size_t)p.materialId < G.hMats.size()) textured = G.hMats[p.materialId].textureIds.x != B200_TEXTURE_NONE;
        if (((meta[i] >> 7) & 1) || (planeClass && (textured || p.type == B200_PT_CAMERA))) opaque = 0;
    }
    G.opaqueShadows = opaque;
    CK(cudaMemcpyAsync(G.dMeta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, G.stream));
    // the unit walk reads the packed word from the primitive's record ((n1, packed word) is its fourth float4)
    if (G.dPrimRecs && G.capPrimRecs >= PRIM_REC_F4 * meta.size())
        CK(cudaMemcpy2DAsync(reinterpret_cast<char*>(G.dPrimRecs) + 3 * sizeof(float4) + 3 * sizeof(float), PRIM_REC_F4 * sizeof(float4), meta.data(),
                             sizeof(int), sizeof(int), meta.size(), cudaMemcpyHostToDevice, G.stream));
    CK(cudaStreamSynchronize(G.stream)); // meta is a stack-lifetime staging vector
}

void dropMirror()
{
    G.mirrorBitmap = G.mirrorIds = nullptr; G.mirrorDevBitmap = nullptr; G.mirrorDevIds = nullptr;
    G.mirrorArmed = G.mirrorExplicit = G.validBitmap = G.validIds = G.streamedBitmap = G.streamedIds = false;
}

void unregisterHost()
{
    if (G.streamedBitmap || G.streamedIds) cudaStreamSynchronize(G.stream); // a frame in flight may be writing them
    dropMirror();
    // owners unregister their buffers before freeing them; whatever is still listed is released here (an owner that freed first
    // makes cudaHostUnregister fail, which must not surface later as a launch error)
    for (auto& r : G.hostRegistered) if (cudaHostUnregister(r.first) != cudaSuccess) cudaGetLastError();
    G.hostRegistered.clear();
    cudaGetLastError();
}
} // namespace


// Literal box re-layout (fallback): the reference's own hierarchy with single-child chains collapsed.
static int relayoutBoxesLiteral(const b200_BoundingBox* boxes, int nbBoxes, std::vector<float4>& packed)
{
    // 1. survivors of the chain collapse
    std::vector<unsigned char> keep(nbBoxes, 1);
    for (int i = 0; i < nbBoxes; ++i)
    {
        const b200_BoundingBox& b = boxes[i];
        const int skip = b.indexForNextBox.x;
        if (b.nbPrimitives != 0 || skip <= 1 || i + 1 >= nbBoxes || i + skip > nbBoxes) continue;
        const b200_BoundingBox& c = boxes[i + 1];
        if (c.indexForNextBox.x != skip - 1) continue; // more than one child
        const bool contained = b.parameters[0].x <= c.parameters[0].x && b.parameters[0].y <= c.parameters[0].y &&
                               b.parameters[0].z <= c.parameters[0].z && b.parameters[1].x >= c.parameters[1].x &&
                               b.parameters[1].y >= c.parameters[1].y && b.parameters[1].z >= c.parameters[1].z;
        if (contained) keep[i] = 0;
    }
    // 2. new positions; a leaf whose reference skip is not 1 (lights box with children, GPUKernel.cpp:1248-1251)
    //    is emitted as an inner box with identical bounds followed by the leaf.
    std::vector<int> pos(nbBoxes + 1, 0);
    for (int i = 0; i < nbBoxes; ++i)
    {
        int n = keep[i] ? 1 : 0;
        if (keep[i] && boxes[i].nbPrimitives > 0 && boxes[i].indexForNextBox.x != 1) n = 2;
        pos[i + 1] = pos[i] + n;
    }
    const int nOut = pos[nbBoxes];
    packed.assign(2 * (size_t)nOut, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int i = 0; i < nbBoxes; ++i)
    {
        if (!keep[i]) continue;
        const b200_BoundingBox& b = boxes[i];
        int end = i + b.indexForNextBox.x;
        if (end > nbBoxes) end = nbBoxes;
        if (end <= i) end = i + 1; // a zero/negative skip would never terminate in the reference; advance instead
        int o = pos[i];
        const float4 lo = make_float4(b.parameters[0].x, b.parameters[0].y, b.parameters[0].z, 0.f);
        const float4 hi = make_float4(b.parameters[1].x, b.parameters[1].y, b.parameters[1].z, 0.f);
        auto put = [&](int at, int w0, int w1) {
            packed[2 * (size_t)at] = lo; packed[2 * (size_t)at + 1] = hi;
            packed[2 * (size_t)at].w = intBits(w0); packed[2 * (size_t)at + 1].w = intBits(w1);
        };
        if (b.nbPrimitives > 0)
        {
            if (b.indexForNextBox.x != 1) { put(o, pos[end] - o, 0); ++o; }
            put(o, b.startIndex, b.nbPrimitives);
        }
        else
            put(o, pos[end] - o, 0);
    }

    return nOut;
}


// ----------------------------------------------------------------------------------------------------
// Ordered BVH over the reference's leaves.
//
// Only the LEAVES of the reference's flattened hierarchy are observable: a ray tests a leaf's primitives iff
// the leaf box and every ancestor pass the slab test, each against the closest distance known when it is
// reached (GeometryIntersections.cuh:687-770).  Ancestors contain their leaves (GPUKernel.cpp:841-892 builds
// them as min/max unions), the slab arithmetic is monotone in the corner coordinates, and the closest distance
// only shrinks along a walk — so an ancestor can never reject a ray its leaf would accept: the reference
// visits exactly {leaves whose OWN box passes}, in array order.  Any hierarchy over the same leaf sequence
// whose inner boxes contain their leaves therefore reproduces the reference's walk bit for bit, and we are
// free to build a good one: a binary tree over contiguous runs of the (spatially coherent, depth-first grid
// order) leaf sequence, split by the surface-area heuristic, laid out depth-first with skip counts so the
// device walk stays a stackless list walk.  The reference's own levels are uniform grids of N/4^k cells
// whose top levels have hundreds of siblings that every ray must scan.
// Containment of every leaf in all of its reference ancestors is verified; if it fails the literal layout
// is used instead.
// ----------------------------------------------------------------------------------------------------
namespace
{
struct Aabb
{
    float lo[3], hi[3];
    void grow(const Aabb& o)
    {
        for (int k = 0; k < 3; ++k) { lo[k] = o.lo[k] < lo[k] ? o.lo[k] : lo[k]; hi[k] = o.hi[k] > hi[k] ? o.hi[k] : hi[k]; }
    }
    double area() const
    {
        const double x = (double)hi[0] - lo[0], y = (double)hi[1] - lo[1], z = (double)hi[2] - lo[2];
        return (x < 0 || y < 0 || z < 0) ? 0.0 : 2.0 * (x * y + y * z + z * x);
    }
    bool contains(const Aabb& o) const
    {
        for (int k = 0; k < 3; ++k) if (!(lo[k] <= o.lo[k] && hi[k] >= o.hi[k])) return false;
        return true;
    }
};
Aabb aabbOf(const b200_BoundingBox& b)
{
    Aabb a;
    a.lo[0] = b.parameters[0].x; a.lo[1] = b.parameters[0].y; a.lo[2] = b.parameters[0].z;
    a.hi[0] = b.parameters[1].x; a.hi[1] = b.parameters[1].y; a.hi[2] = b.parameters[1].z;
    return a;
}
struct LeafRec { Aabb box; int start, count; };

struct BvhBuilder
{
    const std::vector<LeafRec>& leaves;
    std::vector<float4>& out;
    std::vector<Aabb> suffix;
    BvhBuilder(const std::vector<LeafRec>& l, std::vector<float4>& o) : leaves(l), out(o), suffix(l.size() + 1) {}

    int emit(const Aabb& b, int w0, int w1)
    {
        const int at = (int)(out.size() / 2);
        out.push_back(make_float4(b.lo[0], b.lo[1], b.lo[2], intBits(w0)));
        out.push_back(make_float4(b.hi[0], b.hi[1], b.hi[2], intBits(w1)));
        return at;
    }
    // nodes for leaves [i, j), depth-first; returns the range's bounds
    Aabb build(int i, int j, int depth)
    {
        if (j - i == 1)
        {
            emit(leaves[i].box, leaves[i].start, leaves[i].count);
            return leaves[i].box;
        }
        // surface-area heuristic over the j-i-1 contiguous splits
        int split = (i + j) / 2;
        Aabb all = leaves[i].box;
        for (int k = i + 1; k < j; ++k) all.grow(leaves[k].box);
        if (depth < 56 && j - i > 2)
        {
            suffix[j - 1] = leaves[j - 1].box;
            for (int k = j - 2; k > i; --k) { suffix[k] = leaves[k].box; suffix[k].grow(suffix[k + 1]); }
            Aabb prefix = leaves[i].box;
            double best = 1e300;
            for (int k = i + 1; k < j; ++k) // left = [i, k), right = [k, j)
            {
                const double cost = prefix.area() * (k - i) + suffix[k].area() * (j - k);
                if (cost < best) { best = cost; split = k; }
                prefix.grow(leaves[k].box);
            }
        }
        const int at = emit(all, 0, 0);
        build(i, split, depth + 1);
        build(split, j, depth + 1);
        out[2 * (size_t)at].w = intBits((int)(out.size() / 2) - at); // skip = subtree size
        return all;
    }
};

// returns false if some leaf is not contained in one of its reference ancestors
bool collectLeaves(const b200_BoundingBox* boxes, int nbBoxes, std::vector<LeafRec>& leaves)
{
    struct Anc { int end; Aabb box; };
    std::vector<Anc> stack;
    bool ok = true;
    for (int i = 0; i < nbBoxes; ++i)
    {
        while (!stack.empty() && stack.back().end <= i) stack.pop_back();
        const b200_BoundingBox& b = boxes[i];
        const Aabb box = aabbOf(b);
        int skip = b.indexForNextBox.x;
        if (skip < 1) skip = 1;
        if (b.nbPrimitives > 0)
        {
            for (const Anc& a : stack) if (!a.box.contains(box)) ok = false;
            LeafRec l; l.box = box; l.start = b.startIndex; l.count = b.nbPrimitives;
            leaves.push_back(l);
        }
        if (skip > 1) { Anc a; a.end = i + skip; a.box = box; stack.push_back(a); }
    }
    return ok;
}

// Wide collapse of the binary list (depth-first, skip counts): a node adopts its grandchildren, largest surface area
// first and keeping array order, until it has `width` children (4, or 8 for the unordered trees).  A node is width/4
// consecutive 128-byte records of four children each: lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] refs[4] (pad).
struct WideBuilder
{
    const std::vector<float4>& bin;  // 2 float4 per binary node
    std::vector<float4>& wide;       // 8 float4 per record
    std::vector<int> leafOrdinal;    // binary node index -> leaf ordinal
    std::vector<int>* kidNodes = nullptr; // optional: binary node of every child slot of every record (-1: empty), 4 per record
    int width;
    WideBuilder(const std::vector<float4>& b, std::vector<float4>& w, int wd) : bin(b), wide(w), width(wd) {}
    int w0(int i) const { int v; memcpy(&v, &bin[2 * (size_t)i].w, 4); return v; }
    int w1(int i) const { int v; memcpy(&v, &bin[2 * (size_t)i + 1].w, 4); return v; }
    bool isLeaf(int i) const { return w1(i) > 0; }
    int size(int i) const { return isLeaf(i) ? 1 : w0(i); }
    double area(int i) const
    {
        const float4 lo = bin[2 * (size_t)i], hi = bin[2 * (size_t)i + 1];
        const double x = (double)hi.x - lo.x, y = (double)hi.y - lo.y, z = (double)hi.z - lo.z;
        return (x < 0 || y < 0 || z < 0) ? 0.0 : 2.0 * (x * y + y * z + z * x);
    }
    void writeNode(int at, const int* kids, const int* refs, int n)
    {
        const int recs = width / 4;
        for (int q = 0; q < recs; ++q)
        {
            float rows[6][4];
            int rf[4];
            for (int k = 0; k < 4; ++k)
            {
                const int c = 4 * q + k;
                if (c < n)
                {
                    const float4 lo = bin[2 * (size_t)kids[c]], hi = bin[2 * (size_t)kids[c] + 1];
                    rows[0][k] = lo.x; rows[1][k] = lo.y; rows[2][k] = lo.z; rows[3][k] = hi.x; rows[4][k] = hi.y; rows[5][k] = hi.z;
                    rf[k] = refs[c];
                }
                else
                {
                    rows[0][k] = rows[1][k] = rows[2][k] = 3.0e38f; rows[3][k] = rows[4][k] = rows[5][k] = -3.0e38f;
                    rf[k] = (int)0x80000000;
                }
            }
            float4* rec = &wide[8 * ((size_t)at * recs + q)];
            if (kidNodes)
            {
                if (kidNodes->size() < 4 * ((size_t)at * recs + q + 1)) kidNodes->resize(4 * ((size_t)at * recs + q + 1), -1);
                for (int k = 0; k < 4; ++k) (*kidNodes)[4 * ((size_t)at * recs + q) + k] = (4 * q + k < n) ? kids[4 * q + k] : -1;
            }
            for (int r = 0; r < 6; ++r) rec[r] = make_float4(rows[r][0], rows[r][1], rows[r][2], rows[r][3]);
            rec[6] = make_float4(intBits(rf[0]), intBits(rf[1]), intBits(rf[2]), intBits(rf[3]));
            rec[7] = make_float4(intBits(n), 0.f, 0.f, 0.f);
        }
    }
    int maxDepth = 0; // deepest wide node (root = 1): a walk's stack holds at most 3 entries per level
    int build(int i, int depth = 1) // i: inner binary node
    {
        if (depth > maxDepth) maxDepth = depth;
        int kids[8], n = 2;
        kids[0] = i + 1;
        kids[1] = i + 1 + size(i + 1);
        while (n < width)
        {
            int best = -1;
            double bestArea = -1.0;
            for (int k = 0; k < n; ++k)
                if (!isLeaf(kids[k]) && area(kids[k]) > bestArea) { bestArea = area(kids[k]); best = k; }
            if (best < 0) break;
            const int c = kids[best];
            for (int k = n; k > best + 1; --k) kids[k] = kids[k - 1];
            kids[best] = c + 1;
            kids[best + 1] = c + 1 + size(c + 1);
            ++n;
        }
        const int recs = width / 4;
        const int at = (int)(wide.size() / (8 * (size_t)recs));
        wide.resize(wide.size() + 8 * (size_t)recs);
        int refs[8];
        for (int k = 0; k < n; ++k) refs[k] = isLeaf(kids[k]) ? ~leafOrdinal[kids[k]] : build(kids[k], depth + 1);
        writeNode(at, kids, refs, n);
        return at;
    }
};

// wide nodes + leaf records from the binary list; returns the number of wide nodes
int buildWide(const std::vector<float4>& bin, std::vector<float4>& wide, std::vector<float4>& leafRecs,
              const std::vector<int>* leafOfNode = nullptr, int width = 4, std::vector<int>* kidNodes = nullptr, std::vector<int>* leafNodes = nullptr,
              int* maxDepthOut = nullptr)
{
    if (maxDepthOut) *maxDepthOut = 1;
    wide.clear();
    const int nb = (int)(bin.size() / 2);
    if (nb == 0) return 0;
    WideBuilder b(bin, wide, width);
    b.kidNodes = kidNodes;
    if (kidNodes) kidNodes->clear();
    if (leafNodes) leafNodes->clear();
    b.leafOrdinal.assign(nb, -1);
    if (leafOfNode)
    {
        // leaf ordinals are given (unordered tree over the ordered tree's leaf records)
        for (int i = 0; i < nb; ++i) b.leafOrdinal[i] = (*leafOfNode)[i];
    }
    else
    {
        leafRecs.clear();
        for (int i = 0; i < nb; ++i)
            if (b.isLeaf(i))
            {
                b.leafOrdinal[i] = (int)(leafRecs.size() / 2);
                if (leafNodes) leafNodes->push_back(i);
                leafRecs.push_back(bin[2 * (size_t)i]);
                leafRecs.push_back(bin[2 * (size_t)i + 1]);
            }
    }
    const int recs = width / 4;
    if (b.isLeaf(0))
    {
        // a single leaf: one wide node with one child
        wide.resize(8 * (size_t)recs);
        const int kid = 0, ref = ~b.leafOrdinal[0];
        b.writeNode(0, &kid, &ref, 1);
        return 1;
    }
    b.build(0);
    if (maxDepthOut) *maxDepthOut = b.maxDepth;
    return (int)(wide.size() / (8 * (size_t)recs));
}

// ----------------------------------------------------------------------------------------------------
// Unordered BVH over the same leaves (binned SAH on leaf-box centroids, no ordering constraint), emitted in the
// same depth-first binary list format so buildWide() can collapse it.  Used by the walks whose result does not
// depend on the visiting order (rays with |direction| >= 1, see trace.cuh "order-independent walks").
// ----------------------------------------------------------------------------------------------------
struct SahBuilder
{
    const std::vector<LeafRec>& leaves;
    std::vector<int> order;          // permutation of leaf ordinals, partitioned in place
    std::vector<float4>& out;        // binary list: leaf nodes carry (first primitive, count) in (w0, w1), inner nodes the skip count
    std::vector<int>& leafOfNode;    // per emitted node: leaf ordinal or -1
    SahBuilder(const std::vector<LeafRec>& l, std::vector<float4>& o, std::vector<int>& lon) : leaves(l), out(o), leafOfNode(lon)
    {
        order.resize(l.size());
        for (size_t i = 0; i < l.size(); ++i) order[i] = (int)i;
    }
    static void emit(std::vector<float4>& o, std::vector<int>& lon, const Aabb& b, int w0, int w1, int leaf)
    {
        o.push_back(make_float4(b.lo[0], b.lo[1], b.lo[2], intBits(w0)));
        o.push_back(make_float4(b.hi[0], b.hi[1], b.hi[2], intBits(w1)));
        lon.push_back(leaf);
    }
    // bounds of [i, j) and the split position (binned SAH on box centroids, 16 bins per axis; median when no split helps)
    int split(int i, int j, int depth, Aabb& all)
    {
        all = leaves[order[i]].box;
        Aabb cb;
        for (int k = 0; k < 3; ++k) cb.lo[k] = 3e38f, cb.hi[k] = -3e38f;
        for (int k = i; k < j; ++k)
        {
            const Aabb& b = leaves[order[k]].box;
            all.grow(b);
            for (int a = 0; a < 3; ++a)
            {
                const float c = 0.5f * (b.lo[a] + b.hi[a]);
                cb.lo[a] = c < cb.lo[a] ? c : cb.lo[a]; cb.hi[a] = c > cb.hi[a] ? c : cb.hi[a];
            }
        }
        int mid = (i + j) / 2;
#ifndef SAH_BINS
#define SAH_BINS 16
#endif
        const int NB = SAH_BINS;
        double bestCost = 1e300; int bestAxis = -1, bestBin = -1;
        if (depth < 60)
            for (int a = 0; a < 3; ++a)
            {
                const float ext = cb.hi[a] - cb.lo[a];
                if (!(ext > 0.f)) continue;
                Aabb bb[NB]; int cnt[NB];
                for (int b = 0; b < NB; ++b) { cnt[b] = 0; for (int k = 0; k < 3; ++k) bb[b].lo[k] = 3e38f, bb[b].hi[k] = -3e38f; }
                for (int k = i; k < j; ++k)
                {
                    const Aabb& b = leaves[order[k]].box;
                    int bin = (int)(NB * ((0.5f * (b.lo[a] + b.hi[a]) - cb.lo[a]) / ext));
                    bin = bin < 0 ? 0 : (bin >= NB ? NB - 1 : bin);
                    cnt[bin]++; bb[bin].grow(b);
                }
                Aabb right[NB]; int rc[NB];
                Aabb acc; for (int k = 0; k < 3; ++k) acc.lo[k] = 3e38f, acc.hi[k] = -3e38f;
                int c = 0;
                for (int b = NB - 1; b > 0; --b) { if (cnt[b]) acc.grow(bb[b]); c += cnt[b]; right[b] = acc; rc[b] = c; }
                for (int k = 0; k < 3; ++k) acc.lo[k] = 3e38f, acc.hi[k] = -3e38f;
                c = 0;
                for (int b = 0; b < NB - 1; ++b)
                {
                    if (cnt[b]) acc.grow(bb[b]);
                    c += cnt[b];
                    if (c == 0 || rc[b + 1] == 0) continue;
                    const double cost = acc.area() * c + right[b + 1].area() * rc[b + 1];
                    if (cost < bestCost) { bestCost = cost; bestAxis = a; bestBin = b; }
                }
            }
        if (bestAxis >= 0)
        {
            const int a = bestAxis;
            const float ext = cb.hi[a] - cb.lo[a];
            auto binOf = [&](int leaf) {
                const Aabb& b = leaves[leaf].box;
                int bin = (int)(NB * ((0.5f * (b.lo[a] + b.hi[a]) - cb.lo[a]) / ext));
                return bin < 0 ? 0 : (bin >= NB ? NB - 1 : bin);
            };
            int l = i, r = j - 1;
            while (l <= r)
            {
                if (binOf(order[l]) <= bestBin) ++l;
                else { std::swap(order[l], order[r]); --r; }
            }
            if (l > i && l < j) mid = l;
        }
        return mid;
    }
    // depth-first list of the subtree over order[i, j) appended to (o, lon); skip counts are relative, so lists concatenate
    void buildInto(std::vector<float4>& o, std::vector<int>& lon, int i, int j, int depth)
    {
        if (j - i == 1)
        {
            const LeafRec& l = leaves[order[i]];
            emit(o, lon, l.box, l.start, l.count, order[i]);
            return;
        }
        Aabb all;
        const int mid = split(i, j, depth, all);
        const size_t at = o.size() / 2;
        emit(o, lon, all, 0, 0, -1);
        if (depth < 4 && j - i > 8192)
        {
            // the two halves are independent (disjoint ranges of `order`): build the left one on another thread
            std::vector<float4> lo_; std::vector<int> ll_;
            std::thread worker([&]() { buildInto(lo_, ll_, i, mid, depth + 1); });
            std::vector<float4> ro_; std::vector<int> rl_;
            buildInto(ro_, rl_, mid, j, depth + 1);
            worker.join();
            o.insert(o.end(), lo_.begin(), lo_.end()); lon.insert(lon.end(), ll_.begin(), ll_.end());
            o.insert(o.end(), ro_.begin(), ro_.end()); lon.insert(lon.end(), rl_.begin(), rl_.end());
        }
        else
        {
            buildInto(o, lon, i, mid, depth + 1);
            buildInto(o, lon, mid, j, depth + 1);
        }
        o[2 * at].w = intBits((int)(o.size() / 2 - at));
    }
    void build(int i, int j, int depth) { buildInto(out, leafOfNode, i, j, depth); }
};

int g_useWide = 1;
int g_useUnordered = 1;
int g_tileOrder = 0; // order in which a GPU's own tiles are handed out: 0 row-major, 1 along a Z-order curve (neighbouring warps work on neighbouring tiles in both directions)
int g_fuseTailPercent = 300; // k_stage_pass(p) carries its paths to the end in registers when the queue holds at most p times this share of the resident lanes (0: never)
int g_useStaged = 1;   // 0: always the single kernel; 1: one launch per pass over compacted path queues where the camera allows it; 2: every pass in one persistent launch
int g_useBackward = 1; // point query for hits behind the origin (cylinders/cones); 0 drops that reference behaviour from the order-independent walks
int g_packetMask = 0x0; // per-lane wide walks with deferred leaves beat packets once the code working set is small (profiles/r01_history.md) // bit0 primary, bit1 secondary, bit2 shadow of primary hits, bit3 other shadow walks as packets
int g_streamOutput = 1; // frames whose reader reads every frame into pinned buffers are written there by the ray kernels (engine.cu "streamed output")
int g_animateRefit = 1; // device-side animation: 1 re-fits the main walk tree in place (default), 0 rebuilds it (linear BVH)
int g_gpuTrees = 0; // 1: the trees of the order-independent walks are built on the GPU (treebuild.cuh) instead of on host threads
int g_boxLayout = 0; // 0 auto (ordered BVH when provably equivalent), 1 literal, 2 ordered BVH (unchecked)
} // namespace

// Box re-layout shared by h2d_scene and the host-only debug entry point (tests check it without a GPU).
static int relayoutBoxes(const b200_BoundingBox* boxes, int nbBoxes, std::vector<float4>& packed, int* layoutUsed = nullptr,
                         std::vector<LeafRec>* leavesOut = nullptr)
{
    if (g_boxLayout != 1)
    {
        std::vector<LeafRec> leavesLocal;
        std::vector<LeafRec>& leaves = leavesOut ? *leavesOut : leavesLocal;
        leaves.clear();
        const bool contained = collectLeaves(boxes, nbBoxes, leaves);
        if ((contained || g_boxLayout == 2) && !leaves.empty())
        {
            packed.clear();
            packed.reserve(4 * leaves.size());
            BvhBuilder builder(leaves, packed);
            builder.build(0, (int)leaves.size(), 0);
            if (layoutUsed) *layoutUsed = 2;
            return (int)(packed.size() / 2);
        }
    }
    if (layoutUsed) *layoutUsed = 1;
    return relayoutBoxesLiteral(boxes, nbBoxes, packed);
}

// ----------------------------------------------------------------------------------------------------
// the seam
// ----------------------------------------------------------------------------------------------------
#ifndef UW_PAD
#define UW_PAD 0.02f // padding of the walk trees' primitive boxes (buildWalkTrees, treebuild.cuh)
#endif
#include "treebuild.cuh"
#include "animate.cuh"

extern "C" {

void b200_set_device(int device) { G.device = device; }
void b200_set_stream(void* s)
{
    // work already queued (buffer clears, uploads) must be visible to whatever runs on the new stream
    if (G.stream && ensureDevice()) CK(cudaStreamSynchronize(G.stream));
    G.stream = s ? (cudaStream_t)s : G.ownStream;
}
void b200_set_limits(int w, int h) { if (w > 0 && h > 0) { G.maxW = w; G.maxH = h; } }
void b200_set_option(int key, int value)
{
    if (key == 1 && value >= 0 && value <= 2) g_boxLayout = value;
    else if (key == 2) g_packetMask = value & 0xF;
    else if (key == 3) g_useWide = value != 0;
    else if (key == 4) g_useUnordered = value != 0;
    else if (key == 5) g_useBackward = value != 0;
    else if (key == 10 && (value == 0 || value == 1)) g_gpuTrees = value;
    else if (key == 11 && (value == 0 || value == 1)) g_animateRefit = value;
    else if (key == 12 && (value == 0 || value == 1)) g_streamOutput = value;
    else if (key == 6 && value >= 0 && value <= 2) g_useStaged = value; // 2: fused stages (k_stage_fused)
    else if (key == 8 && value >= 0) g_fuseTailPercent = value;
    else if (key == 9 && (value == 0 || value == 1)) g_tileOrder = value;
    else latch(-11, "b200_set_option", "unknown option");
}
void b200_set_partition(int rank, int world)
{
    if (world < 1 || rank < 0 || rank >= world) { latch(-2, "b200_set_partition", "rank/world out of range"); return; }
    G.rank = rank; G.world = world;
}
// The multi-GPU exchange step fused into the ray kernels: the root's device bitmap is mapped into every other process
// (CUDA IPC; peer access over NVLink / NVSwitch) and the kernel that ends a path packs its RGB8 straight into the root's
// frame, so the frame is complete on the root when the last rank's kernels are — no partial bitmaps, no reduce.
static void closePeerFrame()
{
    if (G.dPeerBitmap) { cudaIpcCloseMemHandle(G.dPeerBitmap); cudaGetLastError(); G.dPeerBitmap = nullptr; }
}
int b200_peer_frame_export(void* handle, int handleBytes)
{
    if (!handle || handleBytes < (int)sizeof(cudaIpcMemHandle_t)) { latch(-12, "b200_peer_frame_export", "handle buffer too small (64 bytes)"); return -12; }
    if (!G.dBitmap) { latch(-4, "b200_peer_frame_export", "reshape_scene not called"); return -4; }
    if (!ensureDevice()) return G.err;
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, G.dBitmap);
    if (e != cudaSuccess) { cudaGetLastError(); latch((int)e, "b200_peer_frame_export", cudaGetErrorString(e)); return (int)e; }
    memcpy(handle, &h, sizeof(h));
    return 0;
}
int b200_peer_frame_open(const void* handle, int handleBytes)
{
    if (!ensureDevice()) return G.err;
    CK(cudaStreamSynchronize(G.stream));
    closePeerFrame();
    if (!handle) return 0;
    if (handleBytes < (int)sizeof(cudaIpcMemHandle_t)) { latch(-12, "b200_peer_frame_open", "handle too small (64 bytes)"); return -12; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); latch((int)e, "b200_peer_frame_open", cudaGetErrorString(e)); return (int)e; }
    G.dPeerBitmap = (unsigned char*)p;
    return 0;
}
int b200_last_error(char* msg, int cap)
{
    if (msg && cap > 0) { strncpy(msg, G.errMsg, cap - 1); msg[cap - 1] = 0; }
    return G.err;
}
void b200_clear_error(void) { G.err = 0; G.errMsg[0] = 0; }

void b200_initialize_scene(b200_int2 occ, b200_SceneInfo si, int, int, int)
{
    G.viewDistance = si.viewDistance;
    if (occ.x != 1) latch(-3, "b200_initialize_scene", "one process drives one GPU: occupancyParameters.x must be 1 (use b200_set_partition)");
    if (!ensureDevice()) return;
    if (G.initialised) return;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, G.device));
    G.numSMs = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&G.ownStream, cudaStreamNonBlocking));
    if (!G.stream) G.stream = G.ownStream;
    CK(cudaEventCreate(&G.evStart));
    CK(cudaEventCreate(&G.evStop));
    CK(cudaMalloc(&G.dTileCounter, sizeof(unsigned int)));
    CK(cudaMalloc(&G.dWork, 8 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(G.dWork, 0, 8 * sizeof(unsigned long long), G.stream));
    CK(cudaMalloc(&G.dLights, B200_NB_MAX_LIGHTINFORMATIONS * sizeof(b200_LightInformation)));
    CK(cudaMemsetAsync(G.dLights, 0, B200_NB_MAX_LIGHTINFORMATIONS * sizeof(b200_LightInformation), G.stream));
    int perSM = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_render, CTA_THREADS, 0));
    G.ctasPerSM = perSM > 0 ? perSM : 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_stage_primary<false>, CTA_THREADS, 0)); G.ctasPerSMStage[0] = perSM > 0 ? perSM : 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_stage_pass<false>, CTA_THREADS, 0)); G.ctasPerSMStage[1] = perSM > 0 ? perSM : 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_stage_reflected<false>, CTA_THREADS, 0)); G.ctasPerSMStage[2] = perSM > 0 ? perSM : 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_stage_fused, CTA_THREADS, 0)); G.ctasPerSMStage[5] = perSM > 0 ? perSM : 1;
    G.initialised = true;
    G.launches = 0;
}

void b200_finalize_scene(b200_int2)
{
    if (!G.initialised) return;
    if (!ensureDevice()) return;
    cudaDeviceSynchronize();
    unregisterHost();
    closePeerFrame();
    freeDev(G.dWide); freeDev(G.dLeafRecs); G.capWide = G.capLeafRecs = 0; G.nbWide = 0;
    freeDev(G.dUWide); G.capUWide = 0; G.nbUWide = 0; G.nbUX = 0; freeDev(G.dUWideCH); G.capUWideCH = 0; G.uwDirty = true; freeDev(G.dLeafBoxes); G.capLeafBoxes = 0; treebuild::releaseScratch();
    freeDev(G.dLeafRaw); freeDev(G.dLeafNode); freeDev(G.dPackedParent); freeDev(G.dFitFlags); freeDev(G.dWideKid); G.capLeafMaps = G.capPackedParent = G.capWideKid = 0; G.animatable = false; freeDev(G.dPrimLeaf); G.capPrimLeaf = 0; freeDev(G.dPrimRecs); G.capPrimRecs = 0;
    freeDev(G.dUGroup); G.capUGroup = 0; freeDev(G.dGatherScratch); G.capGatherScratch = 0;
    freeDev(G.dBoxes); freeDev(G.dGeo); freeDev(G.dMeta); freeDev(G.dPrims); freeDev(G.dRawBoxes); freeDev(G.dMats);
    freeDev(G.dLights); freeDev(G.dTex); freeDev(G.dRandoms); freeDev(G.dPost); freeDev(G.dIds); freeDev(G.dBitmap);
    freeDev(G.dTileRemaining); G.capTileRemaining = 0;
    freeDev(G.dTileCounter); freeDev(G.dWork); freeDev(G.dTileOrder); G.capTileOrder = 0; G.tileOrderKey[0] = 0;
    freeDev(G.dPathWords); freeDev(G.dPathColors); freeDev(G.dPathContrib); freeDev(G.dPathQueues); freeDev(G.dQueueCounters);
    G.pathStride = 0; G.pathIterations = 0;
    if (G.evStart) { cudaEventDestroy(G.evStart); G.evStart = nullptr; }
    if (G.evStop) { cudaEventDestroy(G.evStop); G.evStop = nullptr; }
    if (G.stream == G.ownStream) G.stream = nullptr;
    if (G.ownStream) { cudaStreamDestroy(G.ownStream); G.ownStream = nullptr; }
    G.capBoxes = G.capPrims = G.capMats = 0; G.pixelsCap = 0; G.texBytes = 0;
    G.nbBoxes = G.nbPrims = G.nbMats = G.nbLights = 0;
    G.hPrims.clear(); G.hMats.clear();
    G.initialised = false; G.timed = false;
    // no cudaDeviceReset(): the process may share the device with NCCL / PyTorch
}

void b200_reshape_scene(b200_int2, b200_SceneInfo si)
{
    G.viewDistance = si.viewDistance;
    if (!ensureDevice()) return;
    const size_t px = (size_t)G.maxW * (size_t)G.maxH;
    freeDev(G.dRandoms); freeDev(G.dPost); freeDev(G.dIds); freeDev(G.dBitmap);
    unregisterHost();
    CK(cudaMalloc(&G.dRandoms, (px + 4) * sizeof(float)));
    CK(cudaMemsetAsync(G.dRandoms, 0, (px + 4) * sizeof(float), G.stream));
    CK(cudaMalloc(&G.dPost, px * sizeof(b200_PostProcessingBuffer)));
    CK(cudaMemsetAsync(G.dPost, 0, px * sizeof(b200_PostProcessingBuffer), G.stream));
    CK(cudaMalloc(&G.dIds, px * sizeof(int4)));
    CK(cudaMemsetAsync(G.dIds, 0, px * sizeof(int4), G.stream));
    CK(cudaMalloc(&G.dBitmap, px * B200_COLOR_DEPTH));
    CK(cudaMemsetAsync(G.dBitmap, 0, px * B200_COLOR_DEPTH, G.stream));
    CK(cudaStreamSynchronize(G.stream));
    G.pixelsCap = px;
}

// Renumbers the nodes of a wide tree breadth-first (node 0 stays the root), so that the top levels — visited by every
// ray — are the first records of the array (the walks keep them in shared memory) and siblings are neighbours.
static void renumberBreadthFirst(std::vector<float4>& wide, int nbNodes)
{
    const int recs = UW_WIDTH / 4;
    const size_t f4 = 8 * (size_t)recs;
    if (nbNodes <= 1) return;
    std::vector<int> order, newIndex(nbNodes, -1);
    order.reserve(nbNodes);
    order.push_back(0); newIndex[0] = 0;
    for (size_t head = 0; head < order.size(); ++head)
    {
        const int k = order[head];
        for (int q = 0; q < recs; ++q)
        {
            const float4 rf = wide[f4 * k + 8 * q + 6];
            const float f[4] = {rf.x, rf.y, rf.z, rf.w};
            for (int c = 0; c < 4; ++c)
            {
                int v; memcpy(&v, &f[c], 4);
                if (v >= 0 && v < nbNodes && newIndex[v] < 0) { newIndex[v] = (int)order.size(); order.push_back(v); }
            }
        }
    }
    if ((int)order.size() != nbNodes) return; // not a tree over all nodes: leave as is
    std::vector<float4> out(wide.size());
    for (int n = 0; n < nbNodes; ++n)
    {
        const int src = order[n];
        for (size_t j = 0; j < f4; ++j) out[f4 * n + j] = wide[f4 * src + j];
        for (int q = 0; q < recs; ++q)
        {
            float4& rf = out[f4 * n + 8 * q + 6];
            float* f[4] = {&rf.x, &rf.y, &rf.z, &rf.w};
            for (int c = 0; c < 4; ++c)
            {
                int v; memcpy(&v, f[c], 4);
                if (v >= 0) { v = newIndex[v]; memcpy(f[c], &v, 4); }
            }
        }
    }
    wide.swap(out);
}

// The trees of the order-independent walks (h2d_scene step 2c; also behind the host-only b200_debug_build_walk_trees):
// an unconstrained SAH BVH over PRIMITIVES with tight padded boxes — not over the reference's leaves: level-0 cell keys
// wrap modulo 2^32 (GPUKernel.cpp:938-941), so in large scenes a reference leaf can hold primitives from distant cells and
// span a large part of the scene — followed by the point-query tree of grown cylinder/cone boxes.  primLeaf keeps the
// reference leaf each primitive belongs to, because a hit only counts if that leaf's box passes the reference's slab test.
static void buildWalkTrees(const std::vector<LeafRec>& leaves, const b200_Primitive* prims, int nbPrims, std::vector<float4>& uwide,
                           std::vector<int>& primLeaf, int& nbMain, int& nbExt, int* mainDepth = nullptr)
{
    if (mainDepth) *mainDepth = 0;
    std::vector<float4> ubin, xbin, xwide;
    uwide.clear();
    primLeaf.assign(nbPrims > 0 ? nbPrims : 1, 0);
    nbMain = 0; nbExt = 0;
    if (nbPrims <= 0 || leaves.empty()) return;
    for (size_t l = 0; l < leaves.size(); ++l)
        for (int k = 0; k < leaves[l].count; ++k)
            if (leaves[l].start + k >= 0 && leaves[l].start + k < nbPrims) primLeaf[leaves[l].start + k] = (int)l;
    std::vector<LeafRec> primBoxes(nbPrims), extBoxes;
    for (int i = 0; i < nbPrims; ++i)
    {
        const b200_Primitive& p = prims[i];
        Aabb b;
        const float P0[3] = {p.p0.x, p.p0.y, p.p0.z}, P1[3] = {p.p1.x, p.p1.y, p.p1.z}, P2[3] = {p.p2.x, p.p2.y, p.p2.z};
        const float S[3] = {p.size.x, p.size.y, p.size.z};
        for (int a = 0; a < 3; ++a)
        {
            float lo, hi;
            switch (p.type)
            {
            case B200_PT_TRIANGLE: lo = fminf(fminf(P0[a], P1[a]), P2[a]); hi = fmaxf(fmaxf(P0[a], P1[a]), P2[a]); break;
            case B200_PT_CYLINDER:
            case B200_PT_CONE: lo = fminf(P0[a], P1[a]) - fabsf(S[0]); hi = fmaxf(P0[a], P1[a]) + fabsf(S[0]); break;
            case B200_PT_SPHERE:
            case B200_PT_ENVIRONMENT: lo = P0[a] - fabsf(S[0]); hi = P0[a] + fabsf(S[0]); break;
            default: lo = P0[a] - fabsf(S[a]); hi = P0[a] + fabsf(S[a]); break; // ellipsoid, planes
            }
            // conservative: hit points are computed in float and the cylinder caps accept +-geometryEpsilon
            const float pad = UW_PAD + 2e-5f * fmaxf(fabsf(lo), fabsf(hi));
            b.lo[a] = lo - pad; b.hi[a] = hi + pad;
        }
        primBoxes[i].box = b; primBoxes[i].start = i; primBoxes[i].count = 1;
        // Cylinders and cones also register hits BEHIND the origin: the reference only requires the closest approach of
        // the two lines to lie ahead (t >= 0, GeometryIntersections.cuh:316,381) and takes the entry point t - s
        // whatever its sign, measuring its distance with length() (:690-760).  That needs the origin inside the
        // infinite cylinder, and — because the leaf box must still be ahead (t_max > 0) while the hit point behind the
        // origin lies in it — inside the (convex) leaf box.  Those primitives get a second box, grown over {points of
        // the leaf box within one radius of the axis}; the walks look up the boxes that CONTAIN the ray origin in a
        // separate small tree (a point query) and accept backward hits only from there.
        if ((p.type == B200_PT_CYLINDER || p.type == B200_PT_CONE) && (p.n1.x != 0.f || p.n1.y != 0.f || p.n1.z != 0.f))
        {
            Aabb L = leaves[primLeaf[i]].box;
            L.grow(b); // cone leaves are built from p0 only (GPUKernel.cpp:808-811)
            const double N[3] = {p.n1.x, p.n1.y, p.n1.z};
            const double R = fmax(fabs((double)S[0]), fabs((double)S[1])) * 1.001 + 0.05;
            double u0 = -1e300, u1 = 1e300;
            bool empty = false;
            for (int a = 0; a < 3 && !empty; ++a)
            {
                const double slack = R + 1e-5 * fmax(fabs((double)L.lo[a]), fabs((double)L.hi[a]));
                const double lo = L.lo[a] - slack, hi = L.hi[a] + slack;
                if (fabs(N[a]) < 1e-9) { empty = P0[a] < lo || P0[a] > hi; continue; }
                double ua = (lo - P0[a]) / N[a], ub = (hi - P0[a]) / N[a];
                if (ua > ub) std::swap(ua, ub);
                u0 = fmax(u0, ua); u1 = fmin(u1, ub);
            }
            if (!empty && u0 <= u1 && u0 > -1e299 && u1 < 1e299)
            {
                // a thin diagonal region: cover it with short pieces, each in its own box (an origin near a joint lies
                // in two of them; the walk drops the duplicate)
                int pieces = (int)ceil((u1 - u0) / (4.0 * R));
                pieces = pieces < 1 ? 1 : (pieces > 64 ? 64 : pieces);
                for (int k = 0; k < pieces; ++k)
                {
                    const double ua = u0 + (u1 - u0) * k / pieces, ub = u0 + (u1 - u0) * (k + 1) / pieces;
                    LeafRec x; x.start = i; x.count = 1;
                    for (int a = 0; a < 3; ++a)
                    {
                        const double e0 = P0[a] + ua * N[a], e1 = P0[a] + ub * N[a];
                        const double slack = R + 1e-5 * fmax(fabs(e0), fabs(e1));
                        x.box.lo[a] = (float)(fmin(e0, e1) - slack);
                        x.box.hi[a] = (float)(fmax(e0, e1) + slack);
                    }
                    extBoxes.push_back(x);
                }
            }
        }
    }
    // the two trees are independent: the point-query tree is built on a second thread
    std::vector<int> leafOfNode, leafOfNodeX, primOfNode;
    std::vector<float4> unusedLeafRecs, unusedLeafRecsX;
    int nbUX = 0;
    std::thread extThread([&]() {
        if (extBoxes.empty()) return;
        SahBuilder sx(extBoxes, xbin, leafOfNodeX);
        sx.build(0, (int)extBoxes.size(), 0);
        primOfNode.resize(leafOfNodeX.size());
        for (size_t k = 0; k < leafOfNodeX.size(); ++k) primOfNode[k] = leafOfNodeX[k] < 0 ? -1 : extBoxes[leafOfNodeX[k]].start;
        nbUX = buildWide(xbin, xwide, unusedLeafRecsX, &primOfNode, UW_WIDTH);
    });
    ubin.reserve(4 * (size_t)nbPrims);
    SahBuilder sb(primBoxes, ubin, leafOfNode);
    sb.build(0, nbPrims, 0);
    nbMain = buildWide(ubin, uwide, unusedLeafRecs, &leafOfNode, UW_WIDTH, nullptr, nullptr, mainDepth);
    renumberBreadthFirst(uwide, nbMain);
    extThread.join();
    if (!extBoxes.empty())
    {
        nbExt = nbUX;
        // appended to the first tree: inner refs move by its size, leaf refs get bit 30
        for (size_t k = 0; k < xwide.size() / 8; ++k) // every 128-byte record
        {
            float4& rf = xwide[8 * k + 6];
            float* f[4] = {&rf.x, &rf.y, &rf.z, &rf.w};
            for (int c = 0; c < 4; ++c)
            {
                int v; memcpy(&v, f[c], 4);
                if (v == (int)0x80000000) continue;
                v = v >= 0 ? v + nbMain : ~((~v) | 0x40000000);
                memcpy(f[c], &v, 4);
            }
        }
        uwide.insert(uwide.end(), xwide.begin(), xwide.end());
    }
}

// Scene re-layout.  Input: the flattened AoS arrays of GPUKernel::compactBoxes (GPUKernel.cpp:1085-1281):
// boxes in depth-first order with "slots to skip on a miss" counts.  Output (DESIGN.md "Data layout"):
//   boxes   float4[2n']  (min, w0) (max, w1)    leaf: w0 = first primitive, w1 = count
//                                               inner: w0 = skip, w1 = 0
//   geo     float4[4m]   (p0,size.x) (p1,size.y) (p2,size.z) (n1,0)
//   meta    int[m]       type | fast-transparency | procedural | transparent | materialId
// with every inner box that has exactly one child dropped (its child's slab interval is contained in its
// own, so the pair of tests equals the child's test alone), and skip counts recomputed.
void b200_h2d_scene(b200_int2, const b200_BoundingBox* boxes, int nbBoxes, const b200_Primitive* prims, int nbPrims, const int*, int)
{
    G.uwDirty = true; // the walk trees change: their centre / half-extent form is refreshed before the next frame
    if (!G.initialised) { latch(-4, "b200_h2d_scene", "initialize_scene not called"); return; }
    if (!ensureDevice()) return;
    if (nbBoxes < 0 || nbPrims < 0) { latch(-5, "b200_h2d_scene", "negative count"); return; }
    G.nbBoxesIn = nbBoxes;
    const auto uploadT0 = std::chrono::steady_clock::now();
    const bool timing = getenv("SOLR_B200_TIMING") != nullptr;
    auto lapT = uploadT0;
    auto lap = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(G.stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[solr_b200] h2d_scene %-28s %8.2f ms\n", what, std::chrono::duration<float, std::milli>(now - lapT).count());
        lapT = now;
    };

    std::vector<float4> packed;
    std::vector<LeafRec> leaves;
    const int nOut = relayoutBoxes(boxes, nbBoxes, packed, &G.boxLayoutUsed, &leaves);

    lap("ordered tree (relayout)");
    // 2b. the 4-wide form of the ordered BVH for the per-lane walks
    std::vector<float4> wide, leafRecs;
    std::vector<int> wideKidNodes, leafNodes;
    int wideDepth = 0;
    G.nbWide = (G.boxLayoutUsed == 2) ? buildWide(packed, wide, leafRecs, nullptr, 4, &wideKidNodes, &leafNodes, &wideDepth) : 0;
    if (3 * wideDepth + 3 > WIDE_STACK)
    {
        // closestHitWide / shadowWalkWide keep up to three entries per level in a fixed stack (trace.cuh WIDE_STACK): a tree this deep
        // (SAH splits of a badly skewed scene) is walked as the stackless list instead
        G.nbWide = 0; wide.clear(); leafRecs.clear(); wideKidNodes.clear(); leafNodes.clear();
    }
    if (wide.size() > G.capWide) { freeDev(G.dWide); G.capWide = wide.size() + 1024; CK(cudaMalloc(&G.dWide, G.capWide * sizeof(float4))); }
    if (leafRecs.size() > G.capLeafRecs) { freeDev(G.dLeafRecs); G.capLeafRecs = leafRecs.size() + 1024; CK(cudaMalloc(&G.dLeafRecs, G.capLeafRecs * sizeof(float4))); }
    G.nWideF4 = wide.size(); G.nLeafRecsF4 = leafRecs.size();
    if (!wide.empty()) CK(cudaMemcpyAsync(G.dWide, wide.data(), wide.size() * sizeof(float4), cudaMemcpyHostToDevice, G.stream));
    if (!leafRecs.empty()) CK(cudaMemcpyAsync(G.dLeafRecs, leafRecs.data(), leafRecs.size() * sizeof(float4), cudaMemcpyHostToDevice, G.stream));
    lap("4-wide ordered tree + upload");
    // 2c. the trees of the order-independent walks
    std::vector<float4> uwide;
    std::vector<int> primLeaf(nbPrims > 0 ? nbPrims : 1, 0);
    G.nbUWide = 0; G.nbUX = 0;
    const bool wantTrees = G.boxLayoutUsed == 2 && G.nbWide > 0 && g_useUnordered && nbPrims > 0;
    const bool gpuTrees = wantTrees && g_gpuTrees && !UW_GROUP;
    if (gpuTrees)
    {
        for (size_t l = 0; l < leaves.size(); ++l)
            for (int k = 0; k < leaves[l].count; ++k)
                if (leaves[l].start + k >= 0 && leaves[l].start + k < nbPrims) primLeaf[leaves[l].start + k] = (int)l;
    }
    else if (wantTrees) buildWalkTrees(leaves, prims, nbPrims, uwide, primLeaf, G.nbUWide, G.nbUX, &G.walkTreeLevels);
    lap("walk trees on the host");
    if ((size_t)nbPrims > G.capPrimLeaf) { freeDev(G.dPrimLeaf); G.capPrimLeaf = (size_t)nbPrims + 1024; CK(cudaMalloc(&G.dPrimLeaf, G.capPrimLeaf * sizeof(int))); }
    if (nbPrims > 0) CK(cudaMemcpyAsync(G.dPrimLeaf, primLeaf.data(), (size_t)nbPrims * sizeof(int), cudaMemcpyHostToDevice, G.stream));
    if (uwide.size() > G.capUWide) { freeDev(G.dUWide); G.capUWide = uwide.size() + 1024; CK(cudaMalloc(&G.dUWide, G.capUWide * sizeof(float4))); }
    if (!uwide.empty()) CK(cudaMemcpyAsync(G.dUWide, uwide.data(), uwide.size() * sizeof(float4), cudaMemcpyHostToDevice, G.stream));
    if (gpuTrees)
    {
        // the primitives and the reference leaves' boxes go up first; the trees are built from them on the device
        if ((size_t)nbPrims > G.capPrims)
        {
            freeDev(G.dGeo); freeDev(G.dMeta); freeDev(G.dPrims);
            G.capPrims = (size_t)nbPrims + 1024;
            CK(cudaMalloc(&G.dGeo, 4 * G.capPrims * sizeof(float4)));
            CK(cudaMalloc(&G.dMeta, G.capPrims * sizeof(int)));
            CK(cudaMalloc(&G.dPrims, G.capPrims * sizeof(b200_Primitive)));
        }
        CK(cudaMemcpyAsync(G.dPrims, prims, (size_t)nbPrims * sizeof(b200_Primitive), cudaMemcpyHostToDevice, G.stream));
        std::vector<float4> leafBoxes(2 * leaves.size());
        for (size_t l = 0; l < leaves.size(); ++l)
        {
            leafBoxes[2 * l] = make_float4(leaves[l].box.lo[0], leaves[l].box.lo[1], leaves[l].box.lo[2], 0.f);
            leafBoxes[2 * l + 1] = make_float4(leaves[l].box.hi[0], leaves[l].box.hi[1], leaves[l].box.hi[2], 0.f);
        }
        if (leafBoxes.size() > G.capLeafBoxes) { freeDev(G.dLeafBoxes); G.capLeafBoxes = leafBoxes.size() + 1024; CK(cudaMalloc(&G.dLeafBoxes, G.capLeafBoxes * sizeof(float4))); }
        CK(cudaMemcpyAsync(G.dLeafBoxes, leafBoxes.data(), leafBoxes.size() * sizeof(float4), cudaMemcpyHostToDevice, G.stream));
        const int nbExtBoxes = g_useBackward ? treebuild::extCount(G.dPrims, nbPrims, G.dPrimLeaf, G.dLeafBoxes, G.stream) : 0;
        int rc = nbExtBoxes < 0 ? nbExtBoxes : 0;
        if (rc == 0)
        {
            const size_t wantF4 = 8 * ((size_t)nbPrims + (size_t)nbExtBoxes) + 8;
            if (wantF4 > G.capUWide) { CK(cudaStreamSynchronize(G.stream)); freeDev(G.dUWide); G.capUWide = wantF4 + 1024; CK(cudaMalloc(&G.dUWide, G.capUWide * sizeof(float4))); }
            rc = treebuild::buildWalkTreesGpu(G.dPrims, nbPrims, G.dPrimLeaf, G.dLeafBoxes, nbExtBoxes, G.dUWide, G.nbUWide, G.nbUX, G.stream, &G.walkTreeLevels);
        }
        cudaError_t e = cudaGetLastError();
        if (rc != 0 || e != cudaSuccess)
        {
            latch(rc != 0 ? rc : (int)e, "b200_h2d_scene", "building the walk trees on the GPU failed");
            G.nbUWide = 0; G.nbUX = 0;
        }
    }
    lap("walk trees on the GPU");
    // child-major copy for the group walk: child c of node n at float4 (n * width + c) * 2 — (lo.xyz, hi.x) (hi.yz, ref, -)
    std::vector<float4> ugroup;
#if UW_GROUP
    {
        const int recs = UW_WIDTH / 4;
        const size_t nbNodes = uwide.size() / (8 * (size_t)recs);
        ugroup.resize(nbNodes * UW_WIDTH * 2 + 2);
        for (size_t n = 0; n < nbNodes; ++n)
            for (int c = 0; c < UW_WIDTH; ++c)
            {
                const float4* rec = &uwide[8 * (n * recs + c / 4)];
                const int k = c % 4;
                auto comp = [k](const float4& v) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; };
                ugroup[(n * UW_WIDTH + c) * 2] = make_float4(comp(rec[0]), comp(rec[1]), comp(rec[2]), comp(rec[3]));
                ugroup[(n * UW_WIDTH + c) * 2 + 1] = make_float4(comp(rec[4]), comp(rec[5]), comp(rec[6]), 0.f);
            }
        if (ugroup.size() > G.capUGroup) { freeDev(G.dUGroup); G.capUGroup = ugroup.size() + 1024; CK(cudaMalloc(&G.dUGroup, G.capUGroup * sizeof(float4))); }
        if (nbNodes > 0) CK(cudaMemcpyAsync(G.dUGroup, ugroup.data(), ugroup.size() * sizeof(float4), cudaMemcpyHostToDevice, G.stream));
    }
#endif
    CK(cudaStreamSynchronize(G.stream));

    lap("uploads of the trees");
    G.hPrims.assign(prims, prims + nbPrims);

    // 4. device buffers (grow-only) and upload
    if ((size_t)nOut > G.capBoxes || (size_t)nbBoxes > G.capBoxes)
    {
        freeDev(G.dBoxes); freeDev(G.dRawBoxes);
        G.capBoxes = (size_t)(nbBoxes > nOut ? nbBoxes : nOut) + 1024;
        CK(cudaMalloc(&G.dBoxes, 2 * G.capBoxes * sizeof(float4)));
        CK(cudaMalloc(&G.dRawBoxes, G.capBoxes * sizeof(b200_BoundingBox)));
    }
    if ((size_t)nbPrims > G.capPrims)
    {
        freeDev(G.dGeo); freeDev(G.dMeta); freeDev(G.dPrims);
        G.capPrims = (size_t)nbPrims + 1024;
        CK(cudaMalloc(&G.dGeo, 4 * G.capPrims * sizeof(float4)));
        CK(cudaMalloc(&G.dMeta, G.capPrims * sizeof(int)));
        CK(cudaMalloc(&G.dPrims, G.capPrims * sizeof(b200_Primitive)));
    }
    if (nOut) CK(cudaMemcpyAsync(G.dBoxes, packed.data(), packed.size() * sizeof(float4), cudaMemcpyHostToDevice, G.stream));
    if (nbBoxes) CK(cudaMemcpyAsync(G.dRawBoxes, boxes, (size_t)nbBoxes * sizeof(b200_BoundingBox), cudaMemcpyHostToDevice, G.stream));
    if (nbPrims)
    {
        CK(cudaMemcpyAsync(G.dPrims, prims, (size_t)nbPrims * sizeof(b200_Primitive), cudaMemcpyHostToDevice, G.stream));
        // 3. hot geometry, and the unit walk's records — geometry + (packed word, patched in by uploadMeta) + the reference leaf's box
        //    and number + the original id, so that a leaf visit is one dependent access (trace.cuh) — derived on the device from what
        //    is there already (the kernel the animation step uses: animate.cuh)
        if (G.nbUWide > 0)
        {
            if (!gpuTrees)
            {
                std::vector<float4> leafBoxes(2 * leaves.size());
                for (size_t l = 0; l < leaves.size(); ++l)
                {
                    leafBoxes[2 * l] = make_float4(leaves[l].box.lo[0], leaves[l].box.lo[1], leaves[l].box.lo[2], 0.f);
                    leafBoxes[2 * l + 1] = make_float4(leaves[l].box.hi[0], leaves[l].box.hi[1], leaves[l].box.hi[2], 0.f);
                }
                if (leafBoxes.size() > G.capLeafBoxes) { freeDev(G.dLeafBoxes); G.capLeafBoxes = leafBoxes.size() + 1024; CK(cudaMalloc(&G.dLeafBoxes, G.capLeafBoxes * sizeof(float4))); }
                CK(cudaMemcpyAsync(G.dLeafBoxes, leafBoxes.data(), leafBoxes.size() * sizeof(float4), cudaMemcpyHostToDevice, G.stream));
                CK(cudaStreamSynchronize(G.stream)); // staging vector dies here
            }
            const size_t recsF4 = (size_t)PRIM_REC_F4 * nbPrims + 2; // + 32 bytes: a 256-bit load never straddles the end
            if (recsF4 > G.capPrimRecs) { freeDev(G.dPrimRecs); G.capPrimRecs = recsF4 + 1024; CK(cudaMalloc(&G.dPrimRecs, G.capPrimRecs * sizeof(float4))); }
        }
        animate::k_an_records<<<(nbPrims + 255) / 256, 256, 0, G.stream>>>(G.dPrims, nbPrims, G.dPrimLeaf, G.dLeafBoxes, G.dGeo, G.nbUWide > 0 ? G.dPrimRecs : nullptr);
    }
    CK(cudaStreamSynchronize(G.stream)); // staging vectors die at scope exit
    G.nbBoxes = nOut; G.nbPrims = nbPrims;
    lap("boxes, primitives up; geometry + records");
    uploadMeta();
    lap("packed words");
    // what the device-side animation step needs besides the arrays themselves (animate.cuh)
    G.animatable = false;
    if (G.boxLayoutUsed == 2 && G.nbWide > 0 && leafNodes.size() == leaves.size() && !leaves.empty())
    {
        std::vector<int> leafRaw;
        leafRaw.reserve(leaves.size());
        int maxLevel = 0;
        for (int i = 0; i < nbBoxes; ++i)
        {
            if (boxes[i].nbPrimitives > 0) leafRaw.push_back(i);
            else if (boxes[i].startIndex > maxLevel) maxLevel = boxes[i].startIndex;
        }
        const size_t nbPacked = packed.size() / 2;
        std::vector<int> parent(nbPacked, -1);
        {
            // depth-first list with subtree sizes: a node's children are the next node and the one after the first child's subtree
            std::vector<std::pair<int, int>> stack; // (node, end)
            for (size_t i = 0; i < nbPacked; ++i)
            {
                while (!stack.empty() && stack.back().second <= (int)i) stack.pop_back();
                parent[i] = stack.empty() ? -1 : stack.back().first;
                int w0, w1;
                memcpy(&w0, &packed[2 * i].w, 4); memcpy(&w1, &packed[2 * i + 1].w, 4);
                if (w1 <= 0) stack.push_back({(int)i, (int)i + (w0 > 1 ? w0 : 1)});
            }
        }
        bool ok = leafRaw.size() == leaves.size() && wideKidNodes.size() == 4 * (size_t)G.nbWide;
        for (size_t k = 0; ok && k < leaves.size(); ++k) ok = boxes[leafRaw[k]].startIndex == leaves[k].start && boxes[leafRaw[k]].nbPrimitives == leaves[k].count;
        if (ok)
        {
            if (leaves.size() > G.capLeafMaps)
            {
                freeDev(G.dLeafRaw); freeDev(G.dLeafNode);
                G.capLeafMaps = leaves.size() + 1024;
                CK(cudaMalloc(&G.dLeafRaw, G.capLeafMaps * sizeof(int))); CK(cudaMalloc(&G.dLeafNode, G.capLeafMaps * sizeof(int)));
            }
            if (nbPacked > G.capPackedParent)
            {
                freeDev(G.dPackedParent); freeDev(G.dFitFlags);
                G.capPackedParent = nbPacked + 1024;
                CK(cudaMalloc(&G.dPackedParent, G.capPackedParent * sizeof(int))); CK(cudaMalloc(&G.dFitFlags, G.capPackedParent * sizeof(int)));
            }
            if ((size_t)G.nbWide > G.capWideKid) { freeDev(G.dWideKid); G.capWideKid = (size_t)G.nbWide + 1024; CK(cudaMalloc(&G.dWideKid, G.capWideKid * sizeof(int4))); }
            CK(cudaMemcpyAsync(G.dLeafRaw, leafRaw.data(), leafRaw.size() * sizeof(int), cudaMemcpyHostToDevice, G.stream));
            CK(cudaMemcpyAsync(G.dLeafNode, leafNodes.data(), leafNodes.size() * sizeof(int), cudaMemcpyHostToDevice, G.stream));
            CK(cudaMemcpyAsync(G.dPackedParent, parent.data(), nbPacked * sizeof(int), cudaMemcpyHostToDevice, G.stream));
            CK(cudaMemcpyAsync(G.dWideKid, wideKidNodes.data(), wideKidNodes.size() * sizeof(int), cudaMemcpyHostToDevice, G.stream));
            CK(cudaStreamSynchronize(G.stream));
            G.nbLeaves = (int)leaves.size(); G.nbPacked = (int)nbPacked; G.maxBoxLevel = maxLevel;
            G.nbLightPrims = (nbBoxes > 0 && boxes[0].nbPrimitives > 0 && boxes[0].startIndex == 0) ? boxes[0].nbPrimitives : 0;
            G.animatable = G.err == 0;
        }
    }
    lap("animation maps");
    G.treesOnGpu = gpuTrees ? 1 : 0;
    G.uploadMs = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - uploadT0).count();
}

// ----------------------------------------------------------------------------------------------------
// Scene replication for the multi-GPU frame split: one process uploads the scene (b200_h2d_scene: host tree builds and all), the
// others adopt its layout and receive the device arrays from it over NVLink (sol-r_b200/partition.py broadcast_scene: one NCCL
// broadcast per array) instead of each repeating the host work and its own PCIe copies — the reference's dormant multi-GPU path
// uploads every array to every device from the host (CudaRayTracer.cu:1540-1613 inside the per-device loop).
// Materials, lights, textures and randoms are small and go up per process as before.
// ----------------------------------------------------------------------------------------------------
#define SCENE_LAYOUT_ENTRIES 16
static size_t sceneArrayList(void** ptrs, long long* bytes, int cap)
{
    const size_t an = G.animatable ? 1 : 0; // the maps of the device-side animation travel with an animatable scene
    void* p[14] = {G.dBoxes, G.dRawBoxes, G.dGeo, G.dMeta, G.dPrims, G.dWide, G.dLeafRecs, G.dPrimLeaf, G.dPrimRecs, G.dUWide,
                   G.dLeafRaw, G.dLeafNode, G.dPackedParent, G.dWideKid};
    const long long b[14] = {
        (long long)(2 * (size_t)G.nbBoxes * sizeof(float4)), (long long)((size_t)G.nbBoxesIn * sizeof(b200_BoundingBox)),
        (long long)(4 * (size_t)G.nbPrims * sizeof(float4)), (long long)((size_t)G.nbPrims * sizeof(int)),
        (long long)((size_t)G.nbPrims * sizeof(b200_Primitive)), (long long)(G.nWideF4 * sizeof(float4)), (long long)(G.nLeafRecsF4 * sizeof(float4)),
        (long long)((size_t)G.nbPrims * sizeof(int)), (long long)(G.nbUWide > 0 ? ((size_t)PRIM_REC_F4 * G.nbPrims + 2) * sizeof(float4) : 0),
        (long long)(8 * ((size_t)G.nbUWide + (size_t)G.nbUX) * sizeof(float4)),
        (long long)(an * (size_t)G.nbLeaves * sizeof(int)), (long long)(an * (size_t)G.nbLeaves * sizeof(int)),
        (long long)(an * (size_t)G.nbPacked * sizeof(int)), (long long)(an * (size_t)G.nbWide * sizeof(int4))};
    int n = 0;
    for (int k = 0; k < 14 && n < cap; ++k, ++n) { if (ptrs) ptrs[n] = b[k] ? p[k] : nullptr; if (bytes) bytes[n] = b[k]; }
    return (size_t)n;
}

int b200_scene_layout(long long* layout, int capacity)
{
    if (!layout || capacity < SCENE_LAYOUT_ENTRIES) return -4;
    const long long v[SCENE_LAYOUT_ENTRIES] = {G.nbBoxesIn, G.nbBoxes, G.nbPrims, G.boxLayoutUsed, G.nbWide, (long long)G.nWideF4, (long long)G.nLeafRecsF4,
                                               G.nbUWide, G.nbUX, G.treesOnGpu, G.animatable ? 1 : 0, G.nbLeaves, G.nbPacked, G.maxBoxLevel, G.nbLightPrims, G.walkTreeLevels};
    for (int k = 0; k < SCENE_LAYOUT_ENTRIES; ++k) layout[k] = v[k];
    return SCENE_LAYOUT_ENTRIES;
}

int b200_scene_adopt_layout(const long long* layout, int n)
{
    if (!G.initialised) { latch(-4, "b200_scene_adopt_layout", "initialize_scene not called"); return -4; }
    if (!layout || n < SCENE_LAYOUT_ENTRIES) return -4;
    if (!ensureDevice()) return -1;
    CK(cudaStreamSynchronize(G.stream));
    const int nbBoxesIn = (int)layout[0], nOut = (int)layout[1], nbPrims = (int)layout[2];
    if (nbBoxesIn < 0 || nOut < 0 || nbPrims < 0) { latch(-5, "b200_scene_adopt_layout", "negative count"); return -5; }
    G.nbBoxesIn = nbBoxesIn; G.nbBoxes = nOut; G.nbPrims = nbPrims; G.boxLayoutUsed = (int)layout[3];
    G.nbWide = (int)layout[4]; G.nWideF4 = (size_t)layout[5]; G.nLeafRecsF4 = (size_t)layout[6];
    G.nbUWide = (int)layout[7]; G.nbUX = (int)layout[8]; G.treesOnGpu = (int)layout[9];
    if ((size_t)nOut > G.capBoxes || (size_t)nbBoxesIn > G.capBoxes)
    {
        freeDev(G.dBoxes); freeDev(G.dRawBoxes);
        G.capBoxes = (size_t)(nbBoxesIn > nOut ? nbBoxesIn : nOut) + 1024;
        CK(cudaMalloc(&G.dBoxes, 2 * G.capBoxes * sizeof(float4)));
        CK(cudaMalloc(&G.dRawBoxes, G.capBoxes * sizeof(b200_BoundingBox)));
    }
    if ((size_t)nbPrims > G.capPrims)
    {
        freeDev(G.dGeo); freeDev(G.dMeta); freeDev(G.dPrims);
        G.capPrims = (size_t)nbPrims + 1024;
        CK(cudaMalloc(&G.dGeo, 4 * G.capPrims * sizeof(float4)));
        CK(cudaMalloc(&G.dMeta, G.capPrims * sizeof(int)));
        CK(cudaMalloc(&G.dPrims, G.capPrims * sizeof(b200_Primitive)));
    }
    if (G.nWideF4 > G.capWide) { freeDev(G.dWide); G.capWide = G.nWideF4 + 1024; CK(cudaMalloc(&G.dWide, G.capWide * sizeof(float4))); }
    if (G.nLeafRecsF4 > G.capLeafRecs) { freeDev(G.dLeafRecs); G.capLeafRecs = G.nLeafRecsF4 + 1024; CK(cudaMalloc(&G.dLeafRecs, G.capLeafRecs * sizeof(float4))); }
    if ((size_t)nbPrims > G.capPrimLeaf) { freeDev(G.dPrimLeaf); G.capPrimLeaf = (size_t)nbPrims + 1024; CK(cudaMalloc(&G.dPrimLeaf, G.capPrimLeaf * sizeof(int))); }
    const size_t recsF4 = G.nbUWide > 0 ? (size_t)PRIM_REC_F4 * nbPrims + 2 : 0;
    if (recsF4 > G.capPrimRecs) { freeDev(G.dPrimRecs); G.capPrimRecs = recsF4 + 1024; CK(cudaMalloc(&G.dPrimRecs, G.capPrimRecs * sizeof(float4))); }
    const size_t uwF4 = 8 * ((size_t)G.nbUWide + (size_t)G.nbUX);
    if (uwF4 > G.capUWide) { freeDev(G.dUWide); G.capUWide = uwF4 + 1024; CK(cudaMalloc(&G.dUWide, G.capUWide * sizeof(float4))); }
    G.animatable = layout[10] != 0; G.nbLeaves = (int)layout[11]; G.nbPacked = (int)layout[12]; G.maxBoxLevel = (int)layout[13]; G.nbLightPrims = (int)layout[14]; G.walkTreeLevels = (int)layout[15];
    if (G.animatable)
    {
        if ((size_t)G.nbLeaves > G.capLeafMaps)
        {
            freeDev(G.dLeafRaw); freeDev(G.dLeafNode);
            G.capLeafMaps = (size_t)G.nbLeaves + 1024;
            CK(cudaMalloc(&G.dLeafRaw, G.capLeafMaps * sizeof(int))); CK(cudaMalloc(&G.dLeafNode, G.capLeafMaps * sizeof(int)));
        }
        if ((size_t)G.nbPacked > G.capPackedParent)
        {
            freeDev(G.dPackedParent); freeDev(G.dFitFlags);
            G.capPackedParent = (size_t)G.nbPacked + 1024;
            CK(cudaMalloc(&G.dPackedParent, G.capPackedParent * sizeof(int))); CK(cudaMalloc(&G.dFitFlags, G.capPackedParent * sizeof(int)));
        }
        if ((size_t)G.nbWide > G.capWideKid) { freeDev(G.dWideKid); G.capWideKid = (size_t)G.nbWide + 1024; CK(cudaMalloc(&G.dWideKid, G.capWideKid * sizeof(int4))); }
    }
    return G.err;
}

int b200_scene_device_arrays(void** ptrs, long long* bytes, int capacity)
{
    if (!ensureDevice()) return -1;
    CK(cudaStreamSynchronize(G.stream)); // the arrays are about to be read or written by somebody else's stream
    return (int)sceneArrayList(ptrs, bytes, capacity);
}

int b200_scene_adopt_finish(void)
{
    G.uwDirty = true; // the walk trees change: their centre / half-extent form is refreshed before the next frame
    if (!ensureDevice()) return -1;
    // the host copy of the primitives (the packed words are re-derived from it whenever the materials change)
    G.hPrims.resize((size_t)G.nbPrims);
    if (G.nbPrims > 0) CK(cudaMemcpy(G.hPrims.data(), G.dPrims, (size_t)G.nbPrims * sizeof(b200_Primitive), cudaMemcpyDeviceToHost));
    if (!G.hMats.empty()) uploadMeta();
    else
    {
        // no materials here yet: the packed words that arrived are the root's; whether every shadow caster is opaque follows from them
        // only together with the materials, so stay on the safe side until b200_h2d_materials runs
        G.opaqueShadows = 0;
    }
    return G.err;
}

// ----------------------------------------------------------------------------------------------------
// The animation step on the device-resident scene (animate.cuh has the what and why).
// ----------------------------------------------------------------------------------------------------
static int animateScene(const animate::Move& move)
{
    G.uwDirty = true; // the walk trees change: their centre / half-extent form is refreshed before the next frame
    if (!G.initialised || !ensureDevice()) return -1;
    if (!G.animatable || G.nbPrims <= 0) { latch(-13, "device-side animation", "needs a scene uploaded with the ordered-tree layout (b200_h2d_scene, options 1 and 4 at their defaults)"); return -13; }
    const auto t0 = std::chrono::steady_clock::now();
    const int T = 256;
    const int nbMovable = G.nbPrims - (move.mode == animate::MOVE_SCALE ? 0 : G.nbLightPrims);
    const int firstMovable = move.mode == animate::MOVE_SCALE ? 0 : G.nbLightPrims;
    if (nbMovable > 0) animate::k_an_move<<<(nbMovable + T - 1) / T, T, 0, G.stream>>>(G.dPrims, firstMovable, G.nbPrims, move);
    if (move.mode == animate::MOVE_SCALE && G.nbLights > 0) animate::k_an_scale_lights<<<(G.nbLights + T - 1) / T, T, 0, G.stream>>>(G.dLights, G.nbLights, move.s);
    if (move.mode != animate::MOVE_SCALE && G.nbBoxesIn > 1)
    {
        // the reference's scalePrimitives leaves the boxes alone (GPUKernel.cpp:1574-1600); rotate and translate re-fit them
        const int B = (G.nbBoxesIn - 1 + T - 1) / T;
        animate::k_an_leaf_boxes<<<B, T, 0, G.stream>>>(G.dRawBoxes, G.nbBoxesIn, G.dPrims, G.nbPrims);
        for (int level = 1; level <= G.maxBoxLevel; ++level)
            animate::k_an_inner_boxes<<<B, T, 0, G.stream>>>(G.dRawBoxes, G.nbBoxesIn, level, G.viewDistance);
    }
    // derived: ordered tree, its wide form and leaf records, geometry and primitive records, walk trees
    if ((size_t)2 * G.nbLeaves > G.capLeafBoxes) { CK(cudaStreamSynchronize(G.stream)); freeDev(G.dLeafBoxes); G.capLeafBoxes = (size_t)2 * G.nbLeaves + 1024; CK(cudaMalloc(&G.dLeafBoxes, G.capLeafBoxes * sizeof(float4))); }
    const int BL = (G.nbLeaves + T - 1) / T;
    animate::k_an_ordered_leaves<<<BL, T, 0, G.stream>>>(G.dRawBoxes, G.dLeafRaw, G.dLeafNode, G.nbLeaves, G.dBoxes, G.dLeafBoxes);
    CK(cudaMemsetAsync(G.dFitFlags, 0, (size_t)G.nbPacked * sizeof(int), G.stream));
    animate::k_an_ordered_fit<<<BL, T, 0, G.stream>>>(G.dBoxes, G.dPackedParent, G.dLeafNode, G.nbLeaves, G.dFitFlags);
    animate::k_an_ordered_wide<<<(G.nbWide + T - 1) / T, T, 0, G.stream>>>(G.dBoxes, G.dWideKid, G.nbWide, G.dWide);
    animate::k_an_leaf_recs<<<BL, T, 0, G.stream>>>(G.dBoxes, G.dLeafNode, G.nbLeaves, G.dLeafRecs);
    animate::k_an_records<<<(G.nbPrims + T - 1) / T, T, 0, G.stream>>>(G.dPrims, G.nbPrims, G.dPrimLeaf, G.dLeafBoxes, G.dGeo, G.nbUWide > 0 ? G.dPrimRecs : nullptr);
    int rc = 0;
    if (G.nbUWide > 0)
    {
        const int nbExtBoxes = g_useBackward ? treebuild::extCount(G.dPrims, G.nbPrims, G.dPrimLeaf, G.dLeafBoxes, G.stream) : 0;
        rc = nbExtBoxes < 0 ? nbExtBoxes : 0;
        if (rc == 0)
        {
            const bool refit = g_animateRefit && G.walkTreeLevels > 0;
            const size_t wantF4 = 8 * ((refit ? (size_t)G.nbUWide : (size_t)G.nbPrims) + (size_t)nbExtBoxes) + 8;
            if (wantF4 > G.capUWide)
            {
                // more room behind the main tree, which moves along when it is kept
                float4* grown = nullptr;
                CK(cudaMalloc(&grown, (wantF4 + 1024) * sizeof(float4)));
                if (refit && grown && G.dUWide) CK(cudaMemcpyAsync(grown, G.dUWide, 8 * (size_t)G.nbUWide * sizeof(float4), cudaMemcpyDeviceToDevice, G.stream));
                CK(cudaStreamSynchronize(G.stream));
                freeDev(G.dUWide);
                G.dUWide = grown; G.capUWide = wantF4 + 1024;
            }
            // The main tree keeps its shape — a rigid move leaves a good clustering good, and the host's SAH tree is the better one —
            // and is re-fitted in place; the point-query tree's boxes change their number with the orientation and are rebuilt.
            if (refit) rc = treebuild::refitMain(G.dUWide, G.nbUWide, G.walkTreeLevels, G.dPrims, G.stream);
            if (rc == 0)
                rc = treebuild::buildWalkTreesGpu(G.dPrims, G.nbPrims, G.dPrimLeaf, G.dLeafBoxes, nbExtBoxes, G.dUWide, G.nbUWide, G.nbUX, G.stream,
                                                  refit ? nullptr : &G.walkTreeLevels, refit);
            if (!refit) G.treesOnGpu = 1;
        }
    }
    cudaError_t e = cudaGetLastError();
    if (rc != 0 || e != cudaSuccess) { latch(rc != 0 ? rc : (int)e, "device-side animation", "a kernel or the tree build failed"); G.animatable = false; return G.err; }
    // the host copy of the primitives backs the packed material words only (type and material id: untouched by a move)
    CK(cudaStreamSynchronize(G.stream));
    G.animateMs = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return G.err;
}

int b200_rotate_primitives(b200_float3 center, b200_float3 angles)
{
    animate::Move m;
    memset(&m, 0, sizeof(m));
    m.mode = animate::MOVE_ROTATE;
    m.center = make_float3(center.x, center.y, center.z);
    // GPUKernel.cpp:1383-1391: the cosines and sines are taken in double and kept as floats
    m.cosA = make_float3((float)cos((double)angles.x), (float)cos((double)angles.y), (float)cos((double)angles.z));
    m.sinA = make_float3((float)sin((double)angles.x), (float)sin((double)angles.y), (float)sin((double)angles.z));
    return animateScene(m);
}
int b200_translate_primitives(b200_float3 t)
{
    animate::Move m;
    memset(&m, 0, sizeof(m));
    m.mode = animate::MOVE_TRANSLATE;
    m.t = make_float3(t.x, t.y, t.z);
    return animateScene(m);
}
int b200_scale_primitives(float scale)
{
    animate::Move m;
    memset(&m, 0, sizeof(m));
    m.mode = animate::MOVE_SCALE;
    m.s = scale;
    return animateScene(m);
}
int b200_d2h_scene(b200_BoundingBox* boxes, b200_Primitive* prims)
{
    if (!G.initialised || !ensureDevice()) return -1;
    if (boxes && G.nbBoxesIn > 0) CK(cudaMemcpyAsync(boxes, G.dRawBoxes, (size_t)G.nbBoxesIn * sizeof(b200_BoundingBox), cudaMemcpyDeviceToHost, G.stream));
    if (prims && G.nbPrims > 0) CK(cudaMemcpyAsync(prims, G.dPrims, (size_t)G.nbPrims * sizeof(b200_Primitive), cudaMemcpyDeviceToHost, G.stream));
    CK(cudaStreamSynchronize(G.stream));
    return G.err;
}
float b200_last_animation_ms(void) { return G.animateMs; }

void b200_h2d_materials(b200_int2, const b200_Material* materials, int n)
{
    if (!G.initialised) { latch(-4, "b200_h2d_materials", "initialize_scene not called"); return; }
    if (!ensureDevice() || n <= 0) return;
    if ((size_t)n > G.capMats)
    {
        // always NB_MAX_MATERIALS slots, zero-filled, like the reference's device array (CudaRayTracer.cu:1459-1462):
        // the box-debug view indexes it with startIndex % NB_MAX_MATERIALS
        freeDev(G.dMats);
        G.capMats = (size_t)(n > B200_NB_MAX_MATERIALS ? n : B200_NB_MAX_MATERIALS) + 1;
        CK(cudaMalloc(&G.dMats, G.capMats * sizeof(b200_Material)));
        CK(cudaMemsetAsync(G.dMats, 0, G.capMats * sizeof(b200_Material), G.stream));
    }
    CK(cudaMemcpyAsync(G.dMats, materials, (size_t)n * sizeof(b200_Material), cudaMemcpyHostToDevice, G.stream));
    CK(cudaStreamSynchronize(G.stream));
    G.hMats.assign(materials, materials + n);
    G.nbMats = n;
    uploadMeta(); // the packed word caches three material bits
}

void b200_h2d_randoms(b200_int2, const float* randoms)
{
    if (!G.dRandoms) { latch(-4, "b200_h2d_randoms", "reshape_scene not called"); return; }
    if (!ensureDevice()) return;
    CK(cudaMemcpyAsync(G.dRandoms, randoms, (size_t)G.maxW * G.maxH * sizeof(float), cudaMemcpyHostToDevice, G.stream));
    CK(cudaStreamSynchronize(G.stream));
}

void b200_h2d_textures(b200_int2, int nbTextures, const b200_TextureInfo* infos)
{
    if (!G.initialised) { latch(-4, "b200_h2d_textures", "initialize_scene not called"); return; }
    if (!ensureDevice()) return;
    size_t total = 0;
    for (int i = 0; i < nbTextures; ++i)
        if (infos[i].buffer)
        {
            const size_t end = (size_t)infos[i].offset + (size_t)infos[i].size.x * infos[i].size.y * infos[i].size.z;
            if (end > total) total = end;
        }
    freeDev(G.dTex);
    G.texBytes = total;
    if (!total) return;
    CK(cudaMalloc(&G.dTex, total + 16));
    CK(cudaMemsetAsync(G.dTex, 0, total + 16, G.stream));
    for (int i = 0; i < nbTextures; ++i)
        if (infos[i].buffer)
            CK(cudaMemcpyAsync(G.dTex + infos[i].offset, infos[i].buffer, (size_t)infos[i].size.x * infos[i].size.y * infos[i].size.z,
                               cudaMemcpyHostToDevice, G.stream));
    CK(cudaStreamSynchronize(G.stream));
}

void b200_h2d_lightInformation(b200_int2, const b200_LightInformation* li, int n)
{
    if (!G.initialised) { latch(-4, "b200_h2d_lightInformation", "initialize_scene not called"); return; }
    if (!ensureDevice()) return;
    if (n > B200_NB_MAX_LIGHTINFORMATIONS) n = B200_NB_MAX_LIGHTINFORMATIONS;
    if (n > 0)
    {
        CK(cudaMemcpyAsync(G.dLights, li, (size_t)n * sizeof(b200_LightInformation), cudaMemcpyHostToDevice, G.stream));
        CK(cudaStreamSynchronize(G.stream));
    }
    G.nbLights = n;
}

// every pixel of this GPU's tiles, device buffers -> the host buffers named by b200_stream_target (k_stream_own_tiles)
static void streamOwnTiles(int W, int H, bool bgr)
{
    const int tilesX = (W + TILE_W - 1) / TILE_W, tilesY = (H + TILE_H - 1) / TILE_H;
    const int nbLocal = (tilesX * tilesY - G.rank + G.world - 1) / G.world;
    if (nbLocal <= 0 || (!G.mirrorDevBitmap && !G.mirrorDevIds)) return;
    int grid = (nbLocal + 7) / 8;
    if (grid > G.numSMs * 8) grid = G.numSMs * 8;
    k_stream_own_tiles<<<grid, 256, 0, G.stream>>>(G.dIds, G.dBitmap, G.validIds ? G.mirrorDevIds : nullptr, G.validBitmap ? G.mirrorDevBitmap : nullptr,
                                                   W, H, tilesX, nbLocal, G.rank, G.world, bgr ? 1 : 0);
    G.launches++;
}

void b200_render(b200_int2, b200_int4, b200_SceneInfo si, b200_int4 objects, b200_PostProcessingInfo pp, b200_float3 origin,
                 b200_float3 direction, b200_float4 angles)
{
    if (!G.initialised || !G.dPost) { latch(-4, "b200_render", "initialize_scene/reshape_scene not called"); return; }
    G.viewDistance = si.viewDistance;
    if (!ensureDevice()) return;
    if (si.size.x <= 0 || si.size.y <= 0 || (size_t)si.size.x * si.size.y > G.pixelsCap)
    {
        latch(-6, "b200_render", "frame larger than the limits (b200_set_limits before reshape_scene)");
        return;
    }
    if (si.cameraType == B200_CT_VOLUME)
    {
        // k_volumeRenderer's depth-sorted colour list writes one element past its 10-entry array (GeometryIntersections.cuh:1186):
        // undefined behaviour, no defined result to reproduce
        latch(-7, "b200_render", "volume-rendering camera is outside this engine's path: see DESIGN.md scope");
        return;
    }
    if (si.cameraType == B200_CT_VR && G.world > 1)
    {
        latch(-7, "b200_render", "the stereo camera reads its focus depth from one pixel of the frame: render it on one GPU");
        return;
    }
    if (pp.type != B200_PPE_NONE && G.world > 1)
    {
        latch(-8, "b200_render", "post-processing effects gather from neighbouring pixels: render them on one GPU (set_partition world = 1)");
        return;
    }
    if (objects.y > G.nbPrims || objects.w > B200_NB_MAX_LIGHTINFORMATIONS) { latch(-9, "b200_render", "object counts exceed uploaded scene"); return; }
    if (!G.dMats && G.nbPrims > 0) { latch(-10, "b200_render", "materials not uploaded"); return; }

    G.frameW = si.size.x; G.frameH = si.size.y;
    RenderParams P;
    memset(&P, 0, sizeof(P));
    P.scene.boxes = G.dBoxes; P.scene.nbBoxes = G.nbBoxes;
    P.scene.geo = G.dGeo; P.scene.meta = G.dMeta; P.scene.prims = G.dPrims; P.scene.nbPrimitives = objects.y;
    P.scene.mats = G.dMats; P.scene.lights = G.dLights; P.scene.lightInfoSize = objects.w; P.scene.nbLamps = objects.z;
    P.scene.tex = G.dTex; P.scene.randoms = G.dRandoms; P.scene.randomTableSize = G.maxW * G.maxH;
    P.scene.wnodes = G.dWide; P.scene.leafRecs = G.dLeafRecs; P.scene.nbWide = g_useWide ? G.nbWide : 0;
    P.scene.primLeaf = G.dPrimLeaf; P.scene.primRecs = G.dPrimRecs;
    P.scene.uwnodes = G.dUWide; P.scene.nbUWide = (g_useWide && g_useUnordered) ? G.nbUWide : 0; P.scene.opaqueShadows = G.opaqueShadows;
    if (P.scene.nbUWide > 0 && (G.uwDirty || !G.dUWideCH))
    {
        const size_t recs = (size_t)(G.nbUWide + G.nbUX) * (UW_WIDTH / 4);
        if (8 * recs + 2 > G.capUWideCH)
        {
            CK(cudaStreamSynchronize(G.stream));
            freeDev(G.dUWideCH);
            G.capUWideCH = 8 * recs + 1024;
            CK(cudaMalloc(&G.dUWideCH, G.capUWideCH * sizeof(float4)));
        }
        k_nodes_centre_half<<<(unsigned int)((recs + 255) / 256), 256, 0, G.stream>>>(G.dUWide, G.dUWideCH, (int)recs);
        G.launches++;
        G.uwDirty = false;
    }
    P.scene.uwch = G.dUWideCH;
    P.scene.nbUX = g_useBackward ? G.nbUX : 0;
    P.scene.ugnodes = G.dUGroup;
#if UW_GROUP
    if (P.scene.nbUWide > 0)
    {
        // candidate lists of the bounce rays of the group walk: one per ray slot of every warp a launch can hold
        const size_t want = (size_t)G.numSMs * 16 * (CTA_THREADS / 32) * 32 * GW_GATHER_CAP;
        if (want > G.capGatherScratch)
        {
            CK(cudaStreamSynchronize(G.stream));
            freeDev(G.dGatherScratch);
            CK(cudaMalloc(&G.dGatherScratch, want * sizeof(float4)));
            G.capGatherScratch = want;
        }
        P.gatherScratch = G.dGatherScratch;
    }
#endif
    P.scene.rawBoxes = G.dRawBoxes; P.scene.nbRawBoxes = objects.x < G.nbBoxesIn ? objects.x : G.nbBoxesIn;
    P.si = si; P.pp = pp;
    P.eye = make_float3(origin.x, origin.y, origin.z);
    P.target = make_float3(direction.x, direction.y, direction.z);
    P.angles = make_float4(angles.x, angles.y, angles.z, angles.w);
    // the frame's destination: the host buffers named by b200_stream_target (through this GPU's own frame), else the root GPU's frame, else this GPU's
    P.post = G.dPost; P.ids = G.dIds; P.bitmap = (G.dPeerBitmap && !G.mirrorExplicit) ? G.dPeerBitmap : G.dBitmap;
    P.tileCounter = G.dTileCounter; P.workCounters = G.dWork;
    P.tilesX = (si.size.x + TILE_W - 1) / TILE_W;
    P.tilesY = (si.size.y + TILE_H - 1) / TILE_H;
    const int nbTiles = P.tilesX * P.tilesY;
    P.rank = G.rank; P.worldSize = G.world;
    P.fuseTailPercent = g_fuseTailPercent;
    P.packetMask = (G.boxLayoutUsed == 2) ? g_packetMask : 0; // packets need the ordered BVH (only leaf tests observable)
    P.nbLocalTiles = (nbTiles - G.rank + G.world - 1) / G.world;
    P.tileOrder = nullptr;
    if (g_tileOrder == 1 && P.nbLocalTiles > 0)
    {
        // this GPU's tiles (t % world == rank, as ever) sorted along a Z-order curve over the tile grid
        if (G.tileOrderKey[0] != P.tilesX || G.tileOrderKey[1] != P.tilesY || G.tileOrderKey[2] != G.rank || G.tileOrderKey[3] != G.world)
        {
            std::vector<std::pair<unsigned long long, int>> order;
            order.reserve(P.nbLocalTiles);
            auto spread = [](unsigned int v) { unsigned long long x = v; x = (x | (x << 16)) & 0x0000FFFF0000FFFFull; x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
                                               x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full; x = (x | (x << 2)) & 0x3333333333333333ull; x = (x | (x << 1)) & 0x5555555555555555ull; return x; };
            for (int t = G.rank; t < nbTiles; t += G.world) order.push_back({spread(t % P.tilesX) | (spread(t / P.tilesX) << 1), t});
            std::sort(order.begin(), order.end());
            std::vector<int> table(order.size());
            for (size_t i = 0; i < order.size(); ++i) table[i] = order[i].second;
            if (table.size() > G.capTileOrder) { freeDev(G.dTileOrder); CK(cudaMalloc(&G.dTileOrder, table.size() * sizeof(int))); G.capTileOrder = table.size(); }
            CK(cudaMemcpyAsync(G.dTileOrder, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice, G.stream));
            CK(cudaStreamSynchronize(G.stream));
            G.tileOrderKey[0] = P.tilesX; G.tileOrderKey[1] = P.tilesY; G.tileOrderKey[2] = G.rank; G.tileOrderKey[3] = G.world;
        }
        P.tileOrder = G.dTileOrder;
    }

    // staged rendering where a pixel owns one ray tree and more than one pass can happen
    int maxIteration = (si.graphicsLevel < B200_GL_REFLECTIONS) ? 1 : si.nbRayIterations + si.pathTracingIteration;
    maxIteration = maxIteration > B200_NB_MAX_ITERATIONS ? B200_NB_MAX_ITERATIONS : maxIteration;
    const bool giRays = (si.advancedIllumination == B200_AI_BASIC || si.advancedIllumination == B200_AI_FULL);
    bool staged = g_useStaged && maxIteration > 1 && !giRays && si.renderBoxes == 0 &&
                        (si.cameraType == B200_CT_PERSPECTIVE || si.cameraType == B200_CT_ORTHOGRAPHIC || si.cameraType == B200_CT_ANAGLYPH);
    const int eyes = si.cameraType == B200_CT_ANAGLYPH ? 2 : 1; // paths per pixel
    if (staged)
    {
        const size_t eyeStride = (((size_t)P.nbLocalTiles * 32) + 63) & ~(size_t)63;
        const size_t stride = eyeStride * eyes;
        const size_t wantBytes = stride * (size_t)(maxIteration > G.pathIterations ? maxIteration : G.pathIterations);
        if (G.pathFailedBytes && wantBytes >= G.pathFailedBytes) staged = false;
        else if (stride > G.pathStride || maxIteration > G.pathIterations)
        {
            CK(cudaStreamSynchronize(G.stream));
            freeDev(G.dPathWords); freeDev(G.dPathColors); freeDev(G.dPathContrib); freeDev(G.dPathQueues);
            G.pathStride = stride > G.pathStride ? stride : G.pathStride;
            G.pathIterations = maxIteration > G.pathIterations ? maxIteration : G.pathIterations;
            // path state is ~0.3 KB per pixel and pass: if the device cannot hold it (very large frames next to a large scene),
            // this frame size is rendered by the single kernel, which needs none
            const bool ok = cudaMalloc(&G.dPathWords, PATH_WORDS * G.pathStride * sizeof(float)) == cudaSuccess &&
                            cudaMalloc(&G.dPathColors, (size_t)G.pathIterations * G.pathStride * sizeof(float4)) == cudaSuccess &&
                            cudaMalloc(&G.dPathContrib, (size_t)G.pathIterations * G.pathStride * sizeof(float)) == cudaSuccess &&
                            cudaMalloc(&G.dPathQueues, ((size_t)G.pathIterations + 2) * G.pathStride * sizeof(int)) == cudaSuccess; // + reflected-ray queue, shadow queue
            if (!ok)
            {
                cudaGetLastError();
                freeDev(G.dPathWords); freeDev(G.dPathColors); freeDev(G.dPathContrib); freeDev(G.dPathQueues);
                G.pathStride = 0; G.pathIterations = 0;
                G.pathFailedBytes = wantBytes;
                staged = false;
                fprintf(stderr, "solr_b200: no device memory for the staged renderer's path state at this frame size; using the single kernel\n");
            }
        }
    }
    if (staged)
    {
        if (!G.dQueueCounters) CK(cudaMalloc(&G.dQueueCounters, QUEUE_COUNTERS * sizeof(unsigned int)));
        P.pathWords = G.dPathWords; P.pathColors = G.dPathColors; P.pathContributions = G.dPathContrib; P.pathQueues = G.dPathQueues;
        P.queueCounters = G.dQueueCounters; P.pathStride = G.pathStride; P.maxIteration = maxIteration;
        P.eyeStride = (((size_t)P.nbLocalTiles * 32) + 63) & ~(size_t)63;
    }

    const bool fused = staged && g_useStaged == 2 && P.scene.nbUWide > 0 && eyes == 1 && P.packetMask == 0 && !UW_GROUP;
    P.fusedQueues = fused ? 1 : 0;
    // Streamed output: the reader took the previous frame into pinned buffers (b200_d2h_bitmap armed them: host == device right now)
    // and this frame is the staged kernels' alone (no effect pass, no other GPU's pixels) in whole tiles of RGB
    // (or the buffers were named outright, b200_stream_target: then every frame goes there, whole after the kernels if need be)
    const bool sizeOk = (size_t)si.size.x * si.size.y == G.mirrorPixels;
    const bool stale = G.mirrorExplicit && ((G.mirrorBitmap && !G.validBitmap) || (G.mirrorIds && !G.validIds)); // a frame of another size came between
    const bool tiled = staged && !fused && pp.type == B200_PPE_NONE && si.size.x % TILE_W == 0 && si.size.y % TILE_H == 0 &&
                       si.frameBufferType == B200_FT_RGB && sizeOk && !stale;
    const bool whole = G.mirrorExplicit && !tiled && sizeOk; // k_stream_own_tiles after the frame's other kernels
    // (a reader of the pixels alone is better served by the copy: 6 MB at 1080p cost 0.13 ms, the counting kernels 0.26 ms)
    const bool stream = tiled && (G.mirrorExplicit || (g_streamOutput && G.mirrorArmed && G.validIds && G.world == 1 && !G.dPeerBitmap));
    if (whole) { G.validBitmap = G.mirrorBitmap != nullptr; G.validIds = G.mirrorIds != nullptr; }
    const bool wasStreaming = G.streamedBitmap || G.streamedIds;
    G.mirrorArmed = false;
    G.streamedBitmap = (stream || whole) && G.validBitmap; G.streamedIds = (stream || whole) && G.validIds;
    G.validBitmap = G.streamedBitmap; G.validIds = G.streamedIds; // the device is ahead of a host buffer this frame does not write
    if (stream)
    {
        if ((size_t)nbTiles > G.capTileRemaining)
        {
            CK(cudaStreamSynchronize(G.stream));
            freeDev(G.dTileRemaining);
            CK(cudaMalloc(&G.dTileRemaining, (size_t)nbTiles * sizeof(int)));
            G.capTileRemaining = nbTiles;
        }
        // every count returns to zero with its tile's last path; cleared when streaming (re)starts, in case a frame was cut short
        if (!wasStreaming) CK(cudaMemsetAsync(G.dTileRemaining, 0, G.capTileRemaining * sizeof(int), G.stream));
        P.tileRemaining = G.dTileRemaining;
        P.hostBitmap = G.streamedBitmap ? G.mirrorDevBitmap : nullptr; P.hostIds = G.streamedIds ? G.mirrorDevIds : nullptr;
        // measurement only (tools/gpu/gpu_stream_e2e.py): 1 = count tiles but write nothing to the host, 2 = the counting instance of the kernels with nothing to count
        static const int dbg = getenv("SOLR_B200_STREAM_DEBUG") ? atoi(getenv("SOLR_B200_STREAM_DEBUG")) : 0;
        if (dbg >= 1) P.hostBitmap = nullptr, P.hostIds = nullptr;
        if (dbg >= 2) P.tileRemaining = nullptr;
        G.framesStreamed++;
    }
    CK(cudaEventRecord(G.evStart, G.stream));
    CK(cudaMemsetAsync(G.dTileCounter, 0, sizeof(unsigned int), G.stream));
    // the fused driver's consumers recognise a written entry by its being non-zero
    if (fused) CK(cudaMemsetAsync(G.dPathQueues, 0, ((size_t)maxIteration + 1) * G.pathStride * sizeof(int), G.stream));
    const int warpsPerCta = CTA_THREADS / 32;
    int grid = G.numSMs * G.ctasPerSM;
    const int needed = (P.nbLocalTiles + warpsPerCta - 1) / warpsPerCta;
    if (grid > needed) grid = needed > 0 ? needed : 1;
    CK(cudaMemcpyToSymbolAsync(cP, &P, sizeof(P), 0, cudaMemcpyHostToDevice, G.stream));
    if (!staged)
    {
        k_render<<<grid, CTA_THREADS, 0, G.stream>>>();
        G.launches++;
    }
    else
    {
        CK(cudaMemsetAsync(G.dQueueCounters, 0, QUEUE_COUNTERS * sizeof(unsigned int), G.stream));
        if (fused)
        {
            // persistent: every CTA must be resident, or a sleeping warp could wait for one that never starts
            k_stage_fused<<<G.numSMs * G.ctasPerSMStage[5], CTA_THREADS, 0, G.stream>>>();
            k_stage_reflected<false><<<G.numSMs * G.ctasPerSMStage[2], CTA_THREADS, 0, G.stream>>>();
            G.launches += 2;
        }
        else
        {
            int g0 = G.numSMs * G.ctasPerSMStage[0];
            if (g0 > needed) g0 = needed > 0 ? needed : 1;
            if (stream) k_stage_primary<true><<<g0, CTA_THREADS, 0, G.stream>>>();
            else k_stage_primary<false><<<g0, CTA_THREADS, 0, G.stream>>>();
            // queue lengths are only known on the device: persistent grids sized for the device, each warp takes 32 entries at a time
            for (int pass = 1; pass < maxIteration; ++pass)
            {
                if (stream) k_stage_pass<true><<<G.numSMs * G.ctasPerSMStage[1], CTA_THREADS, 0, G.stream>>>(pass);
                else k_stage_pass<false><<<G.numSMs * G.ctasPerSMStage[1], CTA_THREADS, 0, G.stream>>>(pass);
            }
            if (stream) k_stage_reflected<true><<<G.numSMs * G.ctasPerSMStage[2], CTA_THREADS, 0, G.stream>>>();
            else k_stage_reflected<false><<<G.numSMs * G.ctasPerSMStage[2], CTA_THREADS, 0, G.stream>>>();
            G.launches += maxIteration + 1;
        }
    }
    if (pp.type != B200_PPE_NONE)
    {
        const int px = si.size.x * si.size.y;
        int gp = (px + 255) / 256;
        if (gp > G.numSMs * 8) gp = G.numSMs * 8;
        k_post_process<<<gp, 256, 0, G.stream>>>();
        G.launches++;
    }
    if (whole) streamOwnTiles(si.size.x, si.size.y, si.frameBufferType == B200_FT_BGR);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) latch((int)e, "render kernel launch", cudaGetErrorString(e));
    CK(cudaEventRecord(G.evStop, G.stream));
    G.timed = true;
}

void b200_d2h_bitmap(b200_int2, b200_SceneInfo si, b200_BitmapBuffer* bitmap, b200_PrimitiveXYIdBuffer* ids)
{
    if (!G.dBitmap) { latch(-4, "b200_d2h_bitmap", "reshape_scene not called"); return; }
    if (!ensureDevice()) return;
    const size_t px = (size_t)si.size.x * si.size.y;
    if (px > G.pixelsCap) { latch(-6, "b200_d2h_bitmap", "frame larger than the limits"); return; }
    // The caller's buffers are pageable unless their owner pinned them in place (b200_register_host): the engine never pins
    // memory it does not own on its own initiative — a cached registration outlives a freed buffer whose address is reused.
    // Streamed output: what the frame's own kernels wrote into these very buffers needs no copy, only the end of the stream
    const bool haveBitmap = bitmap && G.validBitmap && bitmap == G.mirrorBitmap && px == G.mirrorPixels;
    const bool haveIds = ids && G.validIds && ids == G.mirrorIds && px == G.mirrorPixels;
    if (bitmap && !haveBitmap) CK(cudaMemcpyAsync(bitmap, G.dBitmap, px * 3, cudaMemcpyDeviceToHost, G.stream));
    if (ids && !haveIds) CK(cudaMemcpyAsync(ids, G.dIds, px * 16, cudaMemcpyDeviceToHost, G.stream));
    CK(cudaStreamSynchronize(G.stream));
    // ... and the buffers now hold what the device holds: the pinned ones may be written by the next frame itself
    if (G.mirrorExplicit) return; // the named target stays what it is, whatever else is read
    if (px != G.mirrorPixels) dropMirror();
    G.mirrorPixels = px;
    auto mapped = [](void* p, size_t bytes) -> void* {
        for (auto& r : G.hostRegistered)
            if (r.first == p && r.second >= bytes)
            {
                void* d = nullptr;
                if (cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
                return d;
            }
        return nullptr;
    };
    if (bitmap && !haveBitmap)
    {
        G.mirrorDevBitmap = (unsigned char*)mapped(bitmap, px * 3);
        G.mirrorBitmap = G.mirrorDevBitmap ? bitmap : nullptr;
        G.validBitmap = G.mirrorBitmap != nullptr;
    }
    if (ids && !haveIds)
    {
        G.mirrorDevIds = (int4*)mapped(ids, px * 16);
        G.mirrorIds = G.mirrorDevIds ? ids : nullptr;
        G.validIds = G.mirrorIds != nullptr;
    }
    G.mirrorArmed = true; // somebody reads the frames as they come
}

// Host buffers the caller owns for as long as they stay registered (the frame and id buffers of a host class: GPUKernel.cpp:344-360
// allocates them once): pinned in place so that d2h_bitmap is a straight DMA instead of a staged copy, and mapped so that the ray
// kernels can write them ("streamed output" above).
int b200_register_host(void* p, size_t bytes)
{
    if (!p || !bytes) return -4;
    if (!ensureDevice()) return -1;
    for (auto& r : G.hostRegistered)
        if (r.first == p)
        {
            if (r.second >= bytes) return 0;
            b200_unregister_host(p);
            break;
        }
    if (cudaHostRegister(p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return -12; } // stays pageable: copies are staged
    G.hostRegistered.push_back({p, bytes});
    return 0;
}
int b200_unregister_host(void* p)
{
    for (size_t i = 0; i < G.hostRegistered.size(); ++i)
        if (G.hostRegistered[i].first == p)
        {
            if (p == G.mirrorBitmap || p == G.mirrorIds)
            {
                if ((G.streamedBitmap || G.streamedIds) && ensureDevice()) cudaStreamSynchronize(G.stream); // a frame in flight may be writing it
                if (p == G.mirrorBitmap) { G.mirrorBitmap = nullptr; G.mirrorDevBitmap = nullptr; G.validBitmap = G.streamedBitmap = false; }
                if (p == G.mirrorIds) { G.mirrorIds = nullptr; G.mirrorDevIds = nullptr; G.validIds = G.streamedIds = false; }
            }
            if (ensureDevice() && cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
            G.hostRegistered[i] = G.hostRegistered.back();
            G.hostRegistered.pop_back();
            return 0;
        }
    return -4;
}

unsigned long long b200_frames_streamed(void) { return G.framesStreamed; }

int b200_stream_target(b200_SceneInfo si, b200_BitmapBuffer* bitmap, b200_PrimitiveXYIdBuffer* ids)
{
    if (!ensureDevice()) return -1;
    if (G.streamedBitmap || G.streamedIds) CK(cudaStreamSynchronize(G.stream));
    dropMirror();
    if (!bitmap && !ids) return 0;
    if (!G.dBitmap) { latch(-4, "b200_stream_target", "reshape_scene not called"); return -4; }
    const size_t px = (size_t)si.size.x * si.size.y;
    if (px == 0 || px > G.pixelsCap) { latch(-6, "b200_stream_target", "frame larger than the limits"); return -6; }
    auto mapped = [](void* p, size_t bytes) -> void* {
        for (auto& r : G.hostRegistered)
            if (r.first == p && r.second >= bytes)
            {
                void* d = nullptr;
                if (cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
                return d;
            }
        return nullptr;
    };
    if (bitmap && !(G.mirrorDevBitmap = (unsigned char*)mapped(bitmap, px * 3))) { latch(-12, "b200_stream_target", "frame buffer not registered (b200_register_host)"); return -12; }
    if (ids && !(G.mirrorDevIds = (int4*)mapped(ids, px * 16))) { G.mirrorDevBitmap = nullptr; latch(-12, "b200_stream_target", "id buffer not registered (b200_register_host)"); return -12; }
    G.mirrorBitmap = bitmap; G.mirrorIds = ids; G.mirrorPixels = px;
    G.validBitmap = bitmap != nullptr; G.validIds = ids != nullptr;
    G.mirrorExplicit = true;
    // what this GPU's tiles hold right now
    streamOwnTiles(si.size.x, si.size.y, si.frameBufferType == B200_FT_BGR);
    CK(cudaStreamSynchronize(G.stream));
    return 0;
}

void b200_d2h_primitive_id(b200_SceneInfo si, int x, int y, b200_PrimitiveXYIdBuffer* id)
{
    if (!G.dIds || !id) { latch(-4, "b200_d2h_primitive_id", "reshape_scene not called"); return; }
    if (!ensureDevice()) return;
    if (x < 0 || y < 0 || x >= si.size.x || y >= si.size.y || (size_t)si.size.x * si.size.y > G.pixelsCap) { latch(-6, "b200_d2h_primitive_id", "pixel outside the frame"); return; }
    CK(cudaMemcpyAsync(id, G.dIds + ((size_t)y * si.size.x + x), sizeof(int4), cudaMemcpyDeviceToHost, G.stream));
    CK(cudaStreamSynchronize(G.stream));
}

void b200_debug_counters(unsigned long long* out8)
{
    if (G.dWork && ensureDevice())
    {
        CK(cudaStreamSynchronize(G.stream));
        CK(cudaMemcpyAsync(out8, G.dWork, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, G.stream));
        CK(cudaStreamSynchronize(G.stream));
    }
}

void b200_d2h_post(b200_SceneInfo si, b200_PostProcessingBuffer* post)
{
    if (!G.dPost || !ensureDevice()) { latch(-4, "b200_d2h_post", "reshape_scene not called"); return; }
    const size_t px = (size_t)si.size.x * si.size.y;
    if (px > G.pixelsCap) { latch(-6, "b200_d2h_post", "frame larger than the limits"); return; }
    CK(cudaStreamSynchronize(G.stream));
    CK(cudaMemcpyAsync(post, G.dPost, px * sizeof(b200_PostProcessingBuffer), cudaMemcpyDeviceToHost, G.stream));
    CK(cudaStreamSynchronize(G.stream));
}

void b200_device_buffers(void** bitmap, void** ids, void** post)
{
    if (bitmap) *bitmap = G.dBitmap;
    if (ids) *ids = G.dIds;
    if (post) *post = G.dPost;
}

void b200_get_counters(unsigned long long* rays, unsigned long long* pixels, int reset)
{
    unsigned long long h[2] = {0, 0};
    if (G.dWork && ensureDevice())
    {
        CK(cudaStreamSynchronize(G.stream));
        CK(cudaMemcpyAsync(h, G.dWork, sizeof(h), cudaMemcpyDeviceToHost, G.stream));
        CK(cudaStreamSynchronize(G.stream));
        if (reset) CK(cudaMemsetAsync(G.dWork, 0, 8 * sizeof(unsigned long long), G.stream));
    }
    if (rays) *rays = h[0];
    if (pixels) *pixels = h[1];
}

float b200_last_render_ms(void)
{
    if (!G.timed || !ensureDevice()) return -1.f;
    float ms = -1.f;
    CK(cudaEventSynchronize(G.evStop));
    CK(cudaEventElapsedTime(&ms, G.evStart, G.evStop));
    return ms;
}

unsigned long long b200_kernel_launches(void) { return G.launches; }
int b200_frame_parameter_bytes(void) { return (int)sizeof(RenderParams); }

void b200_scene_upload_stats(float* ms, int* nodes, int* nodesExt, int* gpu)
{
    if (ms) *ms = G.uploadMs;
    if (nodes) *nodes = G.nbUWide;
    if (nodesExt) *nodesExt = G.nbUX;
    if (gpu) *gpu = G.treesOnGpu;
}

void b200_scene_stats(int* in, int* dev, int* prims, int* reserved)
{
    if (in) *in = G.nbBoxesIn;
    if (dev) *dev = G.nbBoxes;
    if (prims) *prims = G.nbPrims;
    if (reserved) *reserved = G.numSMs * G.ctasPerSM;
}

int b200_debug_relayout_boxes(const b200_BoundingBox* boxes, int nbBoxes, float* outPacked, int capacityBoxes)
{
    std::vector<float4> packed;
    const int nOut = relayoutBoxes(boxes, nbBoxes, packed);
    if (outPacked && nOut <= capacityBoxes) memcpy(outPacked, packed.data(), packed.size() * sizeof(float4));
    return nOut;
}

int b200_debug_build_unordered(const b200_BoundingBox* boxes, int nbBoxes, float* outPacked, int capacityBoxes)
{
    std::vector<float4> packed, ubin;
    std::vector<LeafRec> leaves;
    int used = 0;
    relayoutBoxes(boxes, nbBoxes, packed, &used, &leaves);
    if (used != 2 || leaves.empty()) return 0;
    std::vector<int> leafOfNode;
    SahBuilder sb(leaves, ubin, leafOfNode);
    sb.build(0, (int)leaves.size(), 0);
    const int n = (int)(ubin.size() / 2);
    if (outPacked && n <= capacityBoxes) memcpy(outPacked, ubin.data(), ubin.size() * sizeof(float4));
    return n;
}

int b200_debug_build_walk_trees(const b200_BoundingBox* boxes, int nbBoxes, const b200_Primitive* prims, int nbPrims, float* outNodes,
                                int capacityFloat4, int* primLeafOut, int* nbMainOut, int* nbExtOut)
{
    std::vector<float4> packed, uwide;
    std::vector<LeafRec> leaves;
    std::vector<int> primLeaf;
    int used = 0, nbMain = 0, nbExt = 0;
    relayoutBoxes(boxes, nbBoxes, packed, &used, &leaves);
    if (used != 2) return 0;
    buildWalkTrees(leaves, prims, nbPrims, uwide, primLeaf, nbMain, nbExt);
    if (outNodes && (int)uwide.size() <= capacityFloat4) memcpy(outNodes, uwide.data(), uwide.size() * sizeof(float4));
    if (primLeafOut) memcpy(primLeafOut, primLeaf.data(), (size_t)(nbPrims > 0 ? nbPrims : 0) * sizeof(int));
    if (nbMainOut) *nbMainOut = nbMain;
    if (nbExtOut) *nbExtOut = nbExt;
    return (int)uwide.size();
}

void b200_synchronize(void)
{
    if (G.stream && ensureDevice()) CK(cudaStreamSynchronize(G.stream));
}

static int accumulationPixels()
{
    if (!G.dPost || !G.frameW || !G.frameH) { latch(-4, "b200_accumulation_*", "no frame rendered yet"); return 0; }
    return G.frameW * G.frameH;
}
int b200_accumulation_clear(void)
{
    if (!ensureDevice()) return -1;
    const int n = accumulationPixels();
    if (!n) return -4;
    k_accum_clear<<<G.numSMs * 4, 256, 0, G.stream>>>(G.dPost, n);
    G.launches++;
    return (int)cudaGetLastError();
}
int b200_accumulation_export(void* dstFloat4)
{
    if (!ensureDevice()) return -1;
    const int n = accumulationPixels();
    if (!n || !dstFloat4) return -4;
    k_accum_export<<<G.numSMs * 4, 256, 0, G.stream>>>(G.dPost, (float4*)dstFloat4, n);
    G.launches++;
    return (int)cudaGetLastError();
}
int b200_accumulation_import_and_pack(const void* srcFloat4, int iteration)
{
    if (!ensureDevice()) return -1;
    const int n = accumulationPixels();
    if (!n || !srcFloat4) return -4;
    G.validBitmap = false; // the frame is rewritten on the device
    k_accum_import_pack<<<G.numSMs * 4, 256, 0, G.stream>>>(G.dPost, (const float4*)srcFloat4, G.dPeerBitmap ? G.dPeerBitmap : G.dBitmap, n, iteration);
    G.launches++;
    return (int)cudaGetLastError();
}

// FP32 FMA microbenchmark (BASELINE.md 2 asks for a measured value next to the nominal 148 x 128 x 2 x f): every resident lane runs
// dependent FFMA chains, eight independent ones per thread; flops / event time = achieved TFLOP/s, clock64 ticks / event time = the
// SM clock sustained under that load.
int b200_measure_fp32_peak(float* tflops, float* smMhz)
{
    if (!ensureDevice()) return G.err;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, G.device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    float* dOut = nullptr; long long* dCycles = nullptr;
    CK(cudaMalloc(&dOut, (size_t)blocks * threads * sizeof(float)));
    CK(cudaMalloc(&dCycles, (size_t)blocks * sizeof(long long)));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float bestMs = 1e30f;
    for (int rep = 0; rep < 5; ++rep)
    {
        CK(cudaEventRecord(a, 0));
        k_fp32_peak<<<blocks, threads>>>(dOut, iters, dCycles);
        CK(cudaEventRecord(b, 0));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < bestMs) bestMs = ms;
    }
    std::vector<long long> cyc(blocks);
    CK(cudaMemcpy(cyc.data(), dCycles, (size_t)blocks * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (long long c : cyc) mx = c > mx ? c : mx;
    const double flops = (double)blocks * threads * (double)iters * FP32_PEAK_FMAS_PER_ITER * 2.0;
    if (tflops) *tflops = (float)(flops / (bestMs * 1e-3) / 1e12);
    if (smMhz) *smMhz = (float)((double)mx / (bestMs * 1e-3) / 1e6);
    CK(cudaEventDestroy(a)); CK(cudaEventDestroy(b));
    CK(cudaFree(dOut)); CK(cudaFree(dCycles));
    return G.err;
}
}

#ifdef SOLR_DEBUG_SHADOWCMP
extern "C" int b200_debug_shadowcmp(float* out)
{
    int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_dbgN, sizeof(int));
    cudaMemcpyFromSymbol(out, g_dbgRec, sizeof(float) * 64 * 32);
    return n;
}
#endif
