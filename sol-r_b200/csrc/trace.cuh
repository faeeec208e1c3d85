// trace.cuh — device code of the ray-propagation path: box-list walk, primitive tests, shadow walk,
// shading and the bounce loop.  Semantics follow the reference CUDA engine (cited per function, paths
// relative to /root/reference/solr/engines/cuda/); structure does not:
//   * the walk runs over the compact device layout (engine.cu "scene re-layout"): 32 B per box
//     (two float4: min+w0, max+w1), 64 B of hot geometry per primitive (four float4: (p0,size.x)
//     (p1,size.y) (p2,size.z) (n1,0); a sphere test reads only the first) + one packed int of
//     type/material flags, so the inner loops never touch the 128 B Primitive / 176 B Material records;
//   * single-child box chains are collapsed at upload (exact: a child's slab interval is contained in its
//     parent's, so "parent hit and child hit" == "child hit");
//   * hit attributes (normal, triangle areas) are pure functions of (primitive, hit point, ray) and are
//     computed once for the closest hit instead of for every candidate (the reference recomputes them in
//     every primitive test, GeometryIntersections.cuh:261-282,342-344,629-654);
//   * the walk is a while-while loop (find next leaf / test its primitives) so lanes of a warp re-converge
//     between the two phases.
#pragma once
#include "vec.cuh"
#include "../../include/solr_b200_types.h"

// packed per-primitive word: type | fast-transparency class | flags | material id
#define PM_TYPE(m) ((m) & 0xF)
#define PM_FAST(m) (((m) >> 4) & 3)      // Material.attributes.x: 0, 1, 2 = any other value
#define PM_PROCEDURAL(m) (((m) >> 6) & 1) // Material.attributes.y != 0
#define PM_TRANSPARENT(m) (((m) >> 7) & 1) // Material.transparency != 0
#define PM_MATERIAL(m) ((m) >> 8)

struct SceneDev
{
    const float4* __restrict__ boxes; // [2*nbBoxes] (min.xyz, w0) (max.xyz, w1); leaf: w0=start,w1=count; inner: w0=skip,w1=0
    int nbBoxes;
    const float4* __restrict__ geo;   // [4*nbPrimitives] hot geometry, per type (engine.cu packPrimitive)
    const int* __restrict__ meta;     // [nbPrimitives]
    const b200_Primitive* __restrict__ prims; // cold: full records, read once per shaded hit
    int nbPrimitives;
    const b200_Material* __restrict__ mats;
    const b200_LightInformation* __restrict__ lights;
    int lightInfoSize;
    int nbLamps;
    const unsigned char* __restrict__ tex;
    const float* __restrict__ randoms; // randomTableSize + 4 floats (tail zero)
    int randomTableSize;               // the reference's MAX_BITMAP_SIZE
    const b200_BoundingBox* __restrict__ rawBoxes; // uncollapsed list, only walked when renderBoxes != 0
    int nbRawBoxes;
    // 4-wide form of the ordered BVH (engine.cu buildWide): one 128-byte record per inner node holding its (up to) four
    // children's boxes by axis — rows lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] refs[4] — so one dependent load
    // tests four boxes; ref >= 0: inner node, ref < 0: leaf ~ref in leafRecs, INT_MIN: empty slot.
    const float4* __restrict__ wnodes;
    const float4* __restrict__ leafRecs; // [2*nbLeaves] (min.xyz, first primitive) (max.xyz, count): the reference's leaf boxes, in array order
    int nbWide;
    // unordered 4-wide BVH (binned SAH, no ordering constraint) over the same leaf records, for the walks whose result
    // does not depend on the visiting order
    const float4* __restrict__ uwch;    // uwnodes with every child box as centre + half extent (the unit walk's form, engine.cu k_nodes_centre_half)
    const float4* __restrict__ uwnodes; // leaves are single primitives (ref = ~primitive index), boxes are tight per-primitive boxes
    const int* __restrict__ primLeaf;   // [nbPrimitives] reference leaf (index into leafRecs) each primitive belongs to
    int nbUWide;
    // second unordered tree over the grown boxes of cylinders/cones (engine.cu step 2c): the boxes containing a ray's origin are
    // the only primitives that can register a hit behind it
    int nbUX; // its nodes follow the first tree's in uwnodes (root = nbUWide); leaf refs carry bit 30
    // one 96-byte record per primitive for the unit walk (engine.cu step 2d): everything a leaf visit needs behind ONE dependent
    // load — (p0, size.x) (p1, size.y) (p2, size.z) (n1, packed word) (reference leaf box min, leaf number) (leaf box max, original id)
    const float4* __restrict__ primRecs;
    // the unordered trees once more, child-major for the group walk (tracegroup.cuh): 32 bytes per child — (lo.xyz, hi.x) (hi.yz, ref, -)
    const float4* __restrict__ ugnodes;
    int opaqueShadows; // every shadow caster blocks fully (no transparent material, no textured plane): any-hit is exact
};

// The walks are kept out of line by default (one copy each, own register allocation); -DWALK_INLINE=__forceinline__ to compare.
#ifndef WALK_INLINE
#define WALK_INLINE __noinline__
#endif

struct RenderParams
{
    SceneDev scene;
    b200_SceneInfo si;
    b200_PostProcessingInfo pp;
    float3 eye, target;
    float4 angles;
    b200_PostProcessingBuffer* post;
    int4* ids;
    unsigned char* bitmap;
    unsigned int* tileCounter;        // atomic tile queue head
    const int* tileOrder;             // the k-th tile this GPU hands out (null: k * worldSize + rank)
    unsigned long long* workCounters; // [0] rays, [1] pixels
    int tilesX, tilesY, nbLocalTiles;
    int rank, worldSize;
    int fuseTailPercent; // k_stage_pass: queues up to this share of the resident lanes are carried to the end in registers
    int packetMask; // which walks run warp-synchronously: bit0 primary, bit1 secondary, bit2 shadow of primary hits, bit3 other shadow
    // staged rendering (engine.cu "staged kernels"): path state parked between passes, one array per word / per pass
    float* pathWords;          // [PATH_WORDS][pathStride]
    float4* pathColors;        // [maxIteration][pathStride]
    float* pathContributions;  // [maxIteration][pathStride]
    int* pathQueues;           // [maxIteration + 1][pathStride] path slots; queue q feeds pass q, queue maxIteration the reflected-ray stage
    unsigned int* queueCounters; // [2 * (B200_NB_MAX_ITERATIONS + 2)]: entries pushed, entries handed out
    size_t pathStride;
    size_t eyeStride;          // anaglyph: slots of the right eye's paths start here (two paths per pixel)
    int maxIteration;
    int fusedQueues;           // 1: queue entries are slot + 1 in zeroed queues, published behind a fence (engine.cu k_stage_fused)
    float4* gatherScratch; // group walk: candidate lists of the bounce rays while they are walked
    // streamed frame output (engine.cu "streamed output"): the caller's pinned host buffers as the device sees them, and the number
    // of paths every tile still has to end in this frame; all null when the frame is copied after the render instead
    int4* hostIds;
    unsigned char* hostBitmap;
    int* tileRemaining;
};

// One frame's parameters live in constant memory (uploaded on the render stream before the launch):
// every device function reads scene pointers / SceneInfo fields as uniform constant-bank operands, and
// the two out-of-line walks below need no pointer arguments.
__constant__ RenderParams cP;
#define cS (cP.scene)
#define cSI (cP.si)

struct Ray // direction form used inside a walk (GeometryIntersections.cuh:676-679, :36-44)
{
    float3 o, d, nd, inv; // origin, target-origin (not normalised), normalize(d), inverse direction
};

struct Hit
{
    int prim;   // index into the compacted primitive arrays
    float3 p;   // intersection point
    int flags;  // bit0: sphere hit from inside ("back"), bit1: plane normal flipped
};

struct Counters { unsigned int rays; };
// queue ids of the staged renderer (engine.cu): one queue per pass, the reflected-ray stage after them
SB_DEV int passQueue(const int pass) { return pass; }
SB_DEV int reflectedQueue() { return cP.maxIteration; }

SB_DEV void makeRay(Ray& r, float3 origin, float3 dir)
{
    r.o = origin; r.d = dir;
    r.nd = normalize(dir);
    // GeometryIntersections.cuh:38-40: 1.f (not inf) for zero components
    r.inv.x = dir.x != 0.f ? 1.f / dir.x : 1.f;
    r.inv.y = dir.y != 0.f ? 1.f / dir.y : 1.f;
    r.inv.z = dir.z != 0.f ? 1.f / dir.z : 1.f;
}

// GeometryIntersections.cuh:52-79.  Same twelve subtract/multiply results and the same comparisons as the
// reference's sign-indexed corners with early returns, evaluated without branches: a divergent branch
// inside the walk loop would leave the lanes of a warp un-reconverged for the rest of the walk.
// (t0 = 0 at every call site, :690,:818.)
SB_DEV bool slab(const float4 lo, const float4 hi, const Ray& r, const float t1)
{
    const bool sx = r.inv.x < 0.f, sy = r.inv.y < 0.f, sz = r.inv.z < 0.f;
    const float txmin = ((sx ? hi.x : lo.x) - r.o.x) * r.inv.x;
    const float txmax = ((sx ? lo.x : hi.x) - r.o.x) * r.inv.x;
    const float tymin = ((sy ? hi.y : lo.y) - r.o.y) * r.inv.y;
    const float tymax = ((sy ? lo.y : hi.y) - r.o.y) * r.inv.y;
    const float tzmin = ((sz ? hi.z : lo.z) - r.o.z) * r.inv.z;
    const float tzmax = ((sz ? lo.z : hi.z) - r.o.z) * r.inv.z;
    const bool missXY = (txmin > tymax) | (tymin > txmax);
    const float tmin = (tymin > txmin) ? tymin : txmin;
    const float tmax = (tymax < txmax) ? tymax : txmax;
    const bool missZ = (tmin > tzmax) | (tzmin > tmax);
    const float tmin2 = (tzmin > tmin) ? tzmin : tmin;
    const float tmax2 = (tzmax < tmax) ? tzmax : tmax;
    return !missXY & !missZ & (tmin2 < t1) & (tmax2 > 0.f);
}

// ---------------------------------------------------------------------------------------------------
// Primitive tests: "does the ray hit, and where".  Normals are deferred to hitNormal().
// ---------------------------------------------------------------------------------------------------

// GeometryIntersections.cuh:220-257 (sphere), solve part.  g0 = (centre, radius)
SB_DEV bool sphereTest(const float4 g0, const Ray& r, const float eps, float3& I, int& flags)
{
    const float3 O_C = r.o - xyz(g0);
    const float3 dir = r.nd;
    const float a = 2.f * dot(dir, dir);
    const float b = 2.f * dot(O_C, dir);
    const float c = dot(O_C, O_C) - (g0.w * g0.w);
    const float d = b * b - 2.f * a * c;
    if (d <= 0.f || a == 0.f) return false;
    const float rt = sqrtf(d);
    const float t1 = (-b - rt) / a;
    const float t2 = (-b + rt) / a;
    if (t1 <= eps && t2 <= eps) return false;
    float t = 0.f;
    int back = 0;
    if (t1 <= eps) { t = t2; back = 1; }
    else if (t2 <= eps) t = t1;
    else t = (t1 < t2) ? t1 : t2;
    if (t < eps) return false;
    I = r.o + t * dir;
    flags = back;
    return true;
}

// GeometryIntersections.cuh:159-203 (ellipsoid).  size = (g0.w, g1.w, g2.w)
SB_DEV bool ellipsoidTest(const float4 g0, const float3 g1, const Ray& r, const float eps, float3& I)
{
    const float3 O_C = r.o - xyz(g0);
    const float3 dir = r.nd;
    const float a = ((dir.x * dir.x) / (g1.x * g1.x)) + ((dir.y * dir.y) / (g1.y * g1.y)) + ((dir.z * dir.z) / (g1.z * g1.z));
    const float b = ((2.f * O_C.x * dir.x) / (g1.x * g1.x)) + ((2.f * O_C.y * dir.y) / (g1.y * g1.y)) + ((2.f * O_C.z * dir.z) / (g1.z * g1.z));
    const float c = ((O_C.x * O_C.x) / (g1.x * g1.x)) + ((O_C.y * O_C.y) / (g1.y * g1.y)) + ((O_C.z * O_C.z) / (g1.z * g1.z)) - 1.f;
    float d = ((b * b) - (4.f * a * c));
    if (d < 0.f || a == 0.f || b == 0.f || c == 0.f) return false;
    d = sqrtf(d);
    const float t1 = (-b + d) / (2.f * a);
    const float t2 = (-b - d) / (2.f * a);
    if (t1 <= eps && t2 <= eps) return false;
    float t = 0.f;
    if (t1 <= eps) t = t2;
    else if (t2 <= eps) t = t1;
    else t = (t1 < t2) ? t1 : t2;
    if (t < eps) return false;
    I = r.o + t * dir;
    return true;
}

// GeometryIntersections.cuh:293-340 (cylinder) == :358-405 (cone).
// g0 = (p0, size.x) g1 = (p1, size.y) g3 = (n1 axis, -)
SB_DEV bool cylinderTest(const float4 g0, const float4 g1, const float4 g3, const Ray& r, const float eps, float3& I)
{
    const float3 p0 = xyz(g0), p1 = xyz(g1), n1 = xyz(g3);
    const float3 O_C = r.o - p0;
    const float3 dir = r.d;
    float3 n = cross(dir, n1);
    const float ln = length(n);
    if ((ln < eps) && (ln > -eps)) return false;
    n = normalize(n);
    const float d = fabsf(dot(O_C, n));
    if (d > g1.w) return false;
    float3 O = cross(O_C, n1);
    const float t = -dot(O, n) / ln;
    if (t < 0.f) return false;
    O = normalize(cross(n, n1));
    const float s = fabsf(sqrtf(g0.w * g0.w - d * d) / dot(dir, O));
    const float t1 = t - s;
    const float t2 = t + s;
    I = r.o + t1 * dir;
    float3 HB1 = I - p0, HB2 = I - p1;
    float scale1 = dot(HB1, n1), scale2 = dot(HB2, n1);
    if (scale1 < eps || scale2 > eps)
    {
        I = r.o + t2 * dir;
        HB1 = I - p0; HB2 = I - p1;
        scale1 = dot(HB1, n1); scale2 = dot(HB2, n1);
        if (scale1 < eps || scale2 > eps) return false;
    }
    return true;
}

// GeometryIntersections.cuh:575-626 (triangle), up to the hit point.  The `a + b > 1` branch of the
// reference builds E21 = p1 - p1 = 0 (:604) so det_ = 0 and it always rejects (:607-608).
SB_DEV bool triangleTest(const float4 g0, const float4 g1, const float4 g2, const Ray& r, const float eps, float3& I)
{
    const float3 p0 = xyz(g0), p1 = xyz(g1), p2 = xyz(g2);
    const float3 E01 = p1 - p0;
    const float3 E03 = p2 - p0;
    const float3 P = cross(r.d, E03);
    const float det = dot(E01, P);
    if (fabsf(det) < eps) return false;
    const float3 T = r.o - p0;
    const float a = dot(T, P) / det;
    if (a < 0.f || a > 1.f) return false;
    const float3 Q = cross(T, E01);
    const float b = dot(r.d, Q) / det;
    if (b < 0.f || b > 1.f) return false;
    if ((a + b) > 1.f)
    {
        // det_ = dot(p0 - p1, cross(dir, 0)) = 0 unless a component is non-finite; |0| < eps rejects.
        if (0.f < eps) return false;
    }
    const float t = dot(E03, Q) / det;
    if (t < 0) return false;
    I = r.o + t * r.d;
    return true;
}

SB_DEV bool wireFrameMapping(float x, float y, int width) // TextureMapping.cuh:449-456
{
    const int X = fabsf(x), Y = fabsf(y);
    return (X % 100 <= width) || (Y % 100 <= width);
}

// forward: texture mapping used by textured planes inside the intersection test
__device__ __noinline__ float4 cubeMapping(const b200_Primitive& p, float3 I,
                                           float3& normal, float4& specular, float4& attributes, float4& adv);

// GeometryIntersections.cuh:424-567.  Rare primitive class (the reference's default transparentColor = 0
// makes every plane miss, :561-564); reads the cold records.  Kept out of line so it costs the walk no
// registers.
__device__ __noinline__ bool planeTest(const int primIdx, const Ray& r, float3& I,
                                       int& flags, float& shadowIntensity)
{
    const b200_Primitive& p = cS.prims[primIdx];
    const b200_Material& mat = cS.mats[p.materialId];
    bool collision = false;
    const float reverted = 1.f; // every call site passes reverse = false (:739-740,:861-862)
    float3 normal = f3(p.n0.x, p.n0.y, p.n0.z);
    int flipped = 0;
    const float3 ro = r.o, rd = r.d;
    switch (p.type)
    {
    case B200_PT_MAGICCARPET:
    case B200_PT_CHECKBOARD:
    {
        I.y = p.p0.y;
        const float y = ro.y - p.p0.y;
        if (reverted * rd.y < 0.f && reverted * ro.y > reverted * p.p0.y)
        {
            I.x = ro.x + y * rd.x / -rd.y;
            I.z = ro.z + y * rd.z / -rd.y;
            collision = fabsf(I.x - p.p0.x) < p.size.x && fabsf(I.z - p.p0.z) < p.size.z;
        }
        break;
    }
    case B200_PT_XZPLANE:
    {
        const float y = ro.y - p.p0.y;
        if (reverted * rd.y < 0.f && reverted * ro.y > reverted * p.p0.y)
        {
            I.x = ro.x + y * rd.x / -rd.y; I.y = p.p0.y; I.z = ro.z + y * rd.z / -rd.y;
            collision = fabsf(I.x - p.p0.x) < p.size.x && fabsf(I.z - p.p0.z) < p.size.z;
            if (mat.attributes.z == 2) collision &= wireFrameMapping(I.x, I.z, mat.attributes.w);
        }
        if (!collision && reverted * rd.y > 0.f && reverted * ro.y < reverted * p.p0.y)
        {
            flipped ^= 1;
            I.x = ro.x + y * rd.x / -rd.y; I.y = p.p0.y; I.z = ro.z + y * rd.z / -rd.y;
            collision = fabsf(I.x - p.p0.x) < p.size.x && fabsf(I.z - p.p0.z) < p.size.z;
            if (mat.attributes.z == 2) collision &= wireFrameMapping(I.x, I.z, mat.attributes.w);
        }
        break;
    }
    case B200_PT_YZPLANE:
    {
        const float x = ro.x - p.p0.x;
        if (reverted * rd.x < 0.f && reverted * ro.x > reverted * p.p0.x)
        {
            I.x = p.p0.x; I.y = ro.y + x * rd.y / -rd.x; I.z = ro.z + x * rd.z / -rd.x;
            collision = fabsf(I.y - p.p0.y) < p.size.y && fabsf(I.z - p.p0.z) < p.size.z;
            if (mat.innerIllumination.x != 0.f)
                collision &= int(fabsf(I.z)) % 4000 < 2000 && int(fabsf(I.y)) % 4000 < 2000;
            if (mat.attributes.z == 2) collision &= wireFrameMapping(I.y, I.z, mat.attributes.w);
        }
        if (!collision && reverted * rd.x > 0.f && reverted * ro.x < reverted * p.p0.x)
        {
            flipped ^= 1;
            I.x = p.p0.x; I.y = ro.y + x * rd.y / -rd.x; I.z = ro.z + x * rd.z / -rd.x;
            collision = fabsf(I.y - p.p0.y) < p.size.y && fabsf(I.z - p.p0.z) < p.size.z;
            if (mat.innerIllumination.x != 0.f)
                collision &= int(fabsf(I.z)) % 4000 < 2000 && int(fabsf(I.y)) % 4000 < 2000;
            if (mat.attributes.z == 2) collision &= wireFrameMapping(I.y, I.z, mat.attributes.w);
        }
        break;
    }
    case B200_PT_XYPLANE:
    case B200_PT_CAMERA:
    {
        const float z = ro.z - p.p0.z;
        if (reverted * rd.z < 0.f && reverted * ro.z > reverted * p.p0.z)
        {
            I.z = p.p0.z; I.x = ro.x + z * rd.x / -rd.z; I.y = ro.y + z * rd.y / -rd.z;
            collision = fabsf(I.x - p.p0.x) < p.size.x && fabsf(I.y - p.p0.y) < p.size.y;
            if (mat.attributes.z == 2) collision &= wireFrameMapping(I.x, I.y, mat.attributes.w);
        }
        if (!collision && reverted * rd.z > 0.f && reverted * ro.z < reverted * p.p0.z)
        {
            flipped ^= 1;
            I.z = p.p0.z; I.x = ro.x + z * rd.x / -rd.z; I.y = ro.y + z * rd.y / -rd.z;
            collision = fabsf(I.x - p.p0.x) < p.size.x && fabsf(I.y - p.p0.y) < p.size.y;
            if (mat.attributes.z == 2) collision &= wireFrameMapping(I.x, I.y, mat.attributes.w);
        }
        break;
    }
    }
    if (collision)
    {
        shadowIntensity = 1.f;
        float4 color = f4(mat.color.x, mat.color.y, mat.color.z, mat.color.w);
        if (p.type == B200_PT_CAMERA || mat.textureIds.x != B200_TEXTURE_NONE)
        {
            if (flipped) normal = -normal;
            float4 specular = f4(0.f, 0.f, 0.f, 0.f), attributes = f4(0.f, 0.f, 0.f, 0.f), adv = f4(0.f, 0.f, 0.f, 0.f);
            color = cubeMapping(p, I, normal, specular, attributes, adv);
            shadowIntensity = color.w;
        }
        if ((color.x + color.y + color.z) / 3.f >= cSI.transparentColor) collision = false;
    }
    flags = flipped << 1;
    return collision;
}

// One primitive of a leaf: dispatch on type (GeometryIntersections.cuh:712-747).  Returns hit + point.
#ifndef PRIMTEST_INLINE
#define PRIMTEST_INLINE __forceinline__
#endif
__device__ PRIMTEST_INLINE bool primitiveTest(const int idx, const int meta, const Ray& r, float3& I,
                          int& flags, float& planeShadow)
{
    const float eps = cSI.geometryEpsilon;
    const float4* g = cS.geo + 4 * (size_t)idx;
    flags = 0;
    if (!cSI.extendedGeometry)
        return triangleTest(__ldg(g), __ldg(g + 1), __ldg(g + 2), r, eps, I);
    switch (PM_TYPE(meta))
    {
    case B200_PT_ENVIRONMENT:
    case B200_PT_SPHERE: return sphereTest(__ldg(g), r, eps, I, flags);
    case B200_PT_CYLINDER:
    case B200_PT_CONE: return cylinderTest(__ldg(g), __ldg(g + 1), __ldg(g + 3), r, eps, I);
    case B200_PT_ELLIPSOID:
    {
        const float4 g0 = __ldg(g);
        return ellipsoidTest(g0, f3(g0.w, __ldg(g + 1).w, __ldg(g + 2).w), r, eps, I);
    }
    case B200_PT_TRIANGLE: return triangleTest(__ldg(g), __ldg(g + 1), __ldg(g + 2), r, eps, I);
    default: return planeTest(idx, r, I, flags, planeShadow);
    }
}

// Normal (and triangle areas) of an accepted hit — the tail of each reference test:
// sphere :261-277, ellipsoid :205-210, cylinder/cone :342-344, plane :431,:464, triangle :629-654.
SB_DEV void hitNormal(const int idx, const int meta, const float3 I, const int flags, const float3 rayNd, float3& normal,
                      float3& areas)
{
    const float4* g = cS.geo + 4 * (size_t)idx;
    areas = f3(0.f, 0.f, 0.f);
    const int type = cSI.extendedGeometry ? PM_TYPE(meta) : B200_PT_TRIANGLE;
    switch (type)
    {
    case B200_PT_ENVIRONMENT:
    case B200_PT_SPHERE:
    {
        const float4 g0 = __ldg(g);
        if (!PM_PROCEDURAL(meta))
            normal = I - xyz(g0);
        else
        {
            const float sy = __ldg(g + 1).w, sz = __ldg(g + 2).w; // size.y, size.z
            float3 nc;
            nc.x = g0.x + 0.008f * g0.w * cosf(cSI.timestamp + I.x);
            nc.y = g0.y + 0.008f * sy * sinf(cSI.timestamp + I.y);
            nc.z = g0.z + 0.008f * sz * sinf(cosf(cSI.timestamp + I.z));
            normal = I - nc;
        }
        normal = normalize(normal);
        if (flags & 1) normal *= -1.f;
        break;
    }
    case B200_PT_ELLIPSOID:
    {
        const float4 g0 = __ldg(g);
        const float3 g1 = f3(g0.w, __ldg(g + 1).w, __ldg(g + 2).w);
        normal = I - xyz(g0);
        normal.x = 2.f * normal.x / (g1.x * g1.x);
        normal.y = 2.f * normal.y / (g1.y * g1.y);
        normal.z = 2.f * normal.z / (g1.z * g1.z);
        normal = normalize(normal);
        break;
    }
    case B200_PT_CYLINDER:
    case B200_PT_CONE:
    {
        const float3 n1 = xyz(__ldg(g + 3));
        const float3 V = I - xyz(__ldg(g + 2));
        normal = V - n1 * (dot(V, n1) / dot(n1, n1)); // project(), VectorUtils.cuh:92-95
        normal = normalize(normal);
        break;
    }
    case B200_PT_TRIANGLE:
    {
        const float3 p0 = xyz(__ldg(g)), p1 = xyz(__ldg(g + 1)), p2 = xyz(__ldg(g + 2));
        const b200_Primitive& p = cS.prims[idx];
        const float3 v0 = p0 - I, v1 = p1 - I, v2 = p2 - I;
        areas.x = 0.5f * length(cross(v1, v2));
        areas.y = 0.5f * length(cross(v0, v2));
        areas.z = 0.5f * length(cross(v0, v1));
        normal = normalize((f3(p.n0.x, p.n0.y, p.n0.z) * areas.x + f3(p.n1.x, p.n1.y, p.n1.z) * areas.y +
                            f3(p.n2.x, p.n2.y, p.n2.z) * areas.z) / (areas.x + areas.y + areas.z));
        const float rr = dot(rayNd, normal);
        if (rr > 0.f) normal *= -1.f;
        break;
    }
    default:
    {
        const b200_Primitive& p = cS.prims[idx];
        normal = f3(p.n0.x, p.n0.y, p.n0.z);
        if (flags & 2) normal = -normal;
    }
    }
}

// One walk step over the compacted box list shared by both walks: advance `box` to just past the next
// leaf this ray enters; returns that leaf's primitive range (count = 0: list exhausted).  The body has no
// divergent branch other than the loop exit, so a warp stays converged while its lanes step through
// different boxes.
SB_DEV int nextLeaf(const float4* __restrict__ boxes, const int nbBoxes, int& box, const Ray& r, const float minDistance, int& start)
{
    int count = 0;
    while (box < nbBoxes)
    {
        const float4 lo = __ldg(boxes + 2 * box);
        const float4 hi = __ldg(boxes + 2 * box + 1);
        const int w0 = __float_as_int(lo.w), w1 = __float_as_int(hi.w);
        const bool h = slab(lo, hi, r, minDistance);
        // hit: step into the subtree / leaf (box + 1); miss: jump over it (leaf: 1, inner: w0)
        box += (h | (w1 > 0)) ? 1 : w0;
        if (h & (w1 > 0)) { start = w0; count = w1; break; }
    }
    return count;
}

// GeometryIntersections.cuh:667-772 — closest hit over the compacted box list.  Out of line and by value:
// one copy of the walk serves the bounce loop, the extra reflected ray and the GI ray, with its own
// register allocation.  hit.prim = -1: no hit.
__device__ WALK_INLINE Hit closestHit(const float3 origin, const float3 target, const int iteration, const int currentMaterialId)
{
    Hit hit;
    hit.prim = -1; hit.p = f3(0.f, 0.f, 0.f); hit.flags = 0;
    float minDistance = (iteration < 2) ? cSI.viewDistance : cSI.viewDistance / (iteration + 1);
    Ray r;
    makeRay(r, origin, target - origin);
    const float eps = cSI.geometryEpsilon;
    const float4* __restrict__ boxes = cS.boxes;
    const int* __restrict__ metas = cS.meta;
    const int nbBoxes = cS.nbBoxes;
    int box = 0;
    while (true)
    {
        int start = 0;
        const int count = nextLeaf(boxes, nbBoxes, box, r, minDistance, start);
        if (count == 0) break;
        // the leaf's primitives, in array order (first in order wins ties: strict < at :751)
        for (int k = 0; k < count; ++k)
        {
            const int idx = start + k;
            const int meta = __ldg(metas + idx);
            const int fast = PM_FAST(meta);
            if (fast == 0 || (fast == 1 && currentMaterialId != PM_MATERIAL(meta)))
            {
                float3 I;
                int flags;
                float planeShadow;
                if (primitiveTest(idx, meta, r, I, flags, planeShadow))
                {
                    const float distance = length(I - r.o);
                    if (distance > eps && distance < minDistance)
                    {
                        minDistance = distance;
                        hit.prim = idx; hit.p = I; hit.flags = flags;
                    }
                }
            }
        }
    }
    return hit;
}

// renderBoxes != 0 (debug view, GeometryIntersections.cuh:693-696): every box the ray enters adds its
// material colour / 200; primitives are never tested, so the ray always "misses".  Walks the raw list.
__device__ __noinline__ void boxDebugWalk(const float3 origin, const float3 target,
                                          const int iteration, float4& colorBox)
{
    const float minDistance = (iteration < 2) ? cSI.viewDistance : cSI.viewDistance / (iteration + 1);
    Ray r;
    makeRay(r, origin, target - origin);
    int box = 0;
    while (box < cS.nbRawBoxes)
    {
        const b200_BoundingBox& b = cS.rawBoxes[box];
        const float4 lo = f4(b.parameters[0].x, b.parameters[0].y, b.parameters[0].z, 0.f);
        const float4 hi = f4(b.parameters[1].x, b.parameters[1].y, b.parameters[1].z, 0.f);
        if (slab(lo, hi, r, minDistance))
        {
            const b200_float4 c = cS.mats[b.startIndex % B200_NB_MAX_MATERIALS].color;
            colorBox += f4(c.x, c.y, c.z, c.w) / 200.f;
            ++box;
        }
        else
            box += b.indexForNextBox.x;
    }
}

// GeometryIntersections.cuh:798-908 — shadow walk: accumulates blocker opacity until it saturates.
// Returns (tint.xyz, shadow intensity).
__device__ WALK_INLINE float4 shadowWalk(const float3 lampCenter, const float3 origin, const int lightId, const int iteration, const int objectId)
{
    float result = 0.f;
    float3 color = f3(0.f, 0.f, 0.f);
    Ray r;
    const float3 dirv = lampCenter - origin;
    makeRay(r, origin + normalize(dirv) * cSI.rayEpsilon, dirv);
    const float minDistance = (iteration < 2) ? cSI.viewDistance : cSI.viewDistance / (iteration + 1);
    const float eps = cSI.geometryEpsilon;
    const float shadowLimit = cSI.shadowIntensity;
    const float lenOL = length(r.d);
    const float4* __restrict__ boxes = cS.boxes;
    const int* __restrict__ metas = cS.meta;
    const int nbBoxes = cS.nbBoxes;
    const bool extended = cSI.extendedGeometry != 0;
    int box = 0;
    while (result < shadowLimit)
    {
        int start = 0;
        const int count = nextLeaf(boxes, nbBoxes, box, r, minDistance, start);
        if (count == 0) break;
        for (int k = 0; k < count && result < shadowLimit; ++k)
        {
            const int idx = start + k;
            const int meta = __ldg(metas + idx);
            if (PM_FAST(meta) != 0) continue;
            const int origIndex = __ldg(&cS.prims[idx].index);
            // objectId is a compacted index compared with an original id — as the reference does (:829)
            if (origIndex == lightId || origIndex == objectId) continue;
            const int type = extended ? PM_TYPE(meta) : B200_PT_TRIANGLE;
            // :857-859 ptCamera never shadows; :860-863 ptEnvironment falls to planeIntersection, whose switch has no such case
            if (type == B200_PT_CAMERA || type == B200_PT_ENVIRONMENT) continue;
            float3 I;
            int flags = 0;
            float shadowIntensity = 1.f;
            bool hit = primitiveTest(idx, meta, r, I, flags, shadowIntensity);
            if (hit && type == B200_PT_TRIANGLE && cSI.doubleSidedTriangles)
                hit = false; // dangling else (:643-647): with processingShadows both arms return false
            if (hit)
            {
                const float l = length(I - r.o);
                if (l > eps && l < lenOL)
                {
                    float3 normal = f3(0.f, 0.f, 0.f), areas;
                    const bool transparent = PM_TRANSPARENT(meta);
                    if (transparent) hitNormal(idx, meta, I, flags, r.nd, normal, areas);
                    if (type == B200_PT_SPHERE)
                        shadowIntensity = transparent ? (1.f - fabsf(dot(r.nd, normal))) : 1.f; // :280-281
                    else if (type < B200_PT_CHECKBOARD || type == B200_PT_ELLIPSOID || type == B200_PT_CONE)
                        shadowIntensity = 1.f;
                    float ratio = shadowIntensity * shadowLimit;
                    if (transparent)
                    {
                        const b200_Material& m = cS.mats[PM_MATERIAL(meta)];
                        const float3 O_L = normalize(r.d);
                        const float a = fabsf(dot(O_L, normal));
                        const float rr = (m.transparency == 0.f) ? 1.f : (1.f - m.transparency);
                        ratio *= rr * a;
                        color.x += ratio * (0.3f - 0.3f * m.color.x);
                        color.y += ratio * (0.3f - 0.3f * m.color.y);
                        color.z += ratio * (0.3f - 0.3f * m.color.z);
                    }
                    result += ratio;
                }
            }
        }
    }
    result = fmaxf(0.f, fminf(result, shadowLimit));
    return f4(color.x, color.y, color.z, result);
}

// ---------------------------------------------------------------------------------------------------
// Warp-synchronous ("packet") walks.  The 32 rays of an 8x4-pixel tile pierce almost the same boxes, so the
// warp walks the node list ONCE: every lane tests its own ray against the same node, a vote decides whether
// the warp descends, and at a leaf every lane whose own ray enters the leaf box (own closest distance: the
// exact reference test) tests the same primitive.  Node and primitive loads are warp-uniform (one
// transaction), the primitive type switch is uniform, and control flow is driven by votes only, so there is
// no divergence to re-converge from.  Inner-node votes are conservative by construction (see the ordered-BVH
// note in engine.cu): only leaf-box tests are observable, and those stay per lane.
// All 32 lanes must call these together; `active` = this lane carries a ray.
// ---------------------------------------------------------------------------------------------------
#define FULL_MASK 0xffffffffu

__device__ WALK_INLINE Hit closestHitPacket(const float3 origin, const float3 target, const int iteration, const int currentMaterialId,
                                             const bool active)
{
    Hit hit;
    hit.prim = -1; hit.p = f3(0.f, 0.f, 0.f); hit.flags = 0;
    float minDistance = (iteration < 2) ? cSI.viewDistance : cSI.viewDistance / (iteration + 1);
    Ray r;
    makeRay(r, origin, target - origin);
    const float eps = cSI.geometryEpsilon;
    const float4* __restrict__ boxes = cS.boxes;
    const int* __restrict__ metas = cS.meta;
    const int nbBoxes = cS.nbBoxes;
    int box = 0;
    while (box < nbBoxes)
    {
        const float4 lo = __ldg(boxes + 2 * box);
        const float4 hi = __ldg(boxes + 2 * box + 1);
        const int w0 = __float_as_int(lo.w), w1 = __float_as_int(hi.w);
        const bool h = active & slab(lo, hi, r, minDistance);
        const bool any = __any_sync(FULL_MASK, h);
        if (w1 > 0)
        {
            if (any)
            {
                for (int k = 0; k < w1; ++k)
                {
                    const int idx = w0 + k;
                    const int meta = __ldg(metas + idx);
                    const int fast = PM_FAST(meta);
                    if (h && (fast == 0 || (fast == 1 && currentMaterialId != PM_MATERIAL(meta))))
                    {
                        float3 I;
                        int flags;
                        float planeShadow;
                        if (primitiveTest(idx, meta, r, I, flags, planeShadow))
                        {
                            const float distance = length(I - r.o);
                            if (distance > eps && distance < minDistance)
                            {
                                minDistance = distance;
                                hit.prim = idx; hit.p = I; hit.flags = flags;
                            }
                        }
                    }
                }
            }
            box += 1;
        }
        else
            box += any ? 1 : w0;
    }
    return hit;
}

__device__ WALK_INLINE float4 shadowWalkPacket(const float3 lampCenter, const float3 origin, const int lightId, const int iteration,
                                                const int objectId, const bool active)
{
    float result = 0.f;
    float3 color = f3(0.f, 0.f, 0.f);
    Ray r;
    const float3 dirv = lampCenter - origin;
    makeRay(r, origin + normalize(dirv) * cSI.rayEpsilon, dirv);
    const float minDistance = (iteration < 2) ? cSI.viewDistance : cSI.viewDistance / (iteration + 1);
    const float eps = cSI.geometryEpsilon;
    const float shadowLimit = cSI.shadowIntensity;
    const float lenOL = length(r.d);
    const float4* __restrict__ boxes = cS.boxes;
    const int* __restrict__ metas = cS.meta;
    const int nbBoxes = cS.nbBoxes;
    const bool extended = cSI.extendedGeometry != 0;
    int box = 0;
    // a lane drops out once its shadow saturates (:815,:821); the warp walks on while any lane is still open
    while (box < nbBoxes && __any_sync(FULL_MASK, active && result < shadowLimit))
    {
        const float4 lo = __ldg(boxes + 2 * box);
        const float4 hi = __ldg(boxes + 2 * box + 1);
        const int w0 = __float_as_int(lo.w), w1 = __float_as_int(hi.w);
        const bool h = active & (result < shadowLimit) & slab(lo, hi, r, minDistance);
        const bool any = __any_sync(FULL_MASK, h);
        if (w1 > 0)
        {
            if (any)
            {
                for (int k = 0; k < w1; ++k)
                {
                    const int idx = w0 + k;
                    const int meta = __ldg(metas + idx);
                    const int origIndex = __ldg(&cS.prims[idx].index);
                    const int type = extended ? PM_TYPE(meta) : B200_PT_TRIANGLE;
                    bool test = h && result < shadowLimit && PM_FAST(meta) == 0 && origIndex != lightId && origIndex != objectId &&
                                type != B200_PT_CAMERA && type != B200_PT_ENVIRONMENT;
                    if (test)
                    {
                        float3 I;
                        int flags = 0;
                        float shadowIntensity = 1.f;
                        bool hitp = primitiveTest(idx, meta, r, I, flags, shadowIntensity);
                        if (hitp && type == B200_PT_TRIANGLE && cSI.doubleSidedTriangles) hitp = false;
                        if (hitp)
                        {
                            const float l = length(I - r.o);
                            if (l > eps && l < lenOL)
                            {
                                float3 normal = f3(0.f, 0.f, 0.f), areas;
                                const bool transparent = PM_TRANSPARENT(meta);
                                if (transparent) hitNormal(idx, meta, I, flags, r.nd, normal, areas);
                                if (type == B200_PT_SPHERE)
                                    shadowIntensity = transparent ? (1.f - fabsf(dot(r.nd, normal))) : 1.f;
                                else if (type < B200_PT_CHECKBOARD || type == B200_PT_ELLIPSOID || type == B200_PT_CONE)
                                    shadowIntensity = 1.f;
                                float ratio = shadowIntensity * shadowLimit;
                                if (transparent)
                                {
                                    const b200_Material& m = cS.mats[PM_MATERIAL(meta)];
                                    const float3 O_L = normalize(r.d);
                                    const float a = fabsf(dot(O_L, normal));
                                    const float rr = (m.transparency == 0.f) ? 1.f : (1.f - m.transparency);
                                    ratio *= rr * a;
                                    color.x += ratio * (0.3f - 0.3f * m.color.x);
                                    color.y += ratio * (0.3f - 0.3f * m.color.y);
                                    color.z += ratio * (0.3f - 0.3f * m.color.z);
                                }
                                result += ratio;
                            }
                        }
                    }
                }
            }
            box += 1;
        }
        else
            box += any ? 1 : w0;
    }
    result = fmaxf(0.f, fminf(result, shadowLimit));
    return f4(color.x, color.y, color.z, result);
}

// ---------------------------------------------------------------------------------------------------
// Per-lane walks over the 4-wide ordered BVH.  Children are visited in array order (pushed in reverse, popped
// in order), so leaves are reached in exactly the reference's order; inner tests are conservative by
// construction, the leaf box is re-tested with the reference's arithmetic against the closest distance known
// when the leaf is reached.  Compared with the list walk: one dependent 128-byte load per four boxes, no
// per-box sign selects (the near/far rows are picked once per ray), a third of the loop iterations.
// ---------------------------------------------------------------------------------------------------
#define WIDE_STACK 96
#ifndef WIDE_DEFER
#define WIDE_DEFER 4
#endif
#define WIDE_NONE 0x7fffffff

struct WideRay
{
    int nx, ny, nz, fx, fy, fz; // row indices of the near / far planes for this ray's direction signs
};

SB_DEV void wideRows(WideRay& w, const Ray& r)
{
    const bool sx = r.inv.x < 0.f, sy = r.inv.y < 0.f, sz = r.inv.z < 0.f;
    w.nx = sx ? 3 : 0; w.fx = sx ? 0 : 3;
    w.ny = sy ? 4 : 1; w.fy = sy ? 1 : 4;
    w.nz = sz ? 5 : 2; w.fz = sz ? 2 : 5;
}

// tests the four children of node `n`; returns the first hit child (array order) and pushes the others
SB_DEV int wideStep(const float4* __restrict__ n, const WideRay& w, const Ray& r, const float minDistance, int* stack, int& sp)
{
    const float4 ax = __ldg(n + w.nx), bx = __ldg(n + w.fx);
    const float4 ay = __ldg(n + w.ny), by = __ldg(n + w.fy);
    const float4 az = __ldg(n + w.nz), bz = __ldg(n + w.fz);
    const int4 refs = __ldg(reinterpret_cast<const int4*>(n + 6));
    int next = WIDE_NONE;
#define WIDE_CHILD(C, REF)                                                                                           \
    {                                                                                                                \
        const float tnx = (ax.C - r.o.x) * r.inv.x, tfx = (bx.C - r.o.x) * r.inv.x;                                  \
        const float tny = (ay.C - r.o.y) * r.inv.y, tfy = (by.C - r.o.y) * r.inv.y;                                  \
        const float tnz = (az.C - r.o.z) * r.inv.z, tfz = (bz.C - r.o.z) * r.inv.z;                                  \
        const float tmin = fmaxf(fmaxf(tnx, tny), tnz), tmax = fminf(fminf(tfx, tfy), tfz);                          \
        if ((tmin <= tmax) & (tmin < minDistance) & (tmax > 0.f))                                                    \
        {                                                                                                            \
            if (next != WIDE_NONE) stack[sp++] = next;                                                               \
            next = REF;                                                                                              \
        }                                                                                                            \
    }
    WIDE_CHILD(w, refs.w)
    WIDE_CHILD(z, refs.z)
    WIDE_CHILD(y, refs.y)
    WIDE_CHILD(x, refs.x)
#undef WIDE_CHILD
    return next;
}

__device__ WALK_INLINE Hit closestHitWide(const float3 origin, const float3 target, const int iteration, const int currentMaterialId)
{
    Hit hit;
    hit.prim = -1; hit.p = f3(0.f, 0.f, 0.f); hit.flags = 0;
    float minDistance = (iteration < 2) ? cSI.viewDistance : cSI.viewDistance / (iteration + 1);
    Ray r;
    makeRay(r, origin, target - origin);
    WideRay w;
    wideRows(w, r);
    const float eps = cSI.geometryEpsilon;
    const float4* __restrict__ wnodes = cS.wnodes;
    const float4* __restrict__ leafRecs = cS.leafRecs;
    const int* __restrict__ metas = cS.meta;
    int stack[WIDE_STACK];
    int sp = 0;
    int cur = 0;
    int pend[WIDE_DEFER]; // leaves reached but not yet tested (array order)
    int np = 0;
    while (true)
    {
        // inner phase: walk on past up to WIDE_DEFER leaves (with the closest distance known so far — conservative)
        // so the lanes of a warp switch between the two phases a few times per ray instead of once per leaf
        while (cur != WIDE_NONE && np < WIDE_DEFER)
        {
            if (cur >= 0)
            {
                const int next = wideStep(wnodes + 8 * cur, w, r, minDistance, stack, sp);
                cur = (next != WIDE_NONE) ? next : (sp > 0 ? stack[--sp] : WIDE_NONE);
            }
            else
            {
                pend[np++] = cur;
                cur = (sp > 0) ? stack[--sp] : WIDE_NONE;
            }
        }
        if (np == 0) break;
        // leaf phase, in order: the reference's own test of each leaf box against the closest distance known
        // when that leaf is reached, then its primitives
        for (int j = 0; j < np; ++j)
        {
            const int leaf = ~pend[j];
            const float4 lo = __ldg(leafRecs + 2 * leaf);
            const float4 hi = __ldg(leafRecs + 2 * leaf + 1);
            if (slab(lo, hi, r, minDistance))
            {
                const int start = __float_as_int(lo.w), count = __float_as_int(hi.w);
                for (int k = 0; k < count; ++k)
                {
                    const int idx = start + k;
                    const int meta = __ldg(metas + idx);
                    const int fast = PM_FAST(meta);
                    if (fast == 0 || (fast == 1 && currentMaterialId != PM_MATERIAL(meta)))
                    {
                        float3 I;
                        int flags;
                        float planeShadow;
                        if (primitiveTest(idx, meta, r, I, flags, planeShadow))
                        {
                            const float distance = length(I - r.o);
                            if (distance > eps && distance < minDistance)
                            {
                                minDistance = distance;
                                hit.prim = idx; hit.p = I; hit.flags = flags;
                            }
                        }
                    }
                }
            }
        }
        np = 0;
    }
    return hit;
}

__device__ WALK_INLINE float4 shadowWalkWide(const float3 lampCenter, const float3 origin, const int lightId, const int iteration,
                                             const int objectId)
{
    float result = 0.f;
    float3 color = f3(0.f, 0.f, 0.f);
    Ray r;
    const float3 dirv = lampCenter - origin;
    makeRay(r, origin + normalize(dirv) * cSI.rayEpsilon, dirv);
    WideRay w;
    wideRows(w, r);
    const float minDistance = (iteration < 2) ? cSI.viewDistance : cSI.viewDistance / (iteration + 1);
    const float eps = cSI.geometryEpsilon;
    const float shadowLimit = cSI.shadowIntensity;
    const float lenOL = length(r.d);
    const float4* __restrict__ wnodes = cS.wnodes;
    const float4* __restrict__ leafRecs = cS.leafRecs;
    const int* __restrict__ metas = cS.meta;
    const bool extended = cSI.extendedGeometry != 0;
    int stack[WIDE_STACK];
    int sp = 0;
    int cur = 0;
    int pend[WIDE_DEFER];
    int np = 0;
    while (result < shadowLimit)
    {
        while (cur != WIDE_NONE && np < WIDE_DEFER)
        {
            if (cur >= 0)
            {
                const int next = wideStep(wnodes + 8 * cur, w, r, minDistance, stack, sp);
                cur = (next != WIDE_NONE) ? next : (sp > 0 ? stack[--sp] : WIDE_NONE);
            }
            else
            {
                pend[np++] = cur;
                cur = (sp > 0) ? stack[--sp] : WIDE_NONE;
            }
        }
        if (np == 0) break;
        for (int j = 0; j < np && result < shadowLimit; ++j)
        {
            const int leaf = ~pend[j];
            const float4 lo = __ldg(leafRecs + 2 * leaf);
            const float4 hi = __ldg(leafRecs + 2 * leaf + 1);
            if (!slab(lo, hi, r, minDistance)) continue;
            const int start = __float_as_int(lo.w), count = __float_as_int(hi.w);
            for (int k = 0; k < count && result < shadowLimit; ++k)
            {
                const int idx = start + k;
                const int meta = __ldg(metas + idx);
                if (PM_FAST(meta) != 0) continue;
                const int origIndex = __ldg(&cS.prims[idx].index);
                if (origIndex == lightId || origIndex == objectId) continue;
                const int type = extended ? PM_TYPE(meta) : B200_PT_TRIANGLE;
                if (type == B200_PT_CAMERA || type == B200_PT_ENVIRONMENT) continue;
                float3 I;
                int flags = 0;
                float shadowIntensity = 1.f;
                bool hitp = primitiveTest(idx, meta, r, I, flags, shadowIntensity);
                if (hitp && type == B200_PT_TRIANGLE && cSI.doubleSidedTriangles) hitp = false;
                if (hitp)
                {
                    const float l = length(I - r.o);
                    if (l > eps && l < lenOL)
                    {
                        float3 normal = f3(0.f, 0.f, 0.f), areas;
                        const bool transparent = PM_TRANSPARENT(meta);
                        if (transparent) hitNormal(idx, meta, I, flags, r.nd, normal, areas);
                        if (type == B200_PT_SPHERE)
                            shadowIntensity = transparent ? (1.f - fabsf(dot(r.nd, normal))) : 1.f;
                        else if (type < B200_PT_CHECKBOARD || type == B200_PT_ELLIPSOID || type == B200_PT_CONE)
                            shadowIntensity = 1.f;
                        float ratio = shadowIntensity * shadowLimit;
                        if (transparent)
                        {
                            const b200_Material& m = cS.mats[PM_MATERIAL(meta)];
                            const float3 O_L = normalize(r.d);
                            const float a = fabsf(dot(O_L, normal));
                            const float rr = (m.transparency == 0.f) ? 1.f : (1.f - m.transparency);
                            ratio *= rr * a;
                            color.x += ratio * (0.3f - 0.3f * m.color.x);
                            color.y += ratio * (0.3f - 0.3f * m.color.y);
                            color.z += ratio * (0.3f - 0.3f * m.color.z);
                        }
                        result += ratio;
                    }
                }
            }
        }
        np = 0;
    }
    result = fmaxf(0.f, fminf(result, shadowLimit));
    return f4(color.x, color.y, color.z, result);
}

// ---------------------------------------------------------------------------------------------------
// Order-independent walks.
//
// The reference's closest hit is "the last candidate accepted in array order": a candidate is tested iff its leaf
// box passes (slab hit, t_min < closest-so-far, t_max > 0) and accepted iff eps < distance < closest-so-far
// (GeometryIntersections.cuh:690,749-760).  Box t is in units of |direction| while the closest distance is in
// world units (SURVEY finding 7).  When |direction| >= 1, t_min <= world entry distance <= the distance of any
// hit inside the leaf, so the leaf test can never reject a candidate that would be accepted: the result is simply
// the minimum-distance candidate, ties to the lowest array index — independent of the visiting order.  Primary rays
// (|direction| ~ eye distance) and shadow rays (|direction| = distance to the lamp) are of this kind; they walk an
// unconstrained SAH BVH front to back and cull by the best distance found.  Secondary rays have |direction| =
// 1 - rayEpsilon < 1: for those every candidate along the ray is gathered (any order), sorted by array index and the
// reference's accept/reject sequence is replayed exactly, including its per-leaf t_min < closest-so-far test.
// Shadow walks accumulate blocker opacity until it saturates (:815-905); when every caster is opaque the first
// blocker found — in any order — saturates it, otherwise the ordered walk is used.
// ---------------------------------------------------------------------------------------------------
#define UN_STACK 64
#ifdef SOLR_DEBUG_COUNTERS
#define DBG_ADD(i, v) atomicAdd(cP.workCounters + (i), (unsigned long long)(v))
#define DBG_MAX(i, v) atomicMax(cP.workCounters + (i), (unsigned long long)(v))
#define DBG_DECL(x) x
#else
#define DBG_ADD(i, v)
#define DBG_MAX(i, v)
#define DBG_DECL(x)
#endif
#define GATHER_CAP 48 // candidates kept by a bounce-ray walk; a fuller list sends the ray to the ordered walk (24: 573 rays per frame of config 2, 0.2 ms)

// slab test that also returns t_min (same arithmetic as slab())
SB_DEV bool slabT(const float4 lo, const float4 hi, const Ray& r, const float t1, float& tminOut)
{
    const bool sx = r.inv.x < 0.f, sy = r.inv.y < 0.f, sz = r.inv.z < 0.f;
    const float txmin = ((sx ? hi.x : lo.x) - r.o.x) * r.inv.x;
    const float txmax = ((sx ? lo.x : hi.x) - r.o.x) * r.inv.x;
    const float tymin = ((sy ? hi.y : lo.y) - r.o.y) * r.inv.y;
    const float tymax = ((sy ? lo.y : hi.y) - r.o.y) * r.inv.y;
    const float tzmin = ((sz ? hi.z : lo.z) - r.o.z) * r.inv.z;
    const float tzmax = ((sz ? lo.z : hi.z) - r.o.z) * r.inv.z;
    const bool missXY = (txmin > tymax) | (tymin > txmax);
    const float tmin = (tymin > txmin) ? tymin : txmin;
    const float tmax = (tymax < txmax) ? tymax : txmax;
    const bool missZ = (tmin > tzmax) | (tzmin > tmax);
    const float tmin2 = (tzmin > tmin) ? tzmin : tmin;
    const float tmax2 = (tzmax < tmax) ? tzmax : tmax;
    tminOut = tmin2;
    return !missXY & !missZ & (tmin2 < t1) & (tmax2 > 0.f);
}

// tests the four children of an unordered node against [0, tLimit]; pushes the hits far-to-near (nearest on top).
// These boxes only steer the walk (a hit is re-checked against its reference leaf box with the reference's arithmetic), and
// they are padded (engine.cu step 2c), so the test is written for instruction count: slab distances as one FMA
// b * inv - o * inv (o * inv once per ray; the rounding differs from (b - o) * inv by <= 1e-7 |o| in position, far inside
// the padding), rows loaded at fixed offsets from one address and near / far picked by the direction's sign.
// Traversal stack: the first SM_STACK entries of every lane live in shared memory ([entry][thread], 8 bytes each, so a warp's
// push or pop is one conflict-free access), deeper entries in a thread-local array.  The pop is on the critical path of
// the walk (its result is the next node's address) and thread-local arrays are L1-thrashing global memory here: 32 warps
// x ~1.5 KB of stack + spills per SM do not stay in L1, ncu showed ~4 GB of local-memory traffic reaching DRAM per frame.
#ifndef SM_STACK
#define SM_STACK 12
#endif
#define WALK_THREADS 128
struct WalkStack
{
    int2* sm;   // this thread's column of the shared stack
    int* lref;  // overflow, thread-local
    float* lt;
    SB_DEV void push(const int sp, const int ref, const float t) const
    {
        if (sp < SM_STACK) sm[sp * WALK_THREADS] = make_int2(ref, __float_as_int(t));
        else { lref[sp - SM_STACK] = ref; lt[sp - SM_STACK] = t; }
    }
    SB_DEV void pop(const int sp, int& ref, float& t) const
    {
        if (sp < SM_STACK) { const int2 v = sm[sp * WALK_THREADS]; ref = v.x; t = __int_as_float(v.y); }
        else { ref = lref[sp - SM_STACK]; t = lt[sp - SM_STACK]; }
    }
};

struct NodeRay
{
    float ix, iy, iz;    // inverse direction
    float nox, noy, noz; // -origin * inverse direction
};
SB_DEV void nodeRay(NodeRay& q, const Ray& r)
{
    q.ix = r.inv.x; q.iy = r.inv.y; q.iz = r.inv.z;
    q.nox = -r.o.x * r.inv.x; q.noy = -r.o.y * r.inv.y; q.noz = -r.o.z * r.inv.z;
}
// Node records are loaded with an L2 evict-last policy: the scene (~50 MB) fits the 126 MB L2, but the walks also stream
// gigabytes of thread-local memory (stacks, spills) through it per frame, and without a priority the node lines get
// evicted and the next visit pays DRAM latency (ncu: L2 hit rate 56-64 %).
SB_DEV unsigned long long evictLastPolicy()
{
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); // not volatile: one per walk, hoisted out of the node loop
    return p;
}
SB_DEV float4 ldNode(const float4* p, const unsigned long long policy)
{
    float4 v;
#ifdef NODE_LOAD_PLAIN
    v = __ldg(p);
#else
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(policy));
#endif
    return v;
}
#ifndef UW_WIDTH
#define UW_WIDTH 4 // children per node of the unordered trees: 4 (one 128-byte record) or 8 (two consecutive records; measured slower: 5.8 vs 4.7 ms on config 2)
#endif
#define UW_NODE_F4 (2 * UW_WIDTH) // float4 per node

SB_DEV bool wideStepSorted(const float4* __restrict__ n, const NodeRay& q, const float tLimit, const WalkStack& st, int& sp)
{
    const unsigned long long pol = evictLastPolicy();
    const float4 lx = ldNode(n, pol), ly = ldNode(n + 1, pol), lz = ldNode(n + 2, pol);
    const float4 hx = ldNode(n + 3, pol), hy = ldNode(n + 4, pol), hz = ldNode(n + 5, pol);
    const float4 rf = ldNode(n + 6, pol);
    const int4 refs = make_int4(__float_as_int(rf.x), __float_as_int(rf.y), __float_as_int(rf.z), __float_as_int(rf.w));
    const bool sx = q.ix < 0.f, sy = q.iy < 0.f, sz = q.iz < 0.f;
    float t0, t1, t2, t3;
#define UN_CHILD(C, T)                                                                                               \
    {                                                                                                                \
        const float tnx = __fmaf_rn(sx ? hx.C : lx.C, q.ix, q.nox), tfx = __fmaf_rn(sx ? lx.C : hx.C, q.ix, q.nox);  \
        const float tny = __fmaf_rn(sy ? hy.C : ly.C, q.iy, q.noy), tfy = __fmaf_rn(sy ? ly.C : hy.C, q.iy, q.noy);  \
        const float tnz = __fmaf_rn(sz ? hz.C : lz.C, q.iz, q.noz), tfz = __fmaf_rn(sz ? lz.C : hz.C, q.iz, q.noz);  \
        /* entry clamped to the origin, exit to the bound: one comparison says whether the box is crossed within [0, tLimit] */ \
        const float tmin = fmaxf(fmaxf(fmaxf(tnx, tny), tnz), 0.f), tmax = fminf(fminf(fminf(tfx, tfy), tfz), tLimit);           \
        T = (tmin <= tmax) ? tmin : 3.0e38f;                                                                         \
    }
    UN_CHILD(x, t0) UN_CHILD(y, t1) UN_CHILD(z, t2) UN_CHILD(w, t3)
#undef UN_CHILD
    int r0 = refs.x, r1 = refs.y, r2 = refs.z, r3 = refs.w;
    // sorting network, descending t (farthest first)
#define UN_CSWAP(TA, RA, TB, RB) if (TA < TB) { const float tt = TA; TA = TB; TB = tt; const int rr = RA; RA = RB; RB = rr; }
    // nearest into slot 3 (on top of the stack); a full far-to-near order buys nothing measurable
    UN_CSWAP(t0, r0, t1, r1) UN_CSWAP(t2, r2, t3, r3) UN_CSWAP(t1, r1, t3, r3)
#undef UN_CSWAP
    // each level leaves <= 3 entries behind, so UN_STACK - 4 is only exceeded by a degenerate (very deep) tree: the caller
    // then abandons this walk for the ordered one
    if (sp > UN_STACK - 4) return false;
    if (t0 < 3.0e38f) { st.push(sp, r0, t0); ++sp; }
    if (t1 < 3.0e38f) { st.push(sp, r1, t1); ++sp; }
    if (t2 < 3.0e38f) { st.push(sp, r2, t2); ++sp; }
    if (t3 < 3.0e38f) { st.push(sp, r3, t3); ++sp; }
    return true;
}

// Eight children: the two records of the node, tested four at a time; the nearest hit ends up on top of the stack, the
// others below it in no particular order (a full far-to-near order buys nothing measurable: entries whose entry distance
// has fallen behind the best hit are dropped when popped).
SB_DEV bool wideStep8(const float4* __restrict__ n, const NodeRay& q, const float tLimit, const WalkStack& st, int& sp)
{
    const unsigned long long pol = evictLastPolicy();
    const bool sx = q.ix < 0.f, sy = q.iy < 0.f, sz = q.iz < 0.f;
    float t0, t1, t2, t3, t4, t5, t6, t7;
    int r0, r1, r2, r3, r4, r5, r6, r7;
#define UN_CHILD(C, T)                                                                                               \
    {                                                                                                                \
        const float tnx = __fmaf_rn(sx ? hx.C : lx.C, q.ix, q.nox), tfx = __fmaf_rn(sx ? lx.C : hx.C, q.ix, q.nox);  \
        const float tny = __fmaf_rn(sy ? hy.C : ly.C, q.iy, q.noy), tfy = __fmaf_rn(sy ? ly.C : hy.C, q.iy, q.noy);  \
        const float tnz = __fmaf_rn(sz ? hz.C : lz.C, q.iz, q.noz), tfz = __fmaf_rn(sz ? lz.C : hz.C, q.iz, q.noz);  \
        const float tmin = fmaxf(fmaxf(tnx, tny), tnz), tmax = fminf(fminf(tfx, tfy), tfz);                          \
        T = ((tmin <= tmax) & (tmin <= tLimit) & (tmax > 0.f)) ? tmin : 3.0e38f;                                     \
    }
    {
        const float4 lx = ldNode(n, pol), ly = ldNode(n + 1, pol), lz = ldNode(n + 2, pol);
        const float4 hx = ldNode(n + 3, pol), hy = ldNode(n + 4, pol), hz = ldNode(n + 5, pol);
        const float4 rf = ldNode(n + 6, pol);
        r0 = __float_as_int(rf.x); r1 = __float_as_int(rf.y); r2 = __float_as_int(rf.z); r3 = __float_as_int(rf.w);
        UN_CHILD(x, t0) UN_CHILD(y, t1) UN_CHILD(z, t2) UN_CHILD(w, t3)
    }
    {
        const float4 lx = ldNode(n + 8, pol), ly = ldNode(n + 9, pol), lz = ldNode(n + 10, pol);
        const float4 hx = ldNode(n + 11, pol), hy = ldNode(n + 12, pol), hz = ldNode(n + 13, pol);
        const float4 rf = ldNode(n + 14, pol);
        r4 = __float_as_int(rf.x); r5 = __float_as_int(rf.y); r6 = __float_as_int(rf.z); r7 = __float_as_int(rf.w);
        UN_CHILD(x, t4) UN_CHILD(y, t5) UN_CHILD(z, t6) UN_CHILD(w, t7)
    }
#undef UN_CHILD
    // bubble the nearest into slot 7
#define UN_CSWAP(TA, RA, TB, RB) if (TA < TB) { const float tt = TA; TA = TB; TB = tt; const int rr = RA; RA = RB; RB = rr; }
    UN_CSWAP(t0, r0, t1, r1) UN_CSWAP(t1, r1, t2, r2) UN_CSWAP(t2, r2, t3, r3) UN_CSWAP(t3, r3, t4, r4)
    UN_CSWAP(t4, r4, t5, r5) UN_CSWAP(t5, r5, t6, r6) UN_CSWAP(t6, r6, t7, r7)
#undef UN_CSWAP
    if (sp > UN_STACK - 8) return false;
    if (t0 < 3.0e38f) { st.push(sp, r0, t0); ++sp; }
    if (t1 < 3.0e38f) { st.push(sp, r1, t1); ++sp; }
    if (t2 < 3.0e38f) { st.push(sp, r2, t2); ++sp; }
    if (t3 < 3.0e38f) { st.push(sp, r3, t3); ++sp; }
    if (t4 < 3.0e38f) { st.push(sp, r4, t4); ++sp; }
    if (t5 < 3.0e38f) { st.push(sp, r5, t5); ++sp; }
    if (t6 < 3.0e38f) { st.push(sp, r6, t6); ++sp; }
    if (t7 < 3.0e38f) { st.push(sp, r7, t7); ++sp; }
    return true;
}

SB_DEV bool unorderedStep(const float4* __restrict__ nodes, const int ref, const NodeRay& q, const float tLimit, const WalkStack& st, int& sp)
{
#if UW_WIDTH == 8
    return wideStep8(nodes + (size_t)UW_NODE_F4 * ref, q, tLimit, st, sp);
#else
    return wideStepSorted(nodes + (size_t)UW_NODE_F4 * ref, q, tLimit, st, sp);
#endif
}

// ---------------------------------------------------------------------------------------------------
// Lean node round (round 2).  ncu of the round-1 walk (profiles/r01_ncu_frame_v12_staged.json, SASS view) showed ~185 warp
// instructions per node round, 75-80 % of all instructions of the ray kernels, of which only ~70 test boxes: the rest sorted
// the four children (predicated register moves), pushed each hit child behind its own branch with a shared / local-memory split,
// rebuilt the L2 policy descriptor and popped the child it had just pushed.  Here:
//   * a child's result is ONE integer key: the bits of its clamped entry distance (non-negative floats order like integers) with
//     the child number in the two lowest mantissa bits, KEY_MISS for a miss; min over the four keys = nearest hit child, and its
//     two low bits say which — no sort, no (distance, ref) pairs to move;
//   * the nearest hit child is walked next directly (never pushed), the other hit children are pushed in slot order by
//     predicated stores (only "nearest first" matters, profiles/r01_history.md);
//   * the whole stack lives in shared memory ([entry][thread], 8 bytes: ref + key; a popped entry whose key has fallen behind
//     the bound is dropped without a node visit — ~10 % of the pops of primary rays); a walk that would exceed it takes the
//     ordered walk (degenerate trees only: WALK_STACK 16 against a measured maximum depth of 15 with every hit child pushed);
//   * the L2 evict-last policy descriptor is made once per walk.
// The low key bits perturb the stored entry distance by <= 3 ulp; the bounds it is compared with carry 1e-4 relative slack.
// ---------------------------------------------------------------------------------------------------
#ifndef WALK_STACK
#define WALK_STACK 16
#endif
#define WALK_DONE ((int)0x80000000) // no node, no leaf: the walk is over (never a leaf ref: primitive indices stay below 2^30 - 1)
#define WALK_POP 0x7fffffff         // take the next entry from the stack (never a node index)
#define KEY_MISS 0x7fffffff

struct NodeKeys
{
    int k0, k1, k2, k3;
    int4 refs;
};

// 32 bytes of a node record in one instruction (LDG.E.256, new with sm_100) through the read-only path, kept in L2 with the static
// evict-last priority (only the 256-bit form takes it without a policy operand): 4 loads per 128-byte record instead of 7
#ifndef NODE_LOAD_MODE
#define NODE_LOAD_MODE 0
#endif
SB_DEV void ldNode256(const float4* p, float4& a, float4& b)
{
#if defined(NODE_LOAD_PLAIN) || NODE_LOAD_MODE == 3
    a = __ldg(p); b = __ldg(p + 1);
#else
    unsigned long long t0, t1, t2, t3;
    // (the .v8.f32 spelling of the same load crashes ptxas 12.9 inside these kernels; .v4.b64 is the same instruction)
#if NODE_LOAD_MODE == 0
    asm volatile("ld.global.nc.L2::evict_last.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(t0), "=l"(t1), "=l"(t2), "=l"(t3) : "l"(p));
#elif NODE_LOAD_MODE == 1
    asm volatile("ld.global.nc.L1::evict_last.L2::evict_last.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(t0), "=l"(t1), "=l"(t2), "=l"(t3) : "l"(p));
#elif NODE_LOAD_MODE == 2
    asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(t0), "=l"(t1), "=l"(t2), "=l"(t3) : "l"(p));
#else
    asm volatile("ld.global.L1::evict_last.L2::evict_last.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(t0), "=l"(t1), "=l"(t2), "=l"(t3) : "l"(p));
#endif
    a.x = __uint_as_float((unsigned int)t0); a.y = __uint_as_float((unsigned int)(t0 >> 32));
    a.z = __uint_as_float((unsigned int)t1); a.w = __uint_as_float((unsigned int)(t1 >> 32));
    b.x = __uint_as_float((unsigned int)t2); b.y = __uint_as_float((unsigned int)(t2 >> 32));
    b.z = __uint_as_float((unsigned int)t3); b.w = __uint_as_float((unsigned int)(t3 >> 32));
#endif
}

SB_DEV void nodeKeys(const float4* __restrict__ n, const NodeRay& q, const float tLimit, NodeKeys& o)
{
#ifdef NODE_ADDR_SELECT
    // near / far rows picked by address (rows are 16 bytes apart: lo.x lo.y lo.z hi.x hi.y hi.z): 6 address adds instead of 24 selects
    const char* b = reinterpret_cast<const char*>(n);
    const unsigned int ox = q.ix < 0.f ? 48u : 0u, oy = q.iy < 0.f ? 48u : 0u, oz = q.iz < 0.f ? 48u : 0u;
    const float4 nx = __ldg(reinterpret_cast<const float4*>(b + ox)), fx = __ldg(reinterpret_cast<const float4*>(b + (48u - ox)));
    const float4 ny = __ldg(reinterpret_cast<const float4*>(b + (16u + oy))), fy = __ldg(reinterpret_cast<const float4*>(b + (64u - oy)));
    const float4 nz = __ldg(reinterpret_cast<const float4*>(b + (32u + oz))), fz = __ldg(reinterpret_cast<const float4*>(b + (80u - oz)));
    const float4 rf = __ldg(n + 6);
#define UN_NEAR(A, C) n##A.C
#define UN_FAR(A, C) f##A.C
#else
    float4 lx, ly, lz, hx, hy, hz, rf, pad;
    ldNode256(n, lx, ly); ldNode256(n + 2, lz, hx); ldNode256(n + 4, hy, hz); ldNode256(n + 6, rf, pad);
    const bool sx = q.ix < 0.f, sy = q.iy < 0.f, sz = q.iz < 0.f;
#define UN_NEAR(A, C) (s##A ? h##A.C : l##A.C)
#define UN_FAR(A, C) (s##A ? l##A.C : h##A.C)
#endif
    o.refs = make_int4(__float_as_int(rf.x), __float_as_int(rf.y), __float_as_int(rf.z), __float_as_int(rf.w));
    // a box is crossed within [0, tLimit] iff max(entry, 0) <= min(exit, tLimit): one comparison per child
#define UN_KEY(C, K, J)                                                                                              \
    {                                                                                                                \
        const float tnx = __fmaf_rn(UN_NEAR(x, C), q.ix, q.nox), tfx = __fmaf_rn(UN_FAR(x, C), q.ix, q.nox);         \
        const float tny = __fmaf_rn(UN_NEAR(y, C), q.iy, q.noy), tfy = __fmaf_rn(UN_FAR(y, C), q.iy, q.noy);         \
        const float tnz = __fmaf_rn(UN_NEAR(z, C), q.iz, q.noz), tfz = __fmaf_rn(UN_FAR(z, C), q.iz, q.noz);         \
        const float tmin = fmaxf(fmaxf(fmaxf(tnx, tny), tnz), 0.f), tmax = fminf(fminf(fminf(tfx, tfy), tfz), tLimit); \
        K = (tmin <= tmax) ? ((__float_as_int(tmin) & ~3) | J) : KEY_MISS;                                           \
    }
    UN_KEY(x, o.k0, 0) UN_KEY(y, o.k1, 1) UN_KEY(z, o.k2, 2) UN_KEY(w, o.k3, 3)
#undef UN_KEY
#undef UN_NEAR
#undef UN_FAR
}

// The traversal stack: [entry][thread] in shared memory, addressed by 32-bit shared-window addresses so a push is a predicated
// store and an add.  sp = address of the next free entry of this thread's column.
#define WALK_STACK_STRIDE (WALK_THREADS * 8)
SB_DEV void stackPut(const unsigned int at, const int ref, const int key)
{
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(at), "r"(ref), "r"(key) : "memory");
}
SB_DEV int2 stackGet(const unsigned int at)
{
    int2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(at) : "memory");
    return v;
}

// One node round of a lane: `cur` is a node (>= 0, not WALK_POP) or WALK_POP; afterwards it is the next node, a leaf ref (< 0),
// WALK_POP again (the popped entry had fallen behind the bound) or WALK_DONE.  The bottom entry of the stack is a sentinel
// (WALK_DONE, key 0: never behind the bound), so a pop needs no emptiness test.  spLimit = shared address of entry
// WALK_STACK - 3 of column 0: a round pushes at most three entries.  Returns false if the stack would overflow.
SB_DEV bool walkRound(const float4* __restrict__ nodes, const int nbMain, const NodeRay& q, const float cullT, const unsigned int spLimit,
                      unsigned int& sp, int& cur)
{
    // the pop first, so that a lane coming back from a leaf (or from a node without hits) visits its next node in this same round
    if (cur == WALK_POP)
    {
        sp -= WALK_STACK_STRIDE;
        const int2 e = stackGet(sp);
        cur = (__int_as_float(e.y) > cullT) ? WALK_POP : e.x; // the bound shrank since this entry was pushed
    }
    if (cur >= 0 && cur != WALK_POP)
    {
        if (sp >= spLimit) return false; // before the loads: they all go out together
        NodeKeys h;
        nodeKeys(nodes + (size_t)8 * cur, q, (cur >= nbMain) ? 0.f : cullT, h);
        const int kmin = min(min(h.k0, h.k1), min(h.k2, h.k3));
        const int c = kmin & 3;
        const int ra = (c & 1) ? h.refs.y : h.refs.x, rb = (c & 1) ? h.refs.w : h.refs.z;
        cur = (kmin == KEY_MISS) ? WALK_POP : ((c & 2) ? rb : ra);
        // keys are distinct (child number in the low bits), so "!= kmin" singles out the nearest; all KEY_MISS when nothing is hit
        if ((h.k0 != kmin) & (h.k0 != KEY_MISS)) { stackPut(sp, h.refs.x, h.k0); sp += WALK_STACK_STRIDE; }
        if ((h.k1 != kmin) & (h.k1 != KEY_MISS)) { stackPut(sp, h.refs.y, h.k1); sp += WALK_STACK_STRIDE; }
        if ((h.k2 != kmin) & (h.k2 != KEY_MISS)) { stackPut(sp, h.refs.z, h.k2); sp += WALK_STACK_STRIDE; }
        if ((h.k3 != kmin) & (h.k3 != KEY_MISS)) { stackPut(sp, h.refs.w, h.k3); sp += WALK_STACK_STRIDE; }
    }
    return true;
}

// Start of a walk: sentinel, then the root of the point-query tree if there is one; the root of the walk proper is `cur`.
SB_DEV void walkStart(int2* const stack, const int nbMain, const bool pointQuery, unsigned int& spLimit, unsigned int& sp, int& cur)
{
    const unsigned int base = (unsigned int)__cvta_generic_to_shared(stack);
    spLimit = base + (WALK_STACK - 3) * WALK_STACK_STRIDE;
    sp = base + threadIdx.x * 8;
    stackPut(sp, WALK_DONE, 0); sp += WALK_STACK_STRIDE;
    if (pointQuery) { stackPut(sp, nbMain, 0); sp += WALK_STACK_STRIDE; }
    cur = 0;
}

// A walk deeper than the shared stack (the round-2 form first sent those to the ordered walk: 6.3 ms instead of 4.5 ms per frame of
// config 2 — a 4-wide tree leaves up to three entries behind per level and a few percent of the rays get past 11) moves the
// eight oldest entries of its column to a thread-local buffer and slides the rest down; they come back when the sentinel is
// popped.  Rare, so the node round itself only ever sees the shared stack.
#ifndef WALK_SPILL
#define WALK_SPILL 64 // entries a walk can hold in thread-local memory on top of the shared stack
#endif
SB_DEV bool walkSpill(int2* const stack, int2* const spillBuf, int& nSpill, unsigned int& sp)
{
    if (nSpill + 8 > WALK_SPILL) return false;
    const unsigned int col = (unsigned int)__cvta_generic_to_shared(stack) + threadIdx.x * 8;
    for (int k = 0; k < 8; ++k) spillBuf[nSpill + k] = stackGet(col + (1 + k) * WALK_STACK_STRIDE);
    for (unsigned int at = col + 9 * WALK_STACK_STRIDE; at < sp; at += WALK_STACK_STRIDE)
    {
        const int2 e = stackGet(at);
        stackPut(at - 8 * WALK_STACK_STRIDE, e.x, e.y);
    }
    sp -= 8 * WALK_STACK_STRIDE;
    nSpill += 8;
    return true;
}
SB_DEV void walkUnspill(int2* const stack, const int2* const spillBuf, int& nSpill, unsigned int& sp)
{
    const unsigned int col = (unsigned int)__cvta_generic_to_shared(stack) + threadIdx.x * 8;
    nSpill -= 8;
    for (int k = 0; k < 8; ++k) stackPut(col + (1 + k) * WALK_STACK_STRIDE, spillBuf[nSpill + k].x, spillBuf[nSpill + k].y);
    sp = col + 9 * WALK_STACK_STRIDE; // the sentinel stays at entry 0
}

// One walk for the three order-independent ray classes (one copy of the node loop and of the primitive tests keeps the
// instruction working set small — the megakernel is bound by instruction-cache misses otherwise, profiles/r01_history.md):
//   UW_CLOSEST  |direction| >= 1: minimum distance, ties to the lowest array index; culls by the best distance found
//   UW_GATHER   |direction| <  1: gathers the candidates within GATHER_WINDOW x the closest distance, replays the
//               reference's accept/reject sequence in array order (see below); falls back to the ordered walk on overflow
//   UW_SHADOW   every caster opaque: the first blocker between the point and the lamp saturates the shadow
//
// Which candidates can matter for UW_GATHER: with L = |direction| (0.95 for bounce rays), a leaf is skipped when
// t_min(leaf) >= closest-so-far, i.e. when its world entry distance is >= L * closest-so-far.  A candidate X can change the
// fate of a candidate Y only if Y's leaf entry is within [L*d_X, d_X), so influence only reaches candidates whose distances
// are chained by ratios <= 1/L.  Sorting the candidates by distance, everything beyond the first gap wider than 1/L is
// inert: it can neither win nor hide anything that can.  The walk keeps every candidate within GATHER_WINDOW x the closest
// distance (1.5 = eight links of such a chain, each of which would additionally need a matching array order and leaf
// entry) and, when it is over, checks that no chain of kept candidates climbs from the closest one into the window's outermost
// shell (L * window, window] — if one does, something outside the window could hang on to it and the ray takes the ordered walk
// instead (option key 4 = 0 is that literal form for every ray; tests compare the two).
#define GATHER_WINDOW 1.5f
#define UW_CLOSEST 0
#define UW_GATHER 1
#define UW_SHADOW 2

// forward: ordered fallback
__device__ WALK_INLINE Hit closestHitWide(const float3 origin, const float3 target, const int iteration, const int currentMaterialId);

struct WalkOut { Hit hit; float shadow; };

#ifndef UW_INLINE
#define UW_INLINE __noinline__
#endif
#ifndef UW_SHADOW_TLIMIT
#define UW_SHADOW_TLIMIT 1.001f
#endif
#define PRIM_REC_F4 6 // float4 per primitive record of the unit walk
#if defined(UW_LEGACY) || defined(UW_V1)
__device__ UW_INLINE WalkOut unorderedWalk(const int mode, const float3 rayOrigin, const float3 rayDir, const int iteration,
                                              const int currentMaterialId, const int lightId, const int objectId)
{
    WalkOut out;
    out.hit.prim = -1; out.hit.p = f3(0.f, 0.f, 0.f); out.hit.flags = 0; out.shadow = 0.f;
    const float minDistance0 = (iteration < 2) ? cSI.viewDistance : cSI.viewDistance / (iteration + 1);
    const float shadowLimit = cSI.shadowIntensity;
    if (mode == UW_SHADOW && !(0.f < shadowLimit)) return out;
    Ray r;
    makeRay(r, rayOrigin, rayDir);
    NodeRay q;
    nodeRay(q, r);
    const float eps = cSI.geometryEpsilon;
    const float len2 = dot(r.d, r.d);
    const float invLen = rsqrtf(len2) * 1.0001f; // world distance -> t, with slack so culling stays conservative
    const float lenOL = sqrtf(len2);             // shadow: distance to the lamp (length(O_L), :877)
    const float4* __restrict__ leafRecs = cS.leafRecs;
    const int* __restrict__ metas = cS.meta;
    const bool extended = cSI.extendedGeometry != 0;
#ifdef UW_LEGACY
    __shared__ int2 s_stack[SM_STACK * WALK_THREADS];
    int stackRef[UN_STACK - SM_STACK];
    float stackT[UN_STACK - SM_STACK];
    WalkStack st;
    st.sm = s_stack + threadIdx.x; st.lref = stackRef; st.lt = stackT;
    int sp = 1;
    st.push(0, 0, 0.f);
#else
    __shared__ int2 s_stack[WALK_STACK * WALK_THREADS];
    unsigned int sp, spLimit;
    int cur;
    int2 spillBuf[WALK_SPILL];
    int nSpill = 0;
#endif
    int candIdx[GATHER_CAP], candLeaf[GATHER_CAP];
    float candD[GATHER_CAP], candLeafT[GATHER_CAP];
    int n = 0;
    bool overflow = false;
    float best = minDistance0;   // closest accepted / gathered distance so far
    float window = minDistance0; // UW_GATHER: candidates farther than this are inert
    // entry-t bound for nodes: never beyond the reference's own t_min < closest-so-far test; a shadow blocker lies before the lamp (t ~ 1)
#ifndef UW_SHADOW_TLIMIT
#define UW_SHADOW_TLIMIT 1.001f
#endif
    float cullT = (mode == UW_SHADOW) ? fminf(minDistance0, UW_SHADOW_TLIMIT) : fminf(minDistance0, minDistance0 * invLen);
    if (mode == UW_GATHER) cullT = minDistance0;
    // The node array holds two trees: the walk proper [0, nbUWide) and, behind it, the tree of grown cylinder/cone boxes.  The
    // second is a point query — entry bound 0 keeps exactly the boxes that contain the origin — and only hits behind the origin
    // count there; only hits ahead of it count in the first.
    const float4* __restrict__ nodes = cS.uwnodes;
    const int nbMain = cS.nbUWide;
#ifdef UW_LEGACY
    if (cS.nbUX > 0) { st.push(1, nbMain, -3.0e38f); sp = 2; }
#else
    walkStart(s_stack, nbMain, cS.nbUX > 0, spLimit, sp, cur);
#endif
    bool done = false;
    DBG_DECL(int dbgVisits = 0;)
    while (!done)
    {
#ifdef UW_LEGACY
        int cur = WIDE_NONE;
        while (sp > 0)
        {
            --sp;
            int ref;
            float tEntry;
            st.pop(sp, ref, tEntry);
            if (tEntry > cullT) continue; // the bound shrank since this entry was pushed
            if (ref < 0) { cur = ref; break; }
            DBG_ADD(5, 1); DBG_DECL(++dbgVisits;)
            if (!unorderedStep(nodes, ref, q, (ref >= nbMain) ? 0.f : cullT, st, sp)) { overflow = true; sp = 0; }
        }
        if (overflow) break;
        if (cur == WIDE_NONE) break;
#else
        // node rounds until this lane holds a leaf or has nothing left (nodes and WALK_POP are >= 0, leaves and WALK_DONE < 0)
        while (cur >= 0)
        {
            DBG_ADD(5, 1); DBG_DECL(++dbgVisits;)
            if (!walkRound(nodes, nbMain, q, cullT, spLimit, sp, cur))
            {
                // no room for this node's children: make some (the round is repeated), or give up on a degenerate tree
                DBG_ADD(3, 1);
                if (!walkSpill(s_stack, spillBuf, nSpill, sp)) { overflow = true; break; }
            }
        }
        if (overflow) break;
        if (cur == WALK_DONE)
        {
            if (nSpill == 0) break;
            walkUnspill(s_stack, spillBuf, nSpill, sp);
            cur = WALK_POP;
            continue;
        }
#endif
        // a BVH leaf is one primitive
        {
            const bool behind = ((~cur) & 0x40000000) != 0; // from the point-query tree
            const int idx = (~cur) & 0x3FFFFFFF;
#ifndef UW_LEGACY
            cur = WALK_POP; // whatever happens to this leaf, the next round takes an entry from the stack
#endif
            const int meta = __ldg(metas + idx);
            const int fast = PM_FAST(meta);
            bool test;
            if (mode == UW_SHADOW)
            {
                const int origIndex = __ldg(&cS.prims[idx].index);
                const int type = extended ? PM_TYPE(meta) : B200_PT_TRIANGLE;
                // objectId is a compacted index compared with an original id — as the reference does (:829)
                test = fast == 0 && origIndex != lightId && origIndex != objectId && type != B200_PT_CAMERA && type != B200_PT_ENVIRONMENT &&
                       !(type == B200_PT_TRIANGLE && cSI.doubleSidedTriangles);
            }
            else
                test = fast == 0 || (fast == 1 && currentMaterialId != PM_MATERIAL(meta));
            if (!test) continue;
            float3 I;
            int flags;
            float planeShadow;
            DBG_ADD(7, 1);
            if (!primitiveTest(idx, meta, r, I, flags, planeShadow)) continue;
            const float distance = length(I - r.o);
            if (!(distance > eps)) continue;
            if ((dot(I - r.o, r.d) < 0.f) != behind) continue; // hits behind the origin (cylinders/cones only) come from the point query
            // the reference only tests a primitive whose leaf box passes its slab test (:690); checked for hits only
            const int leaf = __ldg(cS.primLeaf + idx);
            const float4 lo = __ldg(leafRecs + 2 * leaf);
            const float4 hi = __ldg(leafRecs + 2 * leaf + 1);
            float leafT;
            // (UW_CLOSEST: t_min(leaf) <= entry distance <= hit distance < closest-so-far whenever the hit would be accepted, so only the
            //  geometric part of the leaf test can reject it)
            if (!slabT(lo, hi, r, (mode == UW_CLOSEST) ? 3.0e38f : minDistance0, leafT)) continue;
            if (mode == UW_SHADOW)
            {
                if (distance < lenOL) { out.shadow = fmaxf(0.f, fminf(shadowLimit, shadowLimit)); done = true; }
            }
            else if (mode == UW_CLOSEST)
            {
                if (distance < best || (distance == best && out.hit.prim >= 0 && idx < out.hit.prim))
                {
                    best = distance;
                    out.hit.prim = idx; out.hit.p = I; out.hit.flags = flags;
                    cullT = fminf(minDistance0, best * invLen);
                }
            }
            else if (distance < minDistance0 && distance <= window)
            {
                if (distance < best)
                {
                    best = distance;
                    window = fminf(minDistance0, GATHER_WINDOW * best);
                    cullT = fminf(minDistance0, window * invLen);
                }
                if (n == GATHER_CAP)
                {
                    int m2 = 0; // full: drop what fell out of the window meanwhile
                    for (int j = 0; j < n; ++j)
                        if (candD[j] <= window)
                        {
                            candIdx[m2] = candIdx[j]; candD[m2] = candD[j]; candLeafT[m2] = candLeafT[j]; candLeaf[m2] = candLeaf[j];
                            ++m2;
                        }
                    n = m2;
                }
                if (n == GATHER_CAP) { overflow = true; done = true; }
                else
                {
                    // appended in visiting order; the replay below picks them in array order
                    candIdx[n] = idx; candD[n] = distance; candLeafT[n] = leafT; candLeaf[n] = leaf;
                    ++n;
                }
            }
        }
    }
    DBG_ADD(2, 1); DBG_ADD(3, overflow ? 1 : 0); DBG_ADD(4, n); DBG_MAX(6, dbgVisits);
    if (overflow)
    {
        out.hit.prim = -2; out.shadow = -1.f; // caller runs the ordered walk
        return out;
    }
    if (mode != UW_GATHER) return out;
    // replay in array order (selection by ascending index: the list is short and every lane of the warp is here at the
    // same time; a cylinder listed twice by the point query is taken once).  Primitives of one leaf are contiguous and
    // share the leaf's fate, decided when the leaf is reached (before any of its primitives): t_min(leaf) < closest-so-far.
    float m = minDistance0;
    bool leafPass = false;
    int prevLeaf = -1;
    int winner = -1;
    int last = -1;
    for (int pass = 0; pass < n; ++pass)
    {
        int bj = -1, bi = 0x7fffffff;
        for (int j = 0; j < n; ++j)
        {
            const int ci = candIdx[j];
            if (ci > last && ci < bi && candD[j] <= window) { bi = ci; bj = j; }
        }
        if (bj < 0) break;
        last = bi;
        if (candLeaf[bj] != prevLeaf)
        {
            leafPass = candLeafT[bj] < m;
            prevLeaf = candLeaf[bj];
        }
        if (leafPass && candD[bj] < m) { m = candD[bj]; winner = bi; }
    }
    if (winner >= 0)
    {
        const int meta = __ldg(metas + winner);
        float3 I;
        int flags;
        float planeShadow;
        primitiveTest(winner, meta, r, I, flags, planeShadow); // deterministic: same hit point as when it was gathered
        out.hit.prim = winner; out.hit.p = I; out.hit.flags = flags;
    }
    return out;
}

#else
// ---------------------------------------------------------------------------------------------------
// The unit walk (round 2).  What the ncu source view of the round-1 walk showed: a node round takes a warp ~2 500 cycles whatever
// the number of lanes in it (16 for primary rays, 6 for bounce rays) — it is one dependent memory access plus a long dependent
// instruction chain, not issue slots — and a leaf visit was FOUR dependent accesses (packed word -> geometry -> leaf number ->
// leaf box).  So the walk is rebuilt around "one dependent access per step, and every lane takes a step in every iteration":
//   * a lane's current item is a node (128-byte record) or a primitive (96-byte record holding geometry, packed word, its
//     reference leaf's box and number, its original id); every iteration each lane loads ITS item with the same 256-bit loads
//     at the top of the loop, then the node lanes run the node round (keys, nearest child next, the other hit children pushed)
//     and the leaf lanes the primitive test, the reference's leaf-box test and the acceptance rule — both from registers;
//   * a lane that pops a leaf or a node works on it in the same iteration; nobody waits for the other lanes to reach a leaf.
// Same candidates, same acceptance rules, same results as the round-1 walk (tests/test_gpu_parity.py compares it with the
// ordered walks and with the reference CUDA engine).
// ---------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------
// Bulk copies (TMA, cp.async.bulk) for the scene arrays the walks read.
//   * SCENE_L2_PREFETCH: the first kernel of a frame asks the L2 for the node records and the primitive records, a slice per
//     CTA (cp.async.bulk.prefetch.L2): the walks' first touches then meet L2 instead of DRAM latency.
//   * TOP_SMEM = K: every persistent CTA copies the first K node records — the tree is numbered breadth-first, so these are
//     its top levels, which every ray walks — into shared memory once (cp.async.bulk.shared::cluster.global with an mbarrier
//     for completion) and the walk reads nodes < K from there.
// ---------------------------------------------------------------------------------------------------
SB_DEV void bulkPrefetchL2(const void* p, const unsigned int bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// one slice of [base, base + bytes) per CTA, in pieces of at most 64 KiB, multiples of 16 bytes; thread 0 of the CTA issues them
SB_DEV void scenePrefetchSlice(const void* base, const size_t bytes)
{
    if (base == nullptr || bytes == 0) return;
    const size_t per = (((bytes + gridDim.x - 1) / gridDim.x) + 127) & ~(size_t)127;
    size_t at = (size_t)blockIdx.x * per;
    const size_t end = (at + per < bytes) ? at + per : (bytes & ~(size_t)15);
    for (; at < end; at += 65536)
        bulkPrefetchL2(reinterpret_cast<const char*>(base) + at, (unsigned int)((end - at < 65536) ? end - at : 65536));
}
#ifdef TOP_SMEM
__shared__ __align__(128) float4 s_topNodes[TOP_SMEM * 8];
__shared__ __align__(8) unsigned long long s_topBar;
// every thread of the CTA calls, once, before its first walk
SB_DEV void topStage()
{
    const int n = cS.nbUWide < TOP_SMEM ? cS.nbUWide : TOP_SMEM;
    if (n <= 0) return;
    const unsigned int bar = (unsigned int)__cvta_generic_to_shared(&s_topBar);
    if (threadIdx.x == 0)
    {
        const unsigned int dst = (unsigned int)__cvta_generic_to_shared(s_topNodes);
        const unsigned int bytes = (unsigned int)n * 128u;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the barrier's initialisation, seen by the copy engine
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(cS.uwnodes), "r"(bytes), "r"(bar) : "memory");
    }
    __syncthreads();
    unsigned int done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
}
#else
SB_DEV void topStage() {}
#endif

#if !defined(TOP_SMEM) && !defined(UW_LOHI)
#define UW_CENTRE_HALF 1
#endif
#ifdef UW_CENTRE_HALF
// Child boxes as centre c and half extent h (>= the exact one: rounded up): entry and exit along an axis are
// (c - o) inv -/+ h |inv| whatever the sign of the direction — three FMAs per axis and child on the FMA pipe instead of two FMAs
// and two sign selects on the ALU pipe, which is the pipe the node round fills (24 selects of ~70 ALU instructions per round).
// An empty child slot has h = -inf: entry +inf, exit -inf.  Conservative like the lo/hi form (two more roundings of ~1e-7 x the
// distance, against boxes padded by 0.02 + 2e-5 |coordinate|); a NaN (0 x inf on an axis the ray does not move along) leaves
// that axis unconstrained, as before.
SB_DEV void nodeKeysRegs(const float4 cx, const float4 cy, const float4 cz, const float4 hx, const float4 hy, const float4 hz, const float4 rf,
                         const NodeRay& q, const float tLimit, NodeKeys& o)
{
    o.refs = make_int4(__float_as_int(rf.x), __float_as_int(rf.y), __float_as_int(rf.z), __float_as_int(rf.w));
    const float ax = fabsf(q.ix), ay = fabsf(q.iy), az = fabsf(q.iz);
#define UN_KEY(C, K, J)                                                                                              \
    {                                                                                                                \
        const float tcx = __fmaf_rn(cx.C, q.ix, q.nox), tcy = __fmaf_rn(cy.C, q.iy, q.noy), tcz = __fmaf_rn(cz.C, q.iz, q.noz); \
        const float tnx = __fmaf_rn(-hx.C, ax, tcx), tfx = __fmaf_rn(hx.C, ax, tcx);                                 \
        const float tny = __fmaf_rn(-hy.C, ay, tcy), tfy = __fmaf_rn(hy.C, ay, tcy);                                 \
        const float tnz = __fmaf_rn(-hz.C, az, tcz), tfz = __fmaf_rn(hz.C, az, tcz);                                 \
        const float tmin = fmaxf(fmaxf(fmaxf(tnx, tny), tnz), 0.f), tmax = fminf(fminf(fminf(tfx, tfy), tfz), tLimit); \
        K = (tmin <= tmax) ? ((__float_as_int(tmin) & ~3) | J) : KEY_MISS;                                           \
    }
    UN_KEY(x, o.k0, 0) UN_KEY(y, o.k1, 1) UN_KEY(z, o.k2, 2) UN_KEY(w, o.k3, 3)
#undef UN_KEY
}
#else
SB_DEV void nodeKeysRegs(const float4 lx, const float4 ly, const float4 lz, const float4 hx, const float4 hy, const float4 hz, const float4 rf,
                         const NodeRay& q, const float tLimit, NodeKeys& o)
{
    const bool sx = q.ix < 0.f, sy = q.iy < 0.f, sz = q.iz < 0.f;
    o.refs = make_int4(__float_as_int(rf.x), __float_as_int(rf.y), __float_as_int(rf.z), __float_as_int(rf.w));
#define UN_NEAR(A, C) (s##A ? h##A.C : l##A.C)
#define UN_FAR(A, C) (s##A ? l##A.C : h##A.C)
#define UN_KEY(C, K, J)                                                                                              \
    {                                                                                                                \
        const float tnx = __fmaf_rn(UN_NEAR(x, C), q.ix, q.nox), tfx = __fmaf_rn(UN_FAR(x, C), q.ix, q.nox);         \
        const float tny = __fmaf_rn(UN_NEAR(y, C), q.iy, q.noy), tfy = __fmaf_rn(UN_FAR(y, C), q.iy, q.noy);         \
        const float tnz = __fmaf_rn(UN_NEAR(z, C), q.iz, q.noz), tfz = __fmaf_rn(UN_FAR(z, C), q.iz, q.noz);         \
        const float tmin = fmaxf(fmaxf(fmaxf(tnx, tny), tnz), 0.f), tmax = fminf(fminf(fminf(tfx, tfy), tfz), tLimit); \
        K = (tmin <= tmax) ? ((__float_as_int(tmin) & ~3) | J) : KEY_MISS;                                           \
    }
    UN_KEY(x, o.k0, 0) UN_KEY(y, o.k1, 1) UN_KEY(z, o.k2, 2) UN_KEY(w, o.k3, 3)
#undef UN_KEY
#undef UN_NEAR
#undef UN_FAR
}

#endif

// dispatch on type with the geometry already in registers (primitiveTest() loads it)
SB_DEV bool primitiveTestRegs(const float4 g0, const float4 g1, const float4 g2, const float4 g3, const int idx, const int meta, const Ray& r,
                              float3& I, int& flags, float& planeShadow)
{
    const float eps = cSI.geometryEpsilon;
    flags = 0;
    if (!cSI.extendedGeometry) return triangleTest(g0, g1, g2, r, eps, I);
    switch (PM_TYPE(meta))
    {
    case B200_PT_ENVIRONMENT:
    case B200_PT_SPHERE: return sphereTest(g0, r, eps, I, flags);
    case B200_PT_CYLINDER:
    case B200_PT_CONE: return cylinderTest(g0, g1, g3, r, eps, I);
    case B200_PT_ELLIPSOID: return ellipsoidTest(g0, f3(g0.w, g1.w, g2.w), r, eps, I);
    case B200_PT_TRIANGLE: return triangleTest(g0, g1, g2, r, eps, I);
    default: return planeTest(idx, r, I, flags, planeShadow);
    }
}


// the traversal stacks of a CTA's lanes ([entry][thread]); the unit walk's
__shared__ int2 s_walkStack[WALK_STACK * WALK_THREADS];

__device__ UW_INLINE WalkOut unorderedWalk(const int mode, const float3 rayOrigin, const float3 rayDir, const int iteration,
                                              const int currentMaterialId, const int lightId, const int objectId)
{
    WalkOut out;
    out.hit.prim = -1; out.hit.p = f3(0.f, 0.f, 0.f); out.hit.flags = 0; out.shadow = 0.f;
    const float minDistance0 = (iteration < 2) ? cSI.viewDistance : cSI.viewDistance / (iteration + 1);
    const float shadowLimit = cSI.shadowIntensity;
    if (mode == UW_SHADOW && !(0.f < shadowLimit)) return out;
    Ray r;
    makeRay(r, rayOrigin, rayDir);
    NodeRay q;
    nodeRay(q, r);
    const float eps = cSI.geometryEpsilon;
    const float len2 = dot(r.d, r.d);
    const float invLen = rsqrtf(len2) * 1.0001f; // world distance -> t, with slack so culling stays conservative
    const float lenOL = sqrtf(len2);             // shadow: distance to the lamp (length(O_L), :877)
    const bool extended = cSI.extendedGeometry != 0;
    int2* const s_stack = s_walkStack;
    unsigned int sp, spLimit;
    int cur;
    int2 spillBuf[WALK_SPILL];
    int nSpill = 0;
    int candIdx[GATHER_CAP], candLeaf[GATHER_CAP];
    float candD[GATHER_CAP], candLeafT[GATHER_CAP];
    int n = 0;
    bool overflow = false;
    // UW_GATHER: the closest candidate so far (in out.hit), its reference leaf and that leaf's entry distance; whether another
    // candidate lies at exactly its distance
    int bestLeaf = -1;
    float bestLeafT = 0.f;
    bool bestTie = false;
    float best = minDistance0;   // closest accepted / gathered distance so far
    float window = minDistance0; // UW_GATHER: candidates farther than this are inert
    // entry-t bound for nodes: never beyond the reference's own t_min < closest-so-far test; a shadow blocker lies before the lamp (t ~ 1)
    float cullT = (mode == UW_SHADOW) ? fminf(minDistance0, UW_SHADOW_TLIMIT) : fminf(minDistance0, minDistance0 * invLen);
    if (mode == UW_GATHER) cullT = minDistance0;
#ifdef UW_CENTRE_HALF
    const float4* __restrict__ nodes = cS.uwch;
#else
    const float4* __restrict__ nodes = cS.uwnodes;
#endif
    const float4* __restrict__ recs = cS.primRecs;
    const int nbMain = cS.nbUWide;
    walkStart(s_stack, nbMain, cS.nbUX > 0, spLimit, sp, cur);
    DBG_DECL(int dbgVisits = 0;)
    while (true)
    {
        if (cur == WALK_POP)
        {
            sp -= WALK_STACK_STRIDE;
            const int2 e = stackGet(sp);
            cur = (__int_as_float(e.y) > cullT) ? WALK_POP : e.x; // the bound shrank since this entry was pushed
            if (cur == WALK_POP) continue;
        }
        if (cur == WALK_DONE)
        {
            if (nSpill == 0) break;
            walkUnspill(s_stack, spillBuf, nSpill, sp);
            cur = WALK_POP;
            continue;
        }
        // ---- this lane's item: one dependent access, the same instructions for nodes and primitives
        const bool isNode = cur >= 0;
        const int idx = (~cur) & 0x3FFFFFFF;
        const float4* item = isNode ? nodes + (size_t)8 * cur : recs + (size_t)PRIM_REC_F4 * idx;
        float4 a0, a1, a2, a3, a4, a5, a6 = f4(0.f, 0.f, 0.f, 0.f), a7;
#ifdef TOP_SMEM
        if (isNode && cur < TOP_SMEM && cur < nbMain)
        {
            // the top of the tree, staged in shared memory by topStage()
            const float4* t = s_topNodes + 8 * cur;
            a0 = t[0]; a1 = t[1]; a2 = t[2]; a3 = t[3]; a4 = t[4]; a5 = t[5]; a6 = t[6];
        }
        else
#endif
        {
            ldNode256(item, a0, a1); ldNode256(item + 2, a2, a3); ldNode256(item + 4, a4, a5);
            if (isNode) ldNode256(item + 6, a6, a7);
        }
        if (isNode)
        {
            DBG_ADD(5, 1); DBG_DECL(++dbgVisits;)
            if (sp >= spLimit)
            {
                DBG_ADD(3, 1);
                if (!walkSpill(s_stack, spillBuf, nSpill, sp)) { overflow = true; break; } // degenerate tree
            }
            NodeKeys h;
            nodeKeysRegs(a0, a1, a2, a3, a4, a5, a6, q, (cur >= nbMain) ? 0.f : cullT, h);
            const int kmin = min(min(h.k0, h.k1), min(h.k2, h.k3));
            const int c = kmin & 3;
            const int ra = (c & 1) ? h.refs.y : h.refs.x, rb = (c & 1) ? h.refs.w : h.refs.z;
            cur = (kmin == KEY_MISS) ? WALK_POP : ((c & 2) ? rb : ra);
            // keys are distinct (child number in the low bits), so "!= kmin" singles out the nearest; all KEY_MISS when nothing is hit
            if ((h.k0 != kmin) & (h.k0 != KEY_MISS)) { stackPut(sp, h.refs.x, h.k0); sp += WALK_STACK_STRIDE; }
            if ((h.k1 != kmin) & (h.k1 != KEY_MISS)) { stackPut(sp, h.refs.y, h.k1); sp += WALK_STACK_STRIDE; }
            if ((h.k2 != kmin) & (h.k2 != KEY_MISS)) { stackPut(sp, h.refs.z, h.k2); sp += WALK_STACK_STRIDE; }
            if ((h.k3 != kmin) & (h.k3 != KEY_MISS)) { stackPut(sp, h.refs.w, h.k3); sp += WALK_STACK_STRIDE; }
            continue;
        }
        // ---- a BVH leaf is one primitive
        const bool behind = ((~cur) & 0x40000000) != 0; // from the point-query tree
        cur = WALK_POP;                                 // whatever happens to this leaf, the next thing is the stack
        const int meta = __float_as_int(a3.w);
        const int fast = PM_FAST(meta);
        bool test;
        if (mode == UW_SHADOW)
        {
            const int origIndex = __float_as_int(a5.w);
            const int type = extended ? PM_TYPE(meta) : B200_PT_TRIANGLE;
            // objectId is a compacted index compared with an original id — as the reference does (:829)
            test = fast == 0 && origIndex != lightId && origIndex != objectId && type != B200_PT_CAMERA && type != B200_PT_ENVIRONMENT &&
                   !(type == B200_PT_TRIANGLE && cSI.doubleSidedTriangles);
        }
        else
            test = fast == 0 || (fast == 1 && currentMaterialId != PM_MATERIAL(meta));
        if (!test) continue;
        float3 I;
        int flags;
        float planeShadow;
        DBG_ADD(7, 1);
        if (!primitiveTestRegs(a0, a1, a2, a3, idx, meta, r, I, flags, planeShadow)) continue;
        const float distance = length(I - r.o);
        if (!(distance > eps)) continue;
        if ((dot(I - r.o, r.d) < 0.f) != behind) continue; // hits behind the origin (cylinders/cones only) come from the point query
        // the reference only tests a primitive whose leaf box passes its slab test (:690); checked for hits only
        // (UW_CLOSEST: t_min(leaf) <= entry distance <= hit distance < closest-so-far whenever the hit would be accepted, so only the
        //  geometric part of the leaf test can reject it)
        float leafT;
        if (!slabT(a4, a5, r, (mode == UW_CLOSEST) ? 3.0e38f : minDistance0, leafT)) continue;
        if (mode == UW_SHADOW)
        {
            if (distance < lenOL) { out.shadow = fmaxf(0.f, fminf(shadowLimit, shadowLimit)); break; }
        }
        else if (mode == UW_CLOSEST)
        {
            if (distance < best || (distance == best && out.hit.prim >= 0 && idx < out.hit.prim))
            {
                best = distance;
                out.hit.prim = idx; out.hit.p = I; out.hit.flags = flags;
                cullT = fminf(minDistance0, best * invLen);
            }
        }
        else if (distance < minDistance0 && distance <= window)
        {
            if (distance < best)
            {
                best = distance;
                window = fminf(minDistance0, GATHER_WINDOW * best);
                cullT = fminf(minDistance0, window * invLen);
                out.hit.prim = idx; out.hit.p = I; out.hit.flags = flags;
                bestLeaf = __float_as_int(a4.w); bestLeafT = leafT; bestTie = false;
            }
            else if (distance == best && idx != out.hit.prim) bestTie = true;
            if (n == GATHER_CAP)
            {
                int m2 = 0; // full: drop what fell out of the window meanwhile
                for (int j = 0; j < n; ++j)
                    if (candD[j] <= window)
                    {
                        candIdx[m2] = candIdx[j]; candD[m2] = candD[j]; candLeafT[m2] = candLeafT[j]; candLeaf[m2] = candLeaf[j];
                        ++m2;
                    }
                n = m2;
            }
            if (n == GATHER_CAP) { overflow = true; break; }
            // appended in visiting order; the replay below picks them in array order
            candIdx[n] = idx; candD[n] = distance; candLeafT[n] = leafT; candLeaf[n] = __float_as_int(a4.w);
            ++n;
        }
    }
    DBG_ADD(2, 1); DBG_ADD(3, overflow ? 1 : 0); DBG_ADD(4, n); DBG_MAX(6, dbgVisits);
    if (overflow)
    {
        out.hit.prim = -2; out.shadow = -1.f; // caller runs the ordered walk
        return out;
    }
    if (mode != UW_GATHER) return out;
    if (n == 0) { out.hit.prim = -1; return out; }
    // The reference's accept/reject sequence almost always ends with the closest candidate X (kept in out.hit with its hit point):
    // every candidate accepted before X's leaf is reached lies farther than X, so the leaf passes its t_min < closest-so-far test
    // unless one of them lies within [d_X, t_min(leaf of X)] — a candidate of another leaf, earlier in the array, at most t_min(leaf
    // of X) away.  One pass over the list looks for such a candidate (any, accepted or not: conservative); only then — or when two
    // candidates share the closest distance — the full replay below decides.  (Round 2: the replay, quadratic in the list length and
    // run by one to three lanes of a warp, was a quarter of a bounce pass's time; profiles/r02_history.md.)
    float m = minDistance0;
    int winner = -1;
    {
        bool slow = bestTie;
        for (int j = 0; j < n; ++j)
            slow |= candIdx[j] < out.hit.prim && candLeaf[j] != bestLeaf && candD[j] <= bestLeafT && candD[j] <= window;
        // Candidates beyond the window are left out.  They can only matter through a chain of candidates each within 1/L of the
        // next (see above): if such a chain climbs from the closest candidate into the outermost shell of the window, (L * window,
        // window], a candidate outside could hang on to it, and then the ordered walk decides rather than a heuristic.  At
        // L = 0.95 — every bounce ray the shader makes — that takes seven links, i.e. eight candidates: the gate keeps the check
        // out of the common case (it cost 2 % of the frame run on every ray).
        if (n >= 8 || lenOL < 0.949f)
        {
            float dmax = 0.f;
            for (int j = 0; j < n; ++j)
                if (candD[j] <= window) dmax = fmaxf(dmax, candD[j]);
            if (dmax > lenOL * window)
            {
                float reach = best;
                for (bool grew = true; grew;)
                {
                    grew = false;
                    float next = reach;
                    const float limit = reach / lenOL * 1.000001f;
                    for (int j = 0; j < n; ++j)
                        if (candD[j] > reach && candD[j] <= limit && candD[j] <= window) next = fmaxf(next, candD[j]);
                    if (next > reach) { reach = next; grew = true; }
                }
                if (reach > lenOL * window)
                {
                    out.hit.prim = -2; out.shadow = -1.f; // caller runs the ordered walk
                    return out;
                }
            }
        }
        DBG_ADD(3, slow ? 1 : 0);
        if (!slow) winner = out.hit.prim;
        else
        {
            // replay in array order (selection by ascending index: the list is short; a primitive listed twice — by the point query,
            // or as two pieces of a long cylinder — is taken once).  Primitives of one leaf are contiguous and share the leaf's fate,
            // decided when the leaf is reached (before any of its primitives): t_min(leaf) < closest-so-far.
            bool leafPass = false;
            int prevLeaf = -1;
            int last = -1;
            for (int pass = 0; pass < n; ++pass)
            {
                int bj = -1, bi = 0x7fffffff;
                for (int j = 0; j < n; ++j)
                {
                    const int ci = candIdx[j];
                    if (ci > last && ci < bi && candD[j] <= window) { bi = ci; bj = j; }
                }
                if (bj < 0) break;
                last = bi;
                if (candLeaf[bj] != prevLeaf)
                {
                    leafPass = candLeafT[bj] < m;
                    prevLeaf = candLeaf[bj];
                }
                if (leafPass && candD[bj] < m) { m = candD[bj]; winner = bi; }
            }
        }
    }
    // the winner's hit point comes from ONE copy of the primitive test whichever way it was found and whatever the tree looked
    // like (two differently scheduled copies may round a grazing hit apart)
    out.hit.prim = -1; out.hit.p = f3(0.f, 0.f, 0.f); out.hit.flags = 0;
    if (winner >= 0)
    {
        const int meta = __ldg(cS.meta + winner);
        float3 I;
        int flags;
        float planeShadow;
        primitiveTest(winner, meta, r, I, flags, planeShadow); // deterministic: same hit point as when it was gathered
        out.hit.prim = winner; out.hit.p = I; out.hit.flags = flags;
    }
    return out;
}
#endif

#ifndef UW_GROUP
#define UW_GROUP 0 // 0: one ray per lane (unorderedWalk above, default); 1: walks shared out over groups of lanes (tracegroup.cuh; exact, measured slower: profiles/r02_history.md)
#endif
#if UW_GROUP
#include "tracegroup.cuh"
#endif

SB_DEV Hit closestHitOrderIndependent(const float3 origin, const float3 target, const int iteration, const int currentMaterialId)
{
    const float3 d = target - origin;
    const int mode = (dot(d, d) >= 1.0002f) ? UW_CLOSEST : UW_GATHER;
    const WalkOut o = unorderedWalk(mode, origin, d, iteration, currentMaterialId, 0, 0);
    if (o.hit.prim == -2) return closestHitWide(origin, target, iteration, currentMaterialId);
    return o.hit;
}
