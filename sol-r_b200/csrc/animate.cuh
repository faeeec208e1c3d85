// animate.cuh — the reference's animation step on the device-resident scene (b200_rotate_primitives / b200_translate_primitives /
// b200_scale_primitives).
//
// What it replaces: GPUKernel::rotatePrimitives / translatePrimitives / scalePrimitives followed by compactBoxes(false)
// (GPUKernel.cpp:1378-1513, :1574-1600, :1041-1083; MoleculeScene.cpp:75-81 does it every frame) and the upload that follows
// (CudaKernel.cpp:197-230): on the host that is a pass over every primitive through std::map look-ups, a re-fit of every box of
// every level, a depth-first re-flatten of the whole hierarchy and a full upload — 0.2 s + 0.26 s for the 216 k primitives of
// config 2.  The step does not change the hierarchy's SHAPE (the reference re-fits its boxes, it does not re-bin), so the
// flattened arrays keep their order, skip links and primitive ranges and only coordinates change.  Here the same arithmetic
// (operation by operation, no contraction: the arrays stay byte-identical to the host container's, tests/test_gpu_parity.py
// test_device_animation_equals_the_host_step) runs on the arrays where they live:
//   k_an_move          every primitive outside the lights box (the reference moves what its level-0 boxes list)
//   k_an_leaf_boxes    every leaf box re-fitted to its primitives (reset to +-1e6, then the per-type extents; GPUKernel.cpp:741-839)
//   k_an_inner_boxes   level by level (an inner box's startIndex is its level), an inner box re-fitted to its children (:841-892)
// and then everything the engine derives from them, without leaving the device: the ordered tree (same shape, leaves from the
// reference leaf boxes, inner nodes bottom-up with one atomic per node), its 4-wide form and leaf records, geometry and primitive
// records, and the walk trees (rebuilt by treebuild.cuh).
#pragma once

namespace animate
{
enum { MOVE_ROTATE = 0, MOVE_TRANSLATE = 1, MOVE_SCALE = 2 };
struct Move
{
    int mode;
    float3 center, cosA, sinA; // rotate
    float3 t;                  // translate
    float s;                   // scale
};

// GPUKernel.cpp:1602-1632, one rounding per operation as the host build does it
static __device__ __forceinline__ void rotateVector(b200_float3& v, const float3 c, const float3 cosA, const float3 sinA)
{
    float3 vec = make_float3(__fsub_rn(v.x, c.x), __fsub_rn(v.y, c.y), __fsub_rn(v.z, c.z));
    float3 r = vec;
    r.y = __fsub_rn(__fmul_rn(vec.y, cosA.x), __fmul_rn(vec.z, sinA.x));
    r.z = __fadd_rn(__fmul_rn(vec.y, sinA.x), __fmul_rn(vec.z, cosA.x));
    vec = r;
    r.z = __fsub_rn(__fmul_rn(vec.z, cosA.y), __fmul_rn(vec.x, sinA.y));
    r.x = __fadd_rn(__fmul_rn(vec.z, sinA.y), __fmul_rn(vec.x, cosA.y));
    vec = r;
    r.x = __fsub_rn(__fmul_rn(vec.x, cosA.z), __fmul_rn(vec.y, sinA.z));
    r.y = __fadd_rn(__fmul_rn(vec.x, sinA.z), __fmul_rn(vec.y, cosA.z));
    v.x = __fadd_rn(r.x, c.x); v.y = __fadd_rn(r.y, c.y); v.z = __fadd_rn(r.z, c.z);
}

static __global__ void k_an_move(b200_Primitive* __restrict__ prims, const int first, const int n, const Move m)
{
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    b200_Primitive p = prims[i];
    if (m.mode == MOVE_SCALE)
    {
        p.p0.x = __fmul_rn(p.p0.x, m.s); p.p0.y = __fmul_rn(p.p0.y, m.s); p.p0.z = __fmul_rn(p.p0.z, m.s);
        p.p1.x = __fmul_rn(p.p1.x, m.s); p.p1.y = __fmul_rn(p.p1.y, m.s); p.p1.z = __fmul_rn(p.p1.z, m.s);
        p.p2.x = __fmul_rn(p.p2.x, m.s); p.p2.y = __fmul_rn(p.p2.y, m.s); p.p2.z = __fmul_rn(p.p2.z, m.s);
        p.size.x = __fmul_rn(p.size.x, m.s); p.size.y = __fmul_rn(p.size.y, m.s); p.size.z = __fmul_rn(p.size.z, m.s);
        prims[i] = p;
        return;
    }
    if (p.type == B200_PT_CAMERA) return;
    if (m.mode == MOVE_TRANSLATE)
    {
        p.p0.x = __fadd_rn(p.p0.x, m.t.x); p.p0.y = __fadd_rn(p.p0.y, m.t.y); p.p0.z = __fadd_rn(p.p0.z, m.t.z);
        p.p1.x = __fadd_rn(p.p1.x, m.t.x); p.p1.y = __fadd_rn(p.p1.y, m.t.y); p.p1.z = __fadd_rn(p.p1.z, m.t.z);
        p.p2.x = __fadd_rn(p.p2.x, m.t.x); p.p2.y = __fadd_rn(p.p2.y, m.t.y); p.p2.z = __fadd_rn(p.p2.z, m.t.z);
        prims[i] = p;
        return;
    }
    // GPUKernel.cpp:1641-1674
    rotateVector(p.p0, m.center, m.cosA, m.sinA);
    if (p.type == B200_PT_CYLINDER || p.type == B200_PT_TRIANGLE)
    {
        rotateVector(p.p1, m.center, m.cosA, m.sinA);
        rotateVector(p.p2, m.center, m.cosA, m.sinA);
        const float3 zero = make_float3(0.f, 0.f, 0.f);
        rotateVector(p.n0, zero, m.cosA, m.sinA);
        rotateVector(p.n1, zero, m.cosA, m.sinA);
        rotateVector(p.n2, zero, m.cosA, m.sinA);
        if (p.type == B200_PT_CYLINDER)
        {
            float ax = __fsub_rn(p.p1.x, p.p0.x), ay = __fsub_rn(p.p1.y, p.p0.y), az = __fsub_rn(p.p1.z, p.p0.z);
            const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az)));
            if (len != 0.f) { ax = __fdiv_rn(ax, len); ay = __fdiv_rn(ay, len); az = __fdiv_rn(az, len); }
            p.n1.x = ax; p.n1.y = ay; p.n1.z = az;
        }
    }
    prims[i] = p;
}

// scalePrimitives scales the lamps too, and the light information the shader reads is their position (GPUKernel.cpp:1199-1212)
static __global__ void k_an_scale_lights(b200_LightInformation* __restrict__ lights, const int n, const float s)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lights[i].location.x = __fmul_rn(lights[i].location.x, s);
    lights[i].location.y = __fmul_rn(lights[i].location.y, s);
    lights[i].location.z = __fmul_rn(lights[i].location.z, s);
}

// the box one primitive contributes to its cell (GPUKernel.cpp:752-826)
// std::min(a, b) = b < a ? b : a, std::max(a, b) = a < b ? b : a (signed zeros fall the same way as on the host)
#define AN_MIN(A, B) ((B) < (A) ? (B) : (A))
#define AN_MAX(A, B) ((A) < (B) ? (B) : (A))
static __device__ __forceinline__ void primitiveExtent(const b200_Primitive& p, float lo[3], float hi[3])
{
    float c0[3], c1[3];
    const float P0[3] = {p.p0.x, p.p0.y, p.p0.z}, P1[3] = {p.p1.x, p.p1.y, p.p1.z}, P2[3] = {p.p2.x, p.p2.y, p.p2.z};
    const float S[3] = {p.size.x, p.size.y, p.size.z};
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        if (p.type == B200_PT_TRIANGLE)
        {
            const float m01 = AN_MIN(P0[a], P1[a]), M01 = AN_MAX(P0[a], P1[a]);
            c0[a] = AN_MIN(m01, P2[a]); c1[a] = AN_MAX(M01, P2[a]);
        }
        else if (p.type == B200_PT_CYLINDER) { c0[a] = AN_MIN(P0[a], P1[a]); c1[a] = AN_MAX(P0[a], P1[a]); }
        else { c0[a] = P0[a]; c1[a] = P0[a]; }
        const float l = (c0[a] <= c1[a]) ? c0[a] : c1[a];
        const float h = (c0[a] > c1[a]) ? c0[a] : c1[a];
        const bool round = p.type == B200_PT_CYLINDER || p.type == B200_PT_SPHERE || p.type == B200_PT_CONE;
        const float s = round ? S[0] : S[a];
        lo[a] = __fsub_rn(l, s); hi[a] = __fadd_rn(h, s);
    }
}

// leaf boxes: every box with primitives except box 0 (the lights box keeps its +-viewDistance bounds)
static __global__ void k_an_leaf_boxes(b200_BoundingBox* __restrict__ boxes, const int nbBoxes, const b200_Primitive* __restrict__ prims, const int nbPrims)
{
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbBoxes) return;
    const int n = boxes[i].nbPrimitives, first = boxes[i].startIndex;
    if (n <= 0) return;
    float lo[3] = {1000000.f, 1000000.f, 1000000.f}, hi[3] = {-1000000.f, -1000000.f, -1000000.f};
    for (int k = 0; k < n; ++k)
    {
        if (first + k < 0 || first + k >= nbPrims) continue;
        float l[3], h[3];
        primitiveExtent(prims[first + k], l, h);
#pragma unroll
        for (int a = 0; a < 3; ++a) { if (l[a] < lo[a]) lo[a] = l[a]; if (h[a] > hi[a]) hi[a] = h[a]; }
    }
    boxes[i].parameters[0].x = lo[0]; boxes[i].parameters[0].y = lo[1]; boxes[i].parameters[0].z = lo[2];
    boxes[i].parameters[1].x = hi[0]; boxes[i].parameters[1].y = hi[1]; boxes[i].parameters[1].z = hi[2];
}

// inner boxes of one level: re-fitted to the children the flattened array lists (children the reference did not emit are empty and
// would contribute nothing)
static __global__ void k_an_inner_boxes(b200_BoundingBox* __restrict__ boxes, const int nbBoxes, const int level, const float vd)
{
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbBoxes) return;
    if (boxes[i].nbPrimitives != 0 || boxes[i].startIndex != level) return;
    const int end = min(nbBoxes, i + boxes[i].indexForNextBox.x);
    float lo[3] = {vd, vd, vd}, hi[3] = {-vd, -vd, -vd};
    for (int c = i + 1; c < end;)
    {
        const b200_BoundingBox& b = boxes[c];
        if (lo[0] > b.parameters[0].x) lo[0] = b.parameters[0].x;
        if (lo[1] > b.parameters[0].y) lo[1] = b.parameters[0].y;
        if (lo[2] > b.parameters[0].z) lo[2] = b.parameters[0].z;
        if (hi[0] < b.parameters[1].x) hi[0] = b.parameters[1].x;
        if (hi[1] < b.parameters[1].y) hi[1] = b.parameters[1].y;
        if (hi[2] < b.parameters[1].z) hi[2] = b.parameters[1].z;
        const int skip = b.indexForNextBox.x;
        c += skip > 0 ? skip : 1;
    }
    boxes[i].parameters[0].x = lo[0]; boxes[i].parameters[0].y = lo[1]; boxes[i].parameters[0].z = lo[2];
    boxes[i].parameters[1].x = hi[0]; boxes[i].parameters[1].y = hi[1]; boxes[i].parameters[1].z = hi[2];
}

// ---- what the engine derives from the reference arrays ----

// leaves of the ordered tree (and the boxes treebuild.cuh grows the cylinder boxes from) from the reference leaf boxes
static __global__ void k_an_ordered_leaves(const b200_BoundingBox* __restrict__ raw, const int* __restrict__ leafRaw, const int* __restrict__ leafNode,
                                           const int nbLeaves, float4* __restrict__ packed, float4* __restrict__ leafBoxes)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nbLeaves) return;
    const b200_BoundingBox& b = raw[leafRaw[k]];
    const int node = leafNode[k];
    packed[2 * (size_t)node].x = b.parameters[0].x; packed[2 * (size_t)node].y = b.parameters[0].y; packed[2 * (size_t)node].z = b.parameters[0].z;
    packed[2 * (size_t)node + 1].x = b.parameters[1].x; packed[2 * (size_t)node + 1].y = b.parameters[1].y; packed[2 * (size_t)node + 1].z = b.parameters[1].z;
    leafBoxes[2 * (size_t)k] = make_float4(b.parameters[0].x, b.parameters[0].y, b.parameters[0].z, 0.f);
    leafBoxes[2 * (size_t)k + 1] = make_float4(b.parameters[1].x, b.parameters[1].y, b.parameters[1].z, 0.f);
}

// inner nodes of the ordered tree (binary, depth-first, skip counts in lo.w), bottom-up: the second arrival merges the children
static __global__ void k_an_ordered_fit(float4* __restrict__ packed, const int* __restrict__ parent, const int* __restrict__ leafNode, const int nbLeaves,
                                        int* __restrict__ flags)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nbLeaves) return;
    __threadfence();
    int node = parent[leafNode[k]];
    while (node >= 0)
    {
        if (atomicAdd(flags + node, 1) == 0) return;
        __threadfence();
        const int left = node + 1;
        const volatile float4* pk = packed;
        const int leftIsLeaf = __float_as_int(pk[2 * (size_t)left + 1].w) > 0;
        const int leftSize = leftIsLeaf ? 1 : __float_as_int(pk[2 * (size_t)left].w);
        const int right = left + leftSize;
        const float lx = fminf(pk[2 * (size_t)left].x, pk[2 * (size_t)right].x), ly = fminf(pk[2 * (size_t)left].y, pk[2 * (size_t)right].y),
                    lz = fminf(pk[2 * (size_t)left].z, pk[2 * (size_t)right].z);
        const float hx = fmaxf(pk[2 * (size_t)left + 1].x, pk[2 * (size_t)right + 1].x), hy = fmaxf(pk[2 * (size_t)left + 1].y, pk[2 * (size_t)right + 1].y),
                    hz = fmaxf(pk[2 * (size_t)left + 1].z, pk[2 * (size_t)right + 1].z);
        packed[2 * (size_t)node].x = lx; packed[2 * (size_t)node].y = ly; packed[2 * (size_t)node].z = lz;
        packed[2 * (size_t)node + 1].x = hx; packed[2 * (size_t)node + 1].y = hy; packed[2 * (size_t)node + 1].z = hz;
        __threadfence();
        node = parent[node];
    }
}

// the 4-wide form: every child slot copies the box of the binary node it was collapsed from; leaf records likewise
static __global__ void k_an_ordered_wide(const float4* __restrict__ packed, const int4* __restrict__ wideKid, const int nbWide, float4* __restrict__ wide)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nbWide) return;
    const int4 kid = wideKid[w];
    const int kids[4] = {kid.x, kid.y, kid.z, kid.w};
    float rows[6][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
    {
        if (kids[c] >= 0)
        {
            const float4 lo = packed[2 * (size_t)kids[c]], hi = packed[2 * (size_t)kids[c] + 1];
            rows[0][c] = lo.x; rows[1][c] = lo.y; rows[2][c] = lo.z; rows[3][c] = hi.x; rows[4][c] = hi.y; rows[5][c] = hi.z;
        }
        else { rows[0][c] = rows[1][c] = rows[2][c] = 3.0e38f; rows[3][c] = rows[4][c] = rows[5][c] = -3.0e38f; }
    }
    float4* rec = wide + 8 * (size_t)w;
#pragma unroll
    for (int r = 0; r < 6; ++r) rec[r] = make_float4(rows[r][0], rows[r][1], rows[r][2], rows[r][3]);
}
static __global__ void k_an_leaf_recs(const float4* __restrict__ packed, const int* __restrict__ leafNode, const int nbLeaves, float4* __restrict__ leafRecs)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nbLeaves) return;
    leafRecs[2 * (size_t)k] = packed[2 * (size_t)leafNode[k]];
    leafRecs[2 * (size_t)k + 1] = packed[2 * (size_t)leafNode[k] + 1];
}

// hot geometry and the unit walk's records (b200_h2d_scene steps 3 / 3b); the packed material word stays what it is
static __global__ void k_an_records(const b200_Primitive* __restrict__ prims, const int n, const int* __restrict__ primLeaf, const float4* __restrict__ leafBoxes,
                                    float4* __restrict__ geo, float4* __restrict__ recs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const b200_Primitive p = prims[i];
    const float4 g0 = make_float4(p.p0.x, p.p0.y, p.p0.z, p.size.x), g1 = make_float4(p.p1.x, p.p1.y, p.p1.z, p.size.y),
                 g2 = make_float4(p.p2.x, p.p2.y, p.p2.z, p.size.z);
    geo[4 * (size_t)i] = g0; geo[4 * (size_t)i + 1] = g1; geo[4 * (size_t)i + 2] = g2; geo[4 * (size_t)i + 3] = make_float4(p.n1.x, p.n1.y, p.n1.z, 0.f);
    if (!recs) return;
    const int l = primLeaf[i];
    float4* rec = recs + (size_t)PRIM_REC_F4 * i;
    const float word = rec[3].w;
    const float4 lo = leafBoxes[2 * (size_t)l], hi = leafBoxes[2 * (size_t)l + 1];
    rec[0] = g0; rec[1] = g1; rec[2] = g2;
    rec[3] = make_float4(p.n1.x, p.n1.y, p.n1.z, word);
    rec[4] = make_float4(lo.x, lo.y, lo.z, __int_as_float(l));
    rec[5] = make_float4(hi.x, hi.y, hi.z, __int_as_float(p.index));
}
} // namespace animate
