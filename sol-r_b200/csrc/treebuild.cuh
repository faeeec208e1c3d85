// treebuild.cuh — the trees of the order-independent walks built on the GPU (b200_set_option(10, 1)).
//
// What it replaces: engine.cu buildWalkTrees(), a binned-SAH build over single primitives on host threads (0.4 s for the 216 k
// primitives of config 2, 1.2 s for 1 M on the GPU box's cores), run inside b200_h2d_scene — i.e. after every compactBoxes() of
// an animated scene (reference: MoleculeScene.cpp:75-81 rotates the primitives and rebuilds the boxes every frame; the walk the
// trees serve replaces GeometryIntersections.cuh:667-772, :798-907).  Here the same two trees — the main tree over one padded box
// per primitive and the point-query tree over the grown cylinder / cone boxes (buildWalkTrees has the why) — come from a linear
// BVH build: 63-bit Morton codes of the box centres, one radix sort (cub), the binary radix tree of Karras (2012) with one thread
// per inner node, boxes bottom-up with one atomic counter per node, and a level-by-level collapse into the 4-wide 128-byte
// records the walks read (the child with the largest surface is opened first, as in the host's collapse).  Level-by-level
// allocation numbers the nodes breadth-first.
// Exactness does not depend on the tree: the walks' results are order-independent and the leaf boxes are the same padded boxes
// the host builds (same float arithmetic), inner boxes are exact unions.  A linear BVH is a worse tree than the SAH one (more node
// visits per ray; measured in profiles/r02_history.md), so the host build stays the default for scenes that are uploaded once.
#pragma once
#include <cub/cub.cuh>

namespace treebuild
{
struct Scratch
{
    char* p = nullptr;
    size_t cap = 0;
};
static Scratch g_scratch;

static __device__ __forceinline__ unsigned int orderedBits(const float f)
{
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static __host__ __device__ __forceinline__ float fromOrderedBits(const unsigned int u)
{
    const unsigned int v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}

// one padded box per primitive: the arithmetic of buildWalkTrees()
static __device__ __forceinline__ void primBox(const b200_Primitive& p, float lo3[3], float hi3[3])
{
    const float P0[3] = {p.p0.x, p.p0.y, p.p0.z}, P1[3] = {p.p1.x, p.p1.y, p.p1.z}, P2[3] = {p.p2.x, p.p2.y, p.p2.z};
    const float S[3] = {p.size.x, p.size.y, p.size.z};
#pragma unroll
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
        const float pad = __fadd_rn(UW_PAD, __fmul_rn(2e-5f, fmaxf(fabsf(lo), fabsf(hi))));
        lo3[a] = __fsub_rn(lo, pad); hi3[a] = __fadd_rn(hi, pad);
    }
}

// the grown boxes of a cylinder / cone (buildWalkTrees(): hits behind the origin): the parameter range of the axis inside the
// grown leaf box and the number of pieces it is covered with; pieces == 0: none
struct ExtRange { double u0, u1, R; int pieces; };
static __device__ __forceinline__ ExtRange extRange(const b200_Primitive& p, const float4 leafLo, const float4 leafHi)
{
    ExtRange e;
    e.u0 = -1e300; e.u1 = 1e300; e.R = 0.0; e.pieces = 0;
    if (!((p.type == B200_PT_CYLINDER || p.type == B200_PT_CONE) && (p.n1.x != 0.f || p.n1.y != 0.f || p.n1.z != 0.f))) return e;
    float blo[3], bhi[3];
    primBox(p, blo, bhi);
    const float L0[3] = {fminf(leafLo.x, blo[0]), fminf(leafLo.y, blo[1]), fminf(leafLo.z, blo[2])};
    const float L1[3] = {fmaxf(leafHi.x, bhi[0]), fmaxf(leafHi.y, bhi[1]), fmaxf(leafHi.z, bhi[2])};
    const float P0[3] = {p.p0.x, p.p0.y, p.p0.z};
    const double N[3] = {p.n1.x, p.n1.y, p.n1.z};
    e.R = fmax(fabs((double)p.size.x), fabs((double)p.size.y)) * 1.001 + 0.05;
    bool empty = false;
    for (int a = 0; a < 3 && !empty; ++a)
    {
        const double slack = e.R + 1e-5 * fmax(fabs((double)L0[a]), fabs((double)L1[a]));
        const double lo = L0[a] - slack, hi = L1[a] + slack;
        if (fabs(N[a]) < 1e-9) { empty = P0[a] < lo || P0[a] > hi; continue; }
        double ua = (lo - P0[a]) / N[a], ub = (hi - P0[a]) / N[a];
        if (ua > ub) { const double t = ua; ua = ub; ub = t; }
        e.u0 = fmax(e.u0, ua); e.u1 = fmin(e.u1, ub);
    }
    if (!empty && e.u0 <= e.u1 && e.u0 > -1e299 && e.u1 < 1e299)
    {
        int pieces = (int)ceil((e.u1 - e.u0) / (4.0 * e.R));
        e.pieces = pieces < 1 ? 1 : (pieces > 64 ? 64 : pieces);
    }
    return e;
}

// boxes[2 i], boxes[2 i + 1] = (lo, leaf ref) (hi, -) of item i; bounds = ordered bits of the min / max of the box centres
static __global__ void k_tb_prim_boxes(const b200_Primitive* __restrict__ prims, const int n, float4* __restrict__ boxes, unsigned int* bounds)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lo[3], hi[3];
    primBox(prims[i], lo, hi);
    boxes[2 * (size_t)i] = make_float4(lo[0], lo[1], lo[2], __int_as_float(~i));
    boxes[2 * (size_t)i + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        const float c = 0.5f * (lo[a] + hi[a]);
        atomicMin(bounds + a, orderedBits(c));
        atomicMax(bounds + 3 + a, orderedBits(c));
    }
}

static __global__ void k_tb_ext_count(const b200_Primitive* __restrict__ prims, const int n, const int* __restrict__ primLeaf,
                                      const float4* __restrict__ leafBoxes, int* __restrict__ counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = primLeaf[i];
    counts[i] = extRange(prims[i], leafBoxes[2 * (size_t)l], leafBoxes[2 * (size_t)l + 1]).pieces;
}

static __global__ void k_tb_ext_boxes(const b200_Primitive* __restrict__ prims, const int n, const int* __restrict__ primLeaf,
                                      const float4* __restrict__ leafBoxes, const int* __restrict__ offsets, float4* __restrict__ boxes,
                                      unsigned int* bounds)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = primLeaf[i];
    const b200_Primitive& p = prims[i];
    const ExtRange e = extRange(p, leafBoxes[2 * (size_t)l], leafBoxes[2 * (size_t)l + 1]);
    const double P0[3] = {p.p0.x, p.p0.y, p.p0.z}, N[3] = {p.n1.x, p.n1.y, p.n1.z};
    for (int k = 0; k < e.pieces; ++k)
    {
        const double ua = e.u0 + (e.u1 - e.u0) * k / e.pieces, ub = e.u0 + (e.u1 - e.u0) * (k + 1) / e.pieces;
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a)
        {
            const double e0 = P0[a] + ua * N[a], e1 = P0[a] + ub * N[a];
            const double slack = e.R + 1e-5 * fmax(fabs(e0), fabs(e1));
            lo[a] = (float)(fmin(e0, e1) - slack);
            hi[a] = (float)(fmax(e0, e1) + slack);
            const float c = 0.5f * (lo[a] + hi[a]);
            atomicMin(bounds + a, orderedBits(c));
            atomicMax(bounds + 3 + a, orderedBits(c));
        }
        const size_t at = (size_t)offsets[i] + k;
        boxes[2 * at] = make_float4(lo[0], lo[1], lo[2], __int_as_float(~(i | 0x40000000)));
        boxes[2 * at + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
}

static __device__ __forceinline__ unsigned long long spread21(unsigned long long x)
{
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

static __global__ void k_tb_morton(const float4* __restrict__ boxes, const int n, const unsigned int* __restrict__ bounds,
                                   unsigned long long* __restrict__ keys, int* __restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 lo = boxes[2 * (size_t)i], hi = boxes[2 * (size_t)i + 1];
    const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    unsigned long long q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        const float b0 = fromOrderedBits(bounds[a]), b1 = fromOrderedBits(bounds[3 + a]);
        const float ext = fmaxf(b1 - b0, 1e-30f);
        const float t = fminf(fmaxf((c[a] - b0) / ext, 0.f), 1.f);
        q[a] = (unsigned long long)fminf(t * 2097152.f, 2097151.f);
    }
    keys[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    vals[i] = i;
}

static __device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, const int n, const int i, const int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// Karras 2012: inner node i of the binary radix tree over the sorted keys; nodes 0 .. n-2 inner, n-1 .. 2n-2 leaves (sorted order)
static __global__ void k_tb_radix_tree(const unsigned long long* __restrict__ keys, const int n, int2* __restrict__ children, int* __restrict__ parent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do
    {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? d : 0);
    const int left = (min(i, j) == gamma) ? (n - 1 + gamma) : gamma;
    const int right = (max(i, j) == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    children[i] = make_int2(left, right);
    parent[left] = i;
    parent[right] = i;
    if (i == 0) parent[0] = -1;
}

// node boxes bottom-up: the second thread to arrive at an inner node merges its children
static __global__ void k_tb_fit(const float4* __restrict__ itemBoxes, const int* __restrict__ sorted, const int n, const int2* __restrict__ children,
                                const int* __restrict__ parent, float4* __restrict__ nodeBoxes, int* __restrict__ flags)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int item = sorted[k];
    int node = n - 1 + k;
    nodeBoxes[2 * (size_t)node] = itemBoxes[2 * (size_t)item];
    nodeBoxes[2 * (size_t)node + 1] = itemBoxes[2 * (size_t)item + 1];
    __threadfence();
    node = parent[node];
    while (node >= 0)
    {
        if (atomicAdd(flags + node, 1) == 0) return;
        __threadfence();
        const int2 c = children[node];
        const volatile float4* nb = nodeBoxes;
        const float4 l0 = make_float4(nb[2 * (size_t)c.x].x, nb[2 * (size_t)c.x].y, nb[2 * (size_t)c.x].z, 0.f);
        const float4 l1 = make_float4(nb[2 * (size_t)c.x + 1].x, nb[2 * (size_t)c.x + 1].y, nb[2 * (size_t)c.x + 1].z, 0.f);
        const float4 r0 = make_float4(nb[2 * (size_t)c.y].x, nb[2 * (size_t)c.y].y, nb[2 * (size_t)c.y].z, 0.f);
        const float4 r1 = make_float4(nb[2 * (size_t)c.y + 1].x, nb[2 * (size_t)c.y + 1].y, nb[2 * (size_t)c.y + 1].z, 0.f);
        nodeBoxes[2 * (size_t)node] = make_float4(fminf(l0.x, r0.x), fminf(l0.y, r0.y), fminf(l0.z, r0.z), 0.f);
        nodeBoxes[2 * (size_t)node + 1] = make_float4(fmaxf(l1.x, r1.x), fmaxf(l1.y, r1.y), fmaxf(l1.z, r1.z), 0.f);
        __threadfence();
        node = parent[node];
    }
}

// One level of the collapse: wide node (levelBase + t) is binary inner node frontier[t]; its children are the two children of the
// binary node with the largest inner one opened until there are four; inner children become the next level's wide nodes.
static __global__ void k_tb_collapse(const int* __restrict__ frontier, const int count, const int levelBase, const int n, const int2* __restrict__ children,
                                     const float4* __restrict__ nodeBoxes, int* __restrict__ nextFrontier, int* __restrict__ nextCount,
                                     float4* __restrict__ wide, const int refOffset)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int b = frontier[t];
    int kids[4];
    int nk = 2;
    {
        const int2 c = children[b];
        kids[0] = c.x; kids[1] = c.y;
    }
    while (nk < 4)
    {
        int best = -1;
        float bestArea = -1.f;
        for (int k = 0; k < nk; ++k)
            if (kids[k] < n - 1)
            {
                const float4 lo = nodeBoxes[2 * (size_t)kids[k]], hi = nodeBoxes[2 * (size_t)kids[k] + 1];
                const float x = hi.x - lo.x, y = hi.y - lo.y, z = hi.z - lo.z;
                const float area = x * y + y * z + z * x;
                if (area > bestArea) { bestArea = area; best = k; }
            }
        if (best < 0) break;
        const int2 c = children[kids[best]];
        for (int k = nk; k > best + 1; --k) kids[k] = kids[k - 1];
        kids[best] = c.x; kids[best + 1] = c.y;
        ++nk;
    }
    float rows[6][4];
    int refs[4];
    const int nextBase = levelBase + count;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        if (k < nk)
        {
            const float4 lo = nodeBoxes[2 * (size_t)kids[k]], hi = nodeBoxes[2 * (size_t)kids[k] + 1];
            rows[0][k] = lo.x; rows[1][k] = lo.y; rows[2][k] = lo.z; rows[3][k] = hi.x; rows[4][k] = hi.y; rows[5][k] = hi.z;
            if (kids[k] >= n - 1) refs[k] = __float_as_int(lo.w); // leaf: the item's ref rides in its box
            else
            {
                const int pos = atomicAdd(nextCount, 1);
                nextFrontier[pos] = kids[k];
                refs[k] = refOffset + nextBase + pos;
            }
        }
        else
        {
            rows[0][k] = rows[1][k] = rows[2][k] = 3.0e38f; rows[3][k] = rows[4][k] = rows[5][k] = -3.0e38f;
            refs[k] = (int)0x80000000;
        }
    }
    float4* rec = wide + 8 * (size_t)(levelBase + t);
#pragma unroll
    for (int r = 0; r < 6; ++r) rec[r] = make_float4(rows[r][0], rows[r][1], rows[r][2], rows[r][3]);
    rec[6] = make_float4(__int_as_float(refs[0]), __int_as_float(refs[1]), __int_as_float(refs[2]), __int_as_float(refs[3]));
    rec[7] = make_float4(__int_as_float(nk), 0.f, 0.f, 0.f);
}

static __global__ void k_tb_single(const float4* __restrict__ itemBoxes, float4* __restrict__ wide)
{
    // a tree of one item: one wide node with one child
    const float4 lo = itemBoxes[0], hi = itemBoxes[1];
    float4* rec = wide;
    const float E = 3.0e38f;
    rec[0] = make_float4(lo.x, E, E, E); rec[1] = make_float4(lo.y, E, E, E); rec[2] = make_float4(lo.z, E, E, E);
    rec[3] = make_float4(hi.x, -E, -E, -E); rec[4] = make_float4(hi.y, -E, -E, -E); rec[5] = make_float4(hi.z, -E, -E, -E);
    rec[6] = make_float4(lo.w, __int_as_float((int)0x80000000), __int_as_float((int)0x80000000), __int_as_float((int)0x80000000));
    rec[7] = make_float4(__int_as_float(1), 0.f, 0.f, 0.f);
}

// One relaxation pass of a re-fit of the main tree in place: every child slot of every node takes the (padded) box of its primitive
// or the union of its child node's slots (empty slots hold +-3e38 and drop out of the min / max).  `levels` passes settle the tree:
// after pass p every node at most p levels above the leaves is final, and a final node is only ever re-written with the same values.
static __global__ void k_tb_refit(float4* __restrict__ wide, const int nbNodes, const b200_Primitive* __restrict__ prims)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbNodes) return;
    float4* rec = wide + 8 * (size_t)i;
    const float4 rf = rec[6];
    const int refs[4] = {__float_as_int(rf.x), __float_as_int(rf.y), __float_as_int(rf.z), __float_as_int(rf.w)};
    float rows[6][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
    {
        float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        if (refs[c] != (int)0x80000000)
        {
            if (refs[c] < 0) primBox(prims[(~refs[c]) & 0x3FFFFFFF], lo, hi);
            else
            {
                const volatile float4* ch = wide + 8 * (size_t)refs[c];
#pragma unroll
                for (int a = 0; a < 3; ++a)
                {
                    const float4 l = make_float4(ch[a].x, ch[a].y, ch[a].z, ch[a].w), h = make_float4(ch[3 + a].x, ch[3 + a].y, ch[3 + a].z, ch[3 + a].w);
                    lo[a] = fminf(fminf(l.x, l.y), fminf(l.z, l.w));
                    hi[a] = fmaxf(fmaxf(h.x, h.y), fmaxf(h.z, h.w));
                }
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) { rows[a][c] = lo[a]; rows[3 + a][c] = hi[a]; }
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) rec[r] = make_float4(rows[r][0], rows[r][1], rows[r][2], rows[r][3]);
}
static int refitMain(float4* wide, const int nbMain, const int levels, const b200_Primitive* dPrims, cudaStream_t stream)
{
    if (nbMain <= 0) return 0;
    for (int p = 0; p < levels; ++p) k_tb_refit<<<(nbMain + 127) / 128, 128, 0, stream>>>(wide, nbMain, dPrims);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

struct Arena
{
    char* base; size_t at, cap;
    template <typename T> T* take(const size_t count)
    {
        at = (at + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + at);
        at += count * sizeof(T);
        return p;
    }
};

static size_t arenaBytes(const size_t n)
{
    size_t sortTemp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sortTemp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n);
    size_t scanTemp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scanTemp, (const int*)nullptr, (int*)nullptr, (int)n);
    const size_t temp = sortTemp > scanTemp ? sortTemp : scanTemp;
    // keys x2, vals x2, children, parent, flags, node boxes, frontiers x2, counters + alignment slack
    return 2 * n * 8 + 2 * n * 4 + n * 8 + 2 * n * 4 + n * 4 + 2 * (2 * n) * 16 + 2 * n * 4 + temp + 64 * 256 + 4096;
}

// LBVH over `n` items whose boxes (with their leaf refs in lo.w) are in itemBoxes and whose centre bounds are in `bounds`; wide
// records are written from node `nodeBase` of `wide` on, inner refs are offset by it.  Returns the number of wide nodes, < 0 on error.
static int buildOne(const float4* itemBoxes, const int n, const unsigned int* bounds, float4* wide, const int nodeBase, Arena& A, cudaStream_t stream,
                    int* levelsOut = nullptr)
{
    if (levelsOut) *levelsOut = n > 0 ? 1 : 0;
    if (n <= 0) return 0;
    if (n == 1)
    {
        k_tb_single<<<1, 1, 0, stream>>>(itemBoxes, wide + 8 * (size_t)nodeBase);
        return 1;
    }
    const size_t at0 = A.at;
    unsigned long long* keys = A.take<unsigned long long>(n);
    unsigned long long* keys2 = A.take<unsigned long long>(n);
    int* vals = A.take<int>(n);
    int* vals2 = A.take<int>(n);
    int2* children = A.take<int2>(n);
    int* parent = A.take<int>(2 * (size_t)n);
    int* flags = A.take<int>(n);
    float4* nodeBoxes = A.take<float4>(2 * (2 * (size_t)n));
    int* frontA = A.take<int>(n);
    int* frontB = A.take<int>(n);
    int* counters = A.take<int>(64);
    size_t tempBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, keys, keys2, vals, vals2, n);
    char* temp = A.take<char>(tempBytes);
    if (A.at > A.cap) return -1;
    const int T = 256, B = (n + T - 1) / T;
    k_tb_morton<<<B, T, 0, stream>>>(itemBoxes, n, bounds, keys, vals);
    if (cub::DeviceRadixSort::SortPairs(temp, tempBytes, keys, keys2, vals, vals2, n, 0, 63, stream) != cudaSuccess) return -2;
    k_tb_radix_tree<<<B, T, 0, stream>>>(keys2, n, children, parent);
    if (cudaMemsetAsync(flags, 0, (size_t)n * sizeof(int), stream) != cudaSuccess) return -2;
    k_tb_fit<<<B, T, 0, stream>>>(itemBoxes, vals2, n, children, parent, nodeBoxes, flags);
    // collapse, level by level
    int count = 1, levelBase = 0;
    const int root = 0;
    if (cudaMemcpyAsync(frontA, &root, sizeof(int), cudaMemcpyHostToDevice, stream) != cudaSuccess) return -2;
    int* cur = frontA; int* nxt = frontB;
    int level = 0;
    while (count > 0)
    {
        int* ctr = counters + (level & 63);
        if (cudaMemsetAsync(ctr, 0, sizeof(int), stream) != cudaSuccess) return -2;
        k_tb_collapse<<<(count + 127) / 128, 128, 0, stream>>>(cur, count, nodeBase + levelBase, n, children, nodeBoxes, nxt, ctr, wide, 0);
        int next = 0;
        if (cudaMemcpyAsync(&next, ctr, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return -2;
        if (cudaStreamSynchronize(stream) != cudaSuccess) return -2;
        levelBase += count;
        count = next;
        int* t = cur; cur = nxt; nxt = t;
        ++level;
    }
    A.at = at0; // the scratch of this tree is free again (everything above ran to completion)
    if (levelsOut) *levelsOut = level;
    return levelBase;
}

// Number of grown cylinder / cone boxes the point-query tree will hold (one pass over the primitives); offsets are left in the
// scratch for buildWalkTreesGpu.  < 0 on error.
static bool ensureScratch(Scratch& sc, const size_t bytes, cudaStream_t stream)
{
    if (bytes <= sc.cap) return true;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return false;
    if (sc.p) cudaFree(sc.p);
    sc.p = nullptr; sc.cap = 0;
    if (cudaMalloc(&sc.p, bytes + bytes / 8) != cudaSuccess) { cudaGetLastError(); return false; }
    sc.cap = bytes + bytes / 8;
    return true;
}
static Scratch g_counts; // per-primitive piece counts and offsets (live from extCount to buildWalkTreesGpu)
static Scratch g_ext;    // item boxes and arena of the point-query tree

static int extCount(const b200_Primitive* dPrims, const int nbPrims, const int* dPrimLeaf, const float4* dLeafBoxes, cudaStream_t stream)
{
    if (nbPrims <= 0) return 0;
    const size_t n = (size_t)nbPrims;
    size_t scanTemp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scanTemp, (const int*)nullptr, (int*)nullptr, nbPrims + 1);
    if (!ensureScratch(g_counts, 2 * (n + 1) * 4 + scanTemp + 1024, stream)) return -4;
    Arena A;
    A.base = g_counts.p; A.at = 0; A.cap = g_counts.cap;
    int* counts = A.take<int>(n + 1);
    int* offsets = A.take<int>(n + 1);
    char* temp = A.take<char>(scanTemp);
    const int T = 256, B = (nbPrims + T - 1) / T;
    k_tb_ext_count<<<B, T, 0, stream>>>(dPrims, nbPrims, dPrimLeaf, dLeafBoxes, counts);
    if (cudaMemsetAsync(counts + nbPrims, 0, sizeof(int), stream) != cudaSuccess) return -2;
    if (cub::DeviceScan::ExclusiveSum(temp, scanTemp, counts, offsets, nbPrims + 1, stream) != cudaSuccess) return -2;
    int nbExtBoxes = 0;
    if (cudaMemcpyAsync(&nbExtBoxes, offsets + nbPrims, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return -2;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return -2;
    return nbExtBoxes;
}

// Both trees into `wide` (room for nbPrims + nbExtBoxes 128-byte records; nbExtBoxes from extCount() just before).  primLeaf /
// leafBoxes: the reference leaf of each primitive and the reference leaves' boxes (2 float4 each).  Returns 0, or < 0 on error.
// keepMain: the main tree in `wide` (nbMain nodes) stays as it is — re-fitted by its owner — and only the point-query tree behind it
// is rebuilt (an animation step: the grown boxes change their number with the orientation, the main tree's shape need not).
static int buildWalkTreesGpu(const b200_Primitive* dPrims, const int nbPrims, const int* dPrimLeaf, const float4* dLeafBoxes, const int nbExtBoxes,
                             float4* wide, int& nbMain, int& nbExt, cudaStream_t stream, int* mainLevels = nullptr, const bool keepMain = false)
{
    if (!keepMain) nbMain = 0;
    nbExt = 0;
    if (nbPrims <= 0) return 0;
    const size_t n = (size_t)nbPrims;
    if (!ensureScratch(g_scratch, 2 * n * 16 + 256 + arenaBytes(n) + 4096, stream)) return -4;
    Arena A;
    A.base = g_scratch.p; A.at = 0; A.cap = g_scratch.cap;
    unsigned int* bounds = A.take<unsigned int>(16);
    float4* itemBoxes = A.take<float4>(2 * n);
    const unsigned int boundsInit[12] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    if (cudaMemcpyAsync(bounds, boundsInit, sizeof(boundsInit), cudaMemcpyHostToDevice, stream) != cudaSuccess) return -2;
    const int T = 256, B = (nbPrims + T - 1) / T;
    if (!keepMain)
    {
        k_tb_prim_boxes<<<B, T, 0, stream>>>(dPrims, nbPrims, itemBoxes, bounds);
        const int m = buildOne(itemBoxes, nbPrims, bounds, wide, 0, A, stream, mainLevels);
        if (m < 0) return m;
        nbMain = m;
    }
    if (nbExtBoxes > 0)
    {
        // the point-query tree, appended: its nodes are numbered from nbMain on
        const size_t e = (size_t)nbExtBoxes;
        if (!ensureScratch(g_ext, 2 * e * 16 + arenaBytes(e) + 4096, stream)) return -4;
        Arena X;
        X.base = g_ext.p; X.at = 0; X.cap = g_ext.cap;
        float4* extBoxes = X.take<float4>(2 * e);
        const int* offsets = reinterpret_cast<const int*>(g_counts.p + ((((n + 1) * 4) + 255) & ~(size_t)255));
        k_tb_ext_boxes<<<B, T, 0, stream>>>(dPrims, nbPrims, dPrimLeaf, dLeafBoxes, offsets, extBoxes, bounds + 6);
        const int x = buildOne(extBoxes, nbExtBoxes, bounds + 6, wide, nbMain, X, stream);
        if (x < 0) return x;
        nbExt = x;
    }
    return 0;
}

static void releaseScratch()
{
    if (g_scratch.p) cudaFree(g_scratch.p);
    if (g_counts.p) cudaFree(g_counts.p);
    if (g_ext.p) cudaFree(g_ext.p);
    g_scratch = Scratch(); g_counts = Scratch(); g_ext = Scratch();
}
} // namespace treebuild
