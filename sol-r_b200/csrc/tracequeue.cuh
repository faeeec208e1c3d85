// tracequeue.cuh — closest-hit walks of a whole queue of rays in one persistent kernel with per-lane refill.
//
// The staged renderer (engine.cu) parks every path between passes, so the closest-hit walks of a pass are just a list of
// rays.  Walk lengths differ a lot between rays (a ray that grazes the molecule lattice visits ten times the nodes of one
// that hits the first atom), and inside unorderedWalk() a warp waits for its slowest lane.  Here a lane that finishes its
// ray writes the hit and takes the next ray from the queue while the other lanes keep walking, so the node loop runs
// with (nearly) all lanes busy — the "persistent threads / dynamic fetch" scheme.  The walk itself is unorderedWalk()'s
// (trace.cuh): same node loop, same primitive tests, same acceptance rules for UW_CLOSEST and UW_GATHER, same point query
// for hits behind the origin, same fallback to the ordered walk when a stack or the gather list overflows.
#pragma once
#include "shade.cuh"

#define HIT_WORDS 5 // primitive, hit point xyz, flags

#ifndef TRACE_MIN_CTAS
#define TRACE_MIN_CTAS 8
#endif
#ifndef TRACE_REFILL_IDLE
#define TRACE_REFILL_IDLE 1 // refill as soon as this many lanes are idle
#endif
#ifndef TRACE_LEAF_VOTE
#define TRACE_LEAF_VOTE 12 // lanes holding a leaf that end the node rounds
#endif
#ifndef TRACE_RETIRE_VOTE
#define TRACE_RETIRE_VOTE 8 // lanes out of work that end the node rounds (while the queue still has rays)
#endif

__global__ void __launch_bounds__(128, TRACE_MIN_CTAS) k_trace_closest(const int pass)
{
#ifdef TRACE_VIA_CALL
    {
        // debug variant: the same queue, walked by calling unorderedWalk() per ray
        const unsigned int count = cP.queueCounters[2 * passQueue(pass)];
        const size_t stride = cP.pathStride;
        for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        {
            const size_t slot = (size_t)cP.pathQueues[(size_t)passQueue(pass) * stride + i];
            const float* pw = cP.pathWords + slot;
            const float3 o = f3(pw[0], pw[stride], pw[2 * stride]);
            const float3 t = f3(pw[3 * stride], pw[4 * stride], pw[5 * stride]);
            const Hit out = closestHitOrderIndependent(o, t, pass, __float_as_int(pw[7 * stride]));
            float* hw = cP.hitWords + slot;
            hw[0] = __int_as_float(out.prim); hw[stride] = out.p.x; hw[2 * stride] = out.p.y; hw[3 * stride] = out.p.z;
            hw[4 * stride] = __int_as_float(out.flags);
        }
        return;
    }
#endif
    const int lane = threadIdx.x & 31;
    const int qid = passQueue(pass);
    const unsigned int count = cP.queueCounters[2 * qid];
    const int* __restrict__ queue = cP.pathQueues + (size_t)qid * cP.pathStride;
    const size_t stride = cP.pathStride;
    const float minDistance0 = (pass < 2) ? cSI.viewDistance : cSI.viewDistance / (pass + 1);
    const float eps = cSI.geometryEpsilon;
    const float4* __restrict__ leafRecs = cS.leafRecs;
    const int* __restrict__ metas = cS.meta;
    const float4* __restrict__ nodes = cS.uwnodes;
    const int nbMain = cS.nbUWide;
    const bool pointQuery = cS.nbUX > 0;

    // ---- per-lane walk state
    bool active = false;
    size_t slot = 0;
    Ray r;
    NodeRay q;
    int mode = UW_CLOSEST, currentMaterialId = 0;
    float invLen = 0.f, best = 0.f, window = 0.f, cullT = 0.f;
    __shared__ int2 s_stack[SM_STACK * WALK_THREADS];
    int stackRef[UN_STACK - SM_STACK];
    float stackT[UN_STACK - SM_STACK];
    WalkStack st;
    st.sm = s_stack + threadIdx.x; st.lref = stackRef; st.lt = stackT;
    int sp = 0;
    int candIdx[GATHER_CAP], candLeaf[GATHER_CAP];
    float candD[GATHER_CAP], candLeafT[GATHER_CAP];
    int n = 0;
    bool overflow = false;
    Hit out;
    out.prim = -1; out.p = f3(0.f, 0.f, 0.f); out.flags = 0;
    r.o = r.d = r.nd = r.inv = f3(0.f, 0.f, 0.f);
    q.ix = q.iy = q.iz = q.nox = q.noy = q.noz = 0.f;
    bool exhausted = false; // warp-uniform: the queue has been handed out

    while (true)
    {
        // ---- refill the idle lanes
        const unsigned int idle = __ballot_sync(FULL_MASK, !active);
        if (idle != 0 && !exhausted && __popc(idle) >= TRACE_REFILL_IDLE)
        {
            const int nIdle = __popc(idle);
            unsigned int base = 0;
            const int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(cP.queueCounters + 2 * qid + 1, (unsigned int)nIdle);
            base = __shfl_sync(FULL_MASK, base, leader);
            if (!active)
            {
                const unsigned int my = base + __popc(idle & ((1u << lane) - 1u));
                if (my < count)
                {
                    slot = (size_t)queue[my];
                    const float* pw = cP.pathWords + slot;
                    const float3 o = f3(pw[0], pw[stride], pw[2 * stride]);
                    const float3 t = f3(pw[3 * stride], pw[4 * stride], pw[5 * stride]);
                    currentMaterialId = __float_as_int(pw[7 * stride]);
                    const float3 d = t - o;
                    makeRay(r, o, d);
                    nodeRay(q, r);
                    const float len2 = dot(d, d);
                    mode = (len2 >= 1.0002f) ? UW_CLOSEST : UW_GATHER;
                    invLen = rsqrtf(len2) * 1.0001f;
                    best = minDistance0; window = minDistance0;
                    cullT = (mode == UW_GATHER) ? minDistance0 : fminf(minDistance0, minDistance0 * invLen);
                    st.push(0, 0, 0.f); sp = 1;
                    if (pointQuery) { st.push(1, nbMain, -3.0e38f); sp = 2; }
                    n = 0; overflow = false;
                    out.prim = -1; out.p = f3(0.f, 0.f, 0.f); out.flags = 0;
                    active = true;
                }
            }
            if (base + nIdle >= count) exhausted = true;
        }
        if (!__any_sync(FULL_MASK, active)) break;

        // ---- node rounds: every lane without a leaf in hand pops one entry per round.  The rounds stop when enough lanes
        //      hold a leaf (they wait meanwhile), enough lanes have run out of work (they wait for the refill), or nobody
        //      is searching — so neither phase runs for the sake of a few stragglers.
        int cur = WIDE_NONE;
        while (true)
        {
            const bool searching = active && cur == WIDE_NONE && sp > 0 && !overflow;
            const unsigned int ms = __ballot_sync(FULL_MASK, searching);
            const unsigned int ml = __ballot_sync(FULL_MASK, active && cur != WIDE_NONE);
            const unsigned int md = __ballot_sync(FULL_MASK, active && cur == WIDE_NONE && (sp == 0 || overflow));
            if (ms == 0 || __popc(ml) >= TRACE_LEAF_VOTE || (__popc(md) >= TRACE_RETIRE_VOTE && !exhausted)) break;
            if (searching)
            {
                --sp;
                int ref;
                float tEntry;
                st.pop(sp, ref, tEntry);
                if (!(tEntry > cullT)) // else: the bound shrank since this entry was pushed
                {
                    if (ref < 0) cur = ref;
                    else if (!unorderedStep(nodes, ref, q, (ref >= nbMain) ? 0.f : cullT, st, sp)) { overflow = true; sp = 0; }
                }
            }
        }

        __syncwarp(); // lanes leave the node loop at different times; the primitive test below should find them together
        // ---- leaf: one primitive
        bool finished = active && ((cur == WIDE_NONE && sp == 0) || overflow);
        if (active && cur != WIDE_NONE && !overflow)
        {
            const bool behind = ((~cur) & 0x40000000) != 0; // from the point-query tree
            const int idx = (~cur) & 0x3FFFFFFF;
            const int meta = __ldg(metas + idx);
            const int fast = PM_FAST(meta);
            const bool test = fast == 0 || (fast == 1 && currentMaterialId != PM_MATERIAL(meta));
            float3 I;
            int flags;
            float planeShadow;
            if (test && primitiveTest(idx, meta, r, I, flags, planeShadow))
            {
                const float distance = length(I - r.o);
                // hits behind the origin (cylinders/cones only) come from the point query
                if (distance > eps && ((dot(I - r.o, r.d) < 0.f) == behind))
                {
                    // the reference only tests a primitive whose leaf box passes its slab test (:690); checked for hits only
                    const int leaf = __ldg(cS.primLeaf + idx);
                    const float4 lo = __ldg(leafRecs + 2 * leaf);
                    const float4 hi = __ldg(leafRecs + 2 * leaf + 1);
                    float leafT;
                    if (slabT(lo, hi, r, (mode == UW_CLOSEST) ? 3.0e38f : minDistance0, leafT))
                    {
                        if (mode == UW_CLOSEST)
                        {
                            if (distance < best || (distance == best && out.prim >= 0 && idx < out.prim))
                            {
                                best = distance;
                                out.prim = idx; out.p = I; out.flags = flags;
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
                            if (n == GATHER_CAP) { overflow = true; finished = true; }
                            else
                            {
                                candIdx[n] = idx; candD[n] = distance; candLeafT[n] = leafT; candLeaf[n] = leaf;
                                ++n;
                            }
                        }
                    }
                }
            }
        }

        __syncwarp();
        // ---- finished rays: replay (UW_GATHER), fallback, result.  Retired in groups: the code below is per-ray overhead, and
        //      run for one lane at a time it costs as many issue slots as the ray's whole walk; a finished lane waits until
        //      TRACE_RETIRE_VOTE lanes are finished, or nothing else in the warp can make progress.
        {
            const unsigned int mf = __ballot_sync(FULL_MASK, finished);
            const unsigned int mp = __ballot_sync(FULL_MASK, active && !finished);
            if (!(__popc(mf) >= TRACE_RETIRE_VOTE || mp == 0 || exhausted)) finished = false;
        }
        if (finished)
        {
            if (overflow)
            {
                const float* pw = cP.pathWords + slot;
                const float3 t = f3(pw[3 * stride], pw[4 * stride], pw[5 * stride]);
                out = closestHitWide(r.o, t, pass, currentMaterialId);
            }
            else if (mode == UW_GATHER)
            {
                float m = minDistance0;
                bool leafPass = false;
                int prevLeaf = -1, winner = -1, last = -1;
                for (int ps = 0; ps < n; ++ps)
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
                out.prim = -1;
                if (winner >= 0)
                {
                    float3 I;
                    int flags;
                    float planeShadow;
                    primitiveTest(winner, __ldg(metas + winner), r, I, flags, planeShadow); // deterministic: same hit point as when it was gathered
                    out.prim = winner; out.p = I; out.flags = flags;
                }
            }
            float* hw = cP.hitWords + slot;
            hw[0] = __int_as_float(out.prim); hw[stride] = out.p.x; hw[2 * stride] = out.p.y; hw[3 * stride] = out.p.z;
            hw[4 * stride] = __int_as_float(out.flags);
            active = false;
        }
        __syncwarp();
    }
}
