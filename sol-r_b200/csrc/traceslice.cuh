// traceslice.cuh — closest-hit walks of a bounce pass in SLICES.  Compiled only with -DWITH_TRACE_SLICE; written from the CPU model
// in tests/analysis_ray_order.py (profiles/r01_history.md), NOT YET RUN ON A GPU — selected by b200_set_option(6, 3) in such a build.
//
// In a bounce pass a warp runs as long as its longest walk: 37 % lane utilisation in the model, a fifth of the rays above 32 node
// visits.  Here a walk stops after `rounds` node visits; the unfinished ones are parked — bound, best candidate or candidate list,
// stack — and queued, and the next slice continues them 32 at a time, so the long walks end up together in full warps (model:
// 0.46-0.62 x the warp-rounds of pass 1).  The walk is unorderedWalk()'s (trace.cuh): same node loop, same primitive tests, same
// acceptance rules, same point query for hits behind the origin, same fallback to the ordered walk on overflow.  Differences in
// bookkeeping only: the candidate list of a gather walk lives in global memory from the start (cSlice.cand[k][slot]) instead of a
// thread-local array, and a closest-hit walk keeps its winner's index and recomputes the hit point when it ends, as the gather
// replay always did (the primitive tests are deterministic functions of (primitive, ray)).
//
// Per pass p >= 1 the host launches k_trace_slice(p, 0, R), (p, 1, R), ... (p, S - 1, unbounded), then k_shade_pass(p) (engine.cu),
// which takes the hits from hitWords.  Slice 0 reads the pass queue; slice s > 0 reads continuation queue (s - 1) & 1 and writes
// queue s & 1 (the host zeroes a queue's counters before the slice that fills it).
#pragma once
#include "tracequeue.cuh"

#ifndef SLICE_STACK
#define SLICE_STACK 16 // stack entries a parked walk carries (model: at most 13 after 24 visits); a deeper one goes to the ordered walk
#endif
#define SLICE_WORDS (4 + 2 * SLICE_STACK) // best, sp, candidates, winner; then (ref, entry t) per stack entry

struct SliceParams
{
    float* state;            // [SLICE_WORDS][pathStride]
    float4* cand;            // [GATHER_CAP][pathStride]: (index, distance, leaf entry t, leaf)
    int* queues;             // [2][pathStride] continuation queues (path slots)
    unsigned int* counters;  // [4]: pushed, handed out — per continuation queue
};
__constant__ SliceParams cSlice;

__global__ void __launch_bounds__(128, TRACE_MIN_CTAS) k_trace_slice(const int pass, const int slice, const int rounds)
{
    const int lane = threadIdx.x & 31;
    const size_t stride = cP.pathStride;
    const int* __restrict__ qin;
    unsigned int count;
    unsigned int* handed;
    if (slice == 0)
    {
        const int qid = passQueue(pass);
        qin = cP.pathQueues + (size_t)qid * stride;
        count = cP.queueCounters[2 * qid];
        handed = cP.queueCounters + 2 * qid + 1;
    }
    else
    {
        const int b = (slice - 1) & 1;
        qin = cSlice.queues + (size_t)b * stride;
        count = cSlice.counters[2 * b];
        handed = cSlice.counters + 2 * b + 1;
    }
    int* __restrict__ qout = cSlice.queues + (size_t)(slice & 1) * stride;
    unsigned int* pushed = cSlice.counters + 2 * (slice & 1);
    const float minDistance0 = (pass < 2) ? cSI.viewDistance : cSI.viewDistance / (pass + 1);
    const float eps = cSI.geometryEpsilon;
    const float4* __restrict__ leafRecs = cS.leafRecs;
    const int* __restrict__ metas = cS.meta;
    const float4* __restrict__ nodes = cS.uwnodes;
    const int nbMain = cS.nbUWide;
    __shared__ int2 s_stack[SM_STACK * WALK_THREADS];
    int stackRef[UN_STACK - SM_STACK];
    float stackT[UN_STACK - SM_STACK];
    WalkStack st;
    st.sm = s_stack + threadIdx.x; st.lref = stackRef; st.lt = stackT;

    while (true)
    {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(handed, 32u);
        base = __shfl_sync(FULL_MASK, base, 0);
        if (base >= count) break;
        const bool has = base + lane < count;
        const size_t slot = has ? (size_t)qin[base + lane] : 0;
        float* sw = cSlice.state + slot;
        float4* cand = cSlice.cand + slot;

        // ---- the ray, and the walk's state: fresh in slice 0, restored afterwards
        const float* pw = cP.pathWords + slot;
        const float3 o = f3(pw[0], pw[stride], pw[2 * stride]);
        const float3 tgt = f3(pw[3 * stride], pw[4 * stride], pw[5 * stride]);
        const int currentMaterialId = __float_as_int(pw[7 * stride]);
        const float3 d = tgt - o;
        Ray r;
        makeRay(r, o, d);
        NodeRay q;
        nodeRay(q, r);
        const float len2 = dot(d, d);
        const int mode = (len2 >= 1.0002f) ? UW_CLOSEST : UW_GATHER;
        const float invLen = rsqrtf(len2) * 1.0001f;
        float best = minDistance0;
        int sp = 0, n = 0, winner = -1;
        if (has)
        {
            if (slice == 0)
            {
                st.push(0, 0, 0.f); sp = 1;
                if (cS.nbUX > 0) { st.push(1, nbMain, -3.0e38f); sp = 2; }
            }
            else
            {
                best = sw[0];
                sp = __float_as_int(sw[stride]);
                n = __float_as_int(sw[2 * stride]);
                winner = __float_as_int(sw[3 * stride]);
                for (int k = 0; k < sp; ++k) st.push(k, __float_as_int(sw[(size_t)(4 + 2 * k) * stride]), sw[(size_t)(5 + 2 * k) * stride]);
            }
        }
        // both bounds follow from the best distance (unorderedWalk keeps them in step the same way; for a gather ray
        // |direction| < 1, so window * invLen >= minDistance0 until a candidate is found)
        float window = fminf(minDistance0, GATHER_WINDOW * best);
        float cullT = (mode == UW_CLOSEST) ? fminf(minDistance0, best * invLen) : fminf(minDistance0, window * invLen);
        bool overflow = false;
        int budget = has ? rounds : 0;

        // ---- at most `rounds` node visits; a leaf in hand is always finished
        while (sp > 0 && budget > 0 && !overflow)
        {
            int cur = WIDE_NONE;
            while (sp > 0 && budget > 0)
            {
                --sp;
                int ref;
                float tEntry;
                st.pop(sp, ref, tEntry);
                if (tEntry > cullT) continue; // the bound shrank since this entry was pushed
                if (ref < 0) { cur = ref; break; }
                --budget;
                if (!unorderedStep(nodes, ref, q, (ref >= nbMain) ? 0.f : cullT, st, sp)) { overflow = true; sp = 0; }
            }
            if (overflow || cur == WIDE_NONE) break;
            const bool behind = ((~cur) & 0x40000000) != 0; // from the point-query tree
            const int idx = (~cur) & 0x3FFFFFFF;
            const int meta = __ldg(metas + idx);
            const int fast = PM_FAST(meta);
            if (!(fast == 0 || (fast == 1 && currentMaterialId != PM_MATERIAL(meta)))) continue;
            float3 I;
            int flags;
            float planeShadow;
            if (!primitiveTest(idx, meta, r, I, flags, planeShadow)) continue;
            const float distance = length(I - r.o);
            if (!(distance > eps)) continue;
            if ((dot(I - r.o, r.d) < 0.f) != behind) continue; // hits behind the origin (cylinders/cones only) come from the point query
            const int leaf = __ldg(cS.primLeaf + idx);
            const float4 lo = __ldg(leafRecs + 2 * leaf);
            const float4 hi = __ldg(leafRecs + 2 * leaf + 1);
            float leafT;
            if (!slabT(lo, hi, r, (mode == UW_CLOSEST) ? 3.0e38f : minDistance0, leafT)) continue;
            if (mode == UW_CLOSEST)
            {
                if (distance < best || (distance == best && winner >= 0 && idx < winner))
                {
                    best = distance;
                    winner = idx;
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
                    {
                        const float4 c = cand[(size_t)j * stride];
                        if (c.y <= window) { cand[(size_t)m2 * stride] = c; ++m2; }
                    }
                    n = m2;
                }
                if (n == GATHER_CAP) overflow = true;
                else
                {
                    cand[(size_t)n * stride] = make_float4(__int_as_float(idx), distance, leafT, __int_as_float(leaf));
                    ++n;
                }
            }
        }
        __syncwarp();

        // ---- unfinished walks are parked and queued for the next slice; one that cannot be parked takes the ordered walk now
        bool unfinished = has && !overflow && sp > 0;
        if (unfinished && sp > SLICE_STACK) { unfinished = false; overflow = true; }
        if (unfinished)
        {
            sw[0] = best;
            sw[stride] = __int_as_float(sp);
            sw[2 * stride] = __int_as_float(n);
            sw[3 * stride] = __int_as_float(winner);
            for (int k = 0; k < sp; ++k)
            {
                int ref;
                float tEntry;
                st.pop(k, ref, tEntry);
                sw[(size_t)(4 + 2 * k) * stride] = __int_as_float(ref);
                sw[(size_t)(5 + 2 * k) * stride] = tEntry;
            }
        }
        {
            const unsigned int m = __ballot_sync(FULL_MASK, unfinished);
            if (m != 0)
            {
                unsigned int at = 0;
                if (lane == __ffs(m) - 1) at = atomicAdd(pushed, (unsigned int)__popc(m));
                at = __shfl_sync(FULL_MASK, at, __ffs(m) - 1);
                if (unfinished) qout[at + __popc(m & ((1u << lane) - 1u))] = (int)slot;
            }
        }

        // ---- finished walks: replay (gather), fallback, hit record
        if (has && !unfinished)
        {
            Hit out;
            out.prim = -1; out.p = f3(0.f, 0.f, 0.f); out.flags = 0;
            if (overflow) out = closestHitWide(r.o, tgt, pass, currentMaterialId);
            else
            {
                if (mode == UW_GATHER)
                {
                    // replay in array order, as unorderedWalk() does
                    float m = minDistance0;
                    bool leafPass = false;
                    int prevLeaf = -1, last = -1;
                    winner = -1;
                    for (int ps = 0; ps < n; ++ps)
                    {
                        int bi = 0x7fffffff, bLeaf = -1;
                        float bD = 0.f, bT = 0.f;
                        for (int j = 0; j < n; ++j)
                        {
                            const float4 c = cand[(size_t)j * stride];
                            const int ci = __float_as_int(c.x);
                            if (ci > last && ci < bi && c.y <= window) { bi = ci; bD = c.y; bT = c.z; bLeaf = __float_as_int(c.w); }
                        }
                        if (bi == 0x7fffffff) break;
                        last = bi;
                        if (bLeaf != prevLeaf)
                        {
                            leafPass = bT < m;
                            prevLeaf = bLeaf;
                        }
                        if (leafPass && bD < m) { m = bD; winner = bi; }
                    }
                }
                if (winner >= 0)
                {
                    float3 I;
                    int flags;
                    float planeShadow;
                    primitiveTest(winner, __ldg(metas + winner), r, I, flags, planeShadow); // deterministic: the hit point it had when it was found
                    out.prim = winner; out.p = I; out.flags = flags;
                }
            }
            float* hw = cP.hitWords + slot;
            hw[0] = __int_as_float(out.prim); hw[stride] = out.p.x; hw[2 * stride] = out.p.y; hw[3 * stride] = out.p.z;
            hw[4 * stride] = __int_as_float(out.flags);
        }
        __syncwarp();
    }
}
