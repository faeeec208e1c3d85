// tracewave.cuh — the walk as its own kernel, in phases: the unit walk of trace.cuh with (a) the rays of a queue handed to the lanes
// one at a time and (b) node rounds, primitive tests and ray hand-over run as three separate warp-wide phases.
// Replaces, like the walks it is built from, /root/reference/solr/engines/cuda/GeometryIntersections.cuh:667-772 (closest hit)
// and :798-907 (shadows).
//
// What the source-level profile of the staged kernels says (profiles/r02_ncu_unit_src.md): in a bounce pass 62 % of the warp
// instructions are node rounds run at 6 of 32 lanes — a walk's length varies 10x between rays (mean 22 node visits, up to 160) and
// a warp whose 32 lanes each own one ray runs as long as its longest walk — and another 20 % are primitive tests run at 1.3-2.1
// lanes, because every loop iteration of the unit walk runs the node code and then the primitive code for whichever lanes are
// there.  Handing a finished lane the next ray inside the staged kernels (a pool of 64 rays per warp in shared memory, round 2
// "pooled stages") raised the node rounds to 13 lanes and still lost (7.2-7.5 vs 4.4 ms per frame, with the phases below as well
// as without): a kernel whose warps sit in the walk, in the shader and in the hand-over at the same time lives on instruction
// fetches from L2 (8.9 no-instruction stalls per issue, issue slots 27 % busy; the instruction caches are ~6 KB per scheduler and
// 32 KB per SM, the staged kernels 200 KB).  Hence a kernel that contains nothing but the walk, and inside it the phases — each
// of them a small piece of code run by many lanes:
//   NODE    every lane with a node runs node rounds; a lane that reaches a primitive puts it in its queue (PH_K entries per lane
//           in shared memory) and goes on with its stack; the phase ends when PH_STUCK lanes can do nothing more (queue full,
//           walk over) or nobody can;
//   LEAF    every lane tests the primitives in its queue (test, reference leaf box, acceptance rule: as the unit walk);
//   REFILL  every lane whose walk is over stores the result and takes the next ray of the queue (a chunk of WAVE_CHUNK rays per
//           warp at a time).
// The node phase keeps only what a node round needs in registers (the ray in slab form, the bound, the stack pointer); the rest
// of a walk's state (the ray, best hit, window, candidate count) waits in shared memory for the leaf phase.
// Testing a primitive late only delays the shrinking of the bound: a walk visits a few more nodes, the candidates and acceptance
// rules — and therefore the results — are those of unorderedWalk() (closest: minimum distance, ties to the lowest index; gather:
// every candidate inside the final window, replayed in array order; shadow: any blocker).  A ray's result does not depend on the
// lane that walks it.
#pragma once

#ifndef WAVE_CHUNK
#define WAVE_CHUNK 128 // rays a warp takes from the queue at a time
#endif
#ifndef WAVE_GATHER_CAP
#define WAVE_GATHER_CAP 32 // candidates kept per bounce ray (global scratch); a fuller list sends the ray to the ordered walk
#endif
#ifndef PH_K
#define PH_K 2 // primitives a lane can hold back for the next leaf phase
#endif
#ifndef PH_STUCK
#define PH_STUCK 8 // lanes that cannot go on before the warp leaves the node phase
#endif
#define WAVE_CLOSEST 0
#define WAVE_SHADOW 1
#define WAVE_NONE (-1)
#define WAVE_OVERFLOW (-2)
#define HIT_WORDS 5    // primitive | WAVE_*, hit point xyz, flags
#define SHADOW_WORDS 9 // origin, direction, lamp id, object id, result

// the walk a lane is on: its ray (origin, material id | lamp id) (direction, - | object id), and between leaf phases
// (primitive, hit point) (flags, candidates, best distance, window)
__shared__ float4 s_waveRay[2][WALK_THREADS];
__shared__ float4 s_waveRun[2][WALK_THREADS];
__shared__ int s_leafQ[PH_K][WALK_THREADS];
__shared__ unsigned int s_waveNext[WALK_THREADS / 32], s_waveEnd[WALK_THREADS / 32];

#define PF_OVERFLOW 1
#define PF_SHADOWED 2
#define PF_GATHER 4 // a bounce ray (|direction| < 1): candidates are gathered and replayed

// The LEAF phase for one lane: tests the nq primitives of its queue against its ray and applies the acceptance rule of its ray
// class — the code of unorderedWalk()'s leaf branch, with the walk's state read from and left in s_waveRun.  Every lane of the
// warp calls (lanes with nq == 0 idle through it).
__device__ __forceinline__ void waveLeaves(const int kind, const float minDistance0, float4* const list, const int nq, int& cur, int& nSpill,
                                           int& fl, float& cullT)
{
    const float eps = cSI.geometryEpsilon;
    const bool extended = cSI.extendedGeometry != 0;
    const float4* __restrict__ recs = cS.primRecs;
    Ray r;
    r.o = r.d = r.nd = r.inv = f3(0.f, 0.f, 0.f);
    int p0 = 0, p1 = 0, mode = UW_CLOSEST, n = 0;
    float invLen = 0.f, lenOL = 0.f, best = 0.f, window = 0.f;
    Hit hit;
    hit.prim = -1; hit.p = f3(0.f, 0.f, 0.f); hit.flags = 0;
    int m = nq; // entries of the queue still to test
    bool dirty = false;
    if (m > 0)
    {
        const float4 A = s_waveRay[0][threadIdx.x], B = s_waveRay[1][threadIdx.x];
        p0 = __float_as_int(A.w); p1 = __float_as_int(B.w);
        makeRay(r, f3(A.x, A.y, A.z), f3(B.x, B.y, B.z));
        const float len2 = dot(r.d, r.d);
        invLen = rsqrtf(len2) * 1.0001f; // world distance -> t, with slack so culling stays conservative
        lenOL = sqrtf(len2);             // shadow: distance to the lamp (length(O_L), :877)
        mode = (kind == WAVE_SHADOW) ? UW_SHADOW : ((fl & PF_GATHER) ? UW_GATHER : UW_CLOSEST);
        if (kind != WAVE_SHADOW)
        {
            const float4 R0 = s_waveRun[0][threadIdx.x], R1 = s_waveRun[1][threadIdx.x];
            hit.prim = __float_as_int(R0.x); hit.p = f3(R0.y, R0.z, R0.w);
            hit.flags = __float_as_int(R1.x); n = __float_as_int(R1.y); best = R1.z; window = R1.w;
        }
    }
#pragma unroll 1
    for (int j = 0; j < PH_K; ++j)
    {
        if (!__any_sync(FULL_MASK, j < m)) break;
        if (!(j < m)) continue;
        const int ref = s_leafQ[j][threadIdx.x];
        const int idx = (~ref) & 0x3FFFFFFF;
        const bool behind = ((~ref) & 0x40000000) != 0; // from the point-query tree
        const float4* item = recs + (size_t)PRIM_REC_F4 * idx;
        float4 a0, a1, a2, a3, a4, a5;
        ldNode256(item, a0, a1); ldNode256(item + 2, a2, a3); ldNode256(item + 4, a4, a5);
        const int meta = __float_as_int(a3.w);
        const int fast = PM_FAST(meta);
        bool test;
        if (mode == UW_SHADOW)
        {
            const int origIndex = __float_as_int(a5.w);
            const int type = extended ? PM_TYPE(meta) : B200_PT_TRIANGLE;
            // objectId is a compacted index compared with an original id — as the reference does (:829)
            test = fast == 0 && origIndex != p0 && origIndex != p1 && type != B200_PT_CAMERA && type != B200_PT_ENVIRONMENT &&
                   !(type == B200_PT_TRIANGLE && cSI.doubleSidedTriangles);
        }
        else
            test = fast == 0 || (fast == 1 && p0 != PM_MATERIAL(meta));
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
        float leafT;
        if (!slabT(a4, a5, r, (mode == UW_CLOSEST) ? 3.0e38f : minDistance0, leafT)) continue;
        if (mode == UW_SHADOW)
        {
            if (distance < lenOL) { fl |= PF_SHADOWED; nSpill = 0; cur = WALK_DONE; m = 0; }
        }
        else if (mode == UW_CLOSEST)
        {
            if (distance < best || (distance == best && hit.prim >= 0 && idx < hit.prim))
            {
                best = distance;
                hit.prim = idx; hit.p = I; hit.flags = flags;
                cullT = fminf(minDistance0, best * invLen);
                dirty = true;
            }
        }
        else if (distance < minDistance0 && distance <= window)
        {
            dirty = true;
            if (distance < best)
            {
                best = distance;
                window = fminf(minDistance0, GATHER_WINDOW * best);
                cullT = fminf(minDistance0, window * invLen);
            }
            if (n == WAVE_GATHER_CAP)
            {
                int m2 = 0; // full: drop what fell out of the window meanwhile
                for (int i = 0; i < n; ++i)
                {
                    const float4 c = __ldcg(list + i);
                    if (c.y <= window) { __stcg(list + m2, c); ++m2; }
                }
                n = m2;
            }
            if (n == WAVE_GATHER_CAP) { fl |= PF_OVERFLOW; nSpill = 0; cur = WALK_DONE; m = 0; }
            else
            {
                // appended in visiting order; the replay picks them in array order
                __stcg(list + n, f4(__int_as_float(idx), distance, leafT, a4.w));
                ++n;
            }
        }
    }
    if (dirty)
    {
        s_waveRun[0][threadIdx.x] = f4(__int_as_float(hit.prim), hit.p.x, hit.p.y, hit.p.z);
        s_waveRun[1][threadIdx.x] = f4(__int_as_float(hit.flags), __int_as_float(n), best, window);
    }
}

// End of a bounce ray's walk: its candidates replayed in array order (selection by ascending index: the list is short; a cylinder
// listed twice by the point query is taken once).  Primitives of one leaf are contiguous and share the leaf's fate, decided when
// the leaf is reached (before any of its primitives): t_min(leaf) < closest-so-far.  The winner's hit point is recomputed (same
// arithmetic as when it was gathered).
__device__ __forceinline__ Hit waveReplay(const float4* const list, const int n, const float window, const float minDistance0)
{
    float m = minDistance0;
    bool leafPass = false;
    int prevLeaf = -1, winner = -1, last = -1;
    for (int pass = 0; pass < n; ++pass)
    {
        int bi = 0x7fffffff;
        float4 bc = f4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < n; ++j)
        {
            const float4 c = __ldcg(list + j);
            const int ci = __float_as_int(c.x);
            if (ci > last && ci < bi && c.y <= window) { bi = ci; bc = c; }
        }
        if (bi == 0x7fffffff) break;
        last = bi;
        if (__float_as_int(bc.w) != prevLeaf)
        {
            leafPass = bc.z < m;
            prevLeaf = __float_as_int(bc.w);
        }
        if (leafPass && bc.y < m) { m = bc.y; winner = bi; }
    }
    Hit hit;
    hit.prim = -1; hit.p = f3(0.f, 0.f, 0.f); hit.flags = 0;
    if (winner >= 0)
    {
        const float4 A = s_waveRay[0][threadIdx.x], B = s_waveRay[1][threadIdx.x];
        Ray r;
        makeRay(r, f3(A.x, A.y, A.z), f3(B.x, B.y, B.z));
        const float4* item = cS.primRecs + (size_t)PRIM_REC_F4 * winner;
        float4 a0, a1, a2, a3;
        ldNode256(item, a0, a1); ldNode256(item + 2, a2, a3);
        float3 I;
        int flags;
        float planeShadow;
        primitiveTestRegs(a0, a1, a2, a3, winner, __float_as_int(a3.w), r, I, flags, planeShadow);
        hit.prim = winner; hit.p = I; hit.flags = flags;
    }
    return hit;
}

// The walk kernel's body.  kind WAVE_CLOSEST: the rays (origin, target, material) of path queue `qid` from pathWords -> hitWords;
// WAVE_SHADOW: the rays of shadowWords listed in the shadow queue -> shadowWords[8].  handOut: this launch's own hand-out counter.
__device__ __forceinline__ void waveWalk(const int kind, const int iteration, const int qid, unsigned int* const handOut)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int count = cP.queueCounters[2 * qid];
    const int* __restrict__ queue = cP.pathQueues + (size_t)qid * cP.pathStride;
    const size_t stride = cP.pathStride;
    if (lane == 0) { s_waveNext[warp] = 0u; s_waveEnd[warp] = 0u; }
    __syncwarp();
    const float minDistance0 = (iteration < 2) ? cSI.viewDistance : cSI.viewDistance / (iteration + 1);
    const float4* __restrict__ nodes = cS.uwnodes;
    const int nbMain = cS.nbUWide;
    int2* const s_stack = s_walkStack;
    float4* const list = cP.gatherScratch + ((size_t)blockIdx.x * WALK_THREADS + threadIdx.x) * WAVE_GATHER_CAP; // this lane's candidates while it walks
    unsigned int sp = 0, spLimit = 0;
    int2 spillBuf[WALK_SPILL];
    int nSpill = 0;
    // the lane's walk, as far as the node phase needs it
    int slot = -1, cur = WALK_DONE, nq = 0, fl = 0;
    NodeRay q;
    q.ix = q.iy = q.iz = 1.f; q.nox = q.noy = q.noz = 0.f;
    float cullT = 0.f;
    bool queueLeft = count > 0u; // rays of the queue not handed out yet

    while (true)
    {
        // ------------------------------------------------------------------ NODE phase
        while (true)
        {
            if (cur == WALK_POP)
            {
                sp -= WALK_STACK_STRIDE;
                const int2 e = stackGet(sp);
                cur = (__int_as_float(e.y) > cullT) ? WALK_POP : e.x; // the bound shrank since this entry was pushed
            }
            if (cur == WALK_DONE && nSpill != 0)
            {
                walkUnspill(s_stack, spillBuf, nSpill, sp);
                cur = WALK_POP;
            }
            bool isLeaf = cur < 0 && cur != WALK_DONE;
            if (isLeaf && nq < PH_K)
            {
                s_leafQ[nq][threadIdx.x] = cur;
                ++nq;
                cur = WALK_POP;
                isLeaf = false;
            }
            const bool isNode = cur >= 0 && cur != WALK_POP;
            // a lane is stuck when its queue is full, or its walk is over and the next phases have something for it
            const bool stuck = isLeaf || (cur == WALK_DONE && (slot >= 0 || queueLeft));
            const unsigned int goM = __ballot_sync(FULL_MASK, isNode || cur == WALK_POP);
            const unsigned int stuckM = __ballot_sync(FULL_MASK, stuck);
            if (goM == 0u) break;
            {
                const int alive = __popc(goM | stuckM);
                const int limit = min(PH_STUCK, (alive + 3) >> 2);
                if (__popc(stuckM) >= limit) break;
            }
            if (isNode)
            {
                if (sp >= spLimit)
                {
                    if (!walkSpill(s_stack, spillBuf, nSpill, sp)) { fl |= PF_OVERFLOW; nSpill = 0; nq = 0; cur = WALK_DONE; continue; } // degenerate tree
                }
                DBG_ADD(5, 1);
                const float4* item = nodes + (size_t)8 * cur;
                float4 a0, a1, a2, a3, a4, a5, a6, a7;
                ldNode256(item, a0, a1); ldNode256(item + 2, a2, a3); ldNode256(item + 4, a4, a5); ldNode256(item + 6, a6, a7);
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
            }
        }
        // ------------------------------------------------------------------ LEAF phase
        if (__any_sync(FULL_MASK, nq > 0))
        {
            if (lane == 0) { DBG_ADD(2, 1); }
            DBG_ADD(4, nq);
            waveLeaves(kind, minDistance0, list, nq, cur, nSpill, fl, cullT);
            nq = 0;
        }
        // ------------------------------------------------------------------ REFILL phase
        const bool fin = cur == WALK_DONE && nSpill == 0;
        if (fin && slot >= 0)
        {
            // this lane's walk is over: leave the result
            if (kind == WAVE_SHADOW)
                cP.shadowWords[8 * stride + slot] = (fl & PF_OVERFLOW) ? -1.f : ((fl & PF_SHADOWED) ? fmaxf(0.f, cSI.shadowIntensity) : 0.f);
            else
            {
                const float4 R0 = s_waveRun[0][threadIdx.x], R1 = s_waveRun[1][threadIdx.x];
                Hit hit;
                hit.prim = __float_as_int(R0.x); hit.p = f3(R0.y, R0.z, R0.w); hit.flags = __float_as_int(R1.x);
                if (fl & PF_OVERFLOW) { hit.prim = WAVE_OVERFLOW; }
                else if (fl & PF_GATHER) hit = waveReplay(list, __float_as_int(R1.y), R1.w, minDistance0);
                float* hw = cP.hitWords + slot;
                hw[0] = __int_as_float(hit.prim); hw[stride] = hit.p.x; hw[2 * stride] = hit.p.y; hw[3 * stride] = hit.p.z;
                hw[4 * stride] = __int_as_float(hit.flags);
            }
            slot = -1;
        }
        int newSlot = -1;
        while (queueLeft)
        {
            const unsigned int want = __ballot_sync(FULL_MASK, fin && newSlot < 0);
            if (want == 0u) break;
            unsigned int nx = s_waveNext[warp], en = s_waveEnd[warp];
            __syncwarp();
            if (nx >= en)
            {
                // the warp's chunk has been handed out: the next one
                unsigned int base = 0;
                if (lane == 0) base = atomicAdd(handOut, (unsigned int)WAVE_CHUNK);
                base = __shfl_sync(FULL_MASK, base, 0);
                if (base >= count) { queueLeft = false; break; }
                nx = base; en = min(base + (unsigned int)WAVE_CHUNK, count);
            }
            const unsigned int k = nx + __popc(want & ((1u << lane) - 1u));
            if (lane == 0) { s_waveNext[warp] = min(nx + (unsigned int)__popc(want), en); s_waveEnd[warp] = en; }
            __syncwarp();
            if (fin && newSlot < 0 && k < en) newSlot = queue[k];
        }
        if (newSlot >= 0)
        {
            slot = newSlot;
            float3 o, d;
            int p0, p1;
            if (kind == WAVE_SHADOW)
            {
                const float* w = cP.shadowWords + slot;
                o = f3(w[0], w[stride], w[2 * stride]); d = f3(w[3 * stride], w[4 * stride], w[5 * stride]);
                p0 = __float_as_int(w[6 * stride]); p1 = __float_as_int(w[7 * stride]);
            }
            else
            {
                const float* w = cP.pathWords + slot;
                o = f3(w[0], w[stride], w[2 * stride]);
                d = f3(w[3 * stride], w[4 * stride], w[5 * stride]) - o;
                p0 = __float_as_int(w[7 * stride]); p1 = 0;
            }
            s_waveRay[0][threadIdx.x] = f4(o.x, o.y, o.z, __int_as_float(p0));
            s_waveRay[1][threadIdx.x] = f4(d.x, d.y, d.z, __int_as_float(p1));
            Ray r;
            makeRay(r, o, d);
            nodeRay(q, r);
            const float len2 = dot(r.d, r.d);
            const float invLen = rsqrtf(len2) * 1.0001f; // world distance -> t, with slack so culling stays conservative
            fl = (kind == WAVE_SHADOW || len2 >= 1.0002f) ? 0 : PF_GATHER;
            // entry-t bound for nodes: never beyond the reference's own t_min < closest-so-far test; a shadow blocker lies before the lamp (t ~ 1)
            cullT = (kind == WAVE_SHADOW) ? fminf(minDistance0, UW_SHADOW_TLIMIT) : ((fl & PF_GATHER) ? minDistance0 : fminf(minDistance0, minDistance0 * invLen));
            s_waveRun[0][threadIdx.x] = f4(__int_as_float(-1), 0.f, 0.f, 0.f);
            s_waveRun[1][threadIdx.x] = f4(__int_as_float(0), __int_as_float(0), minDistance0, minDistance0);
            walkStart(s_stack, nbMain, cS.nbUX > 0, spLimit, sp, cur);
            if (kind == WAVE_SHADOW && !(0.f < cSI.shadowIntensity)) cur = WALK_DONE; // nothing can shade: result 0
        }
        if (__ballot_sync(FULL_MASK, slot >= 0) == 0u) break;
    }
}
