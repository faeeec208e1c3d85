// tracegroup.cuh — the group walk (round 2): the order-independent walks of trace.cuh with the LANES of a warp shared out
// differently.  Replaces, like the walks it is built from, the reference's box-list walk
// (/root/reference/solr/engines/cuda/GeometryIntersections.cuh:667-772 closest hit, :798-907 shadows).
//
// What ncu said about one-ray-per-lane walks (profiles/r02_ncu_unit_v1.json): issue slots 50-60 % busy, but a warp instruction
// carries 14 (primary rays), 5-7 (bounce rays) of 32 lanes — walk lengths differ 10x inside a warp (mean 21 node visits, up to
// 162), and every scheme that re-balanced whole rays between lanes (per-lane refill, sliced walks, hand-overs) cost more than
// the idle lanes.  Here the unit of divergence is no longer a lane:
//   * a ray is walked by a GROUP of GW_LANES (= children per node) neighbouring lanes: in a node round every lane tests ONE
//     child box (one 32-byte load per lane from a child-major node record), the group's nearest hit inner child comes out of
//     a shuffle min-reduction over the integer keys of trace.cuh, the other hit children are pushed by their own lanes at
//     ballot-ranked positions of the group's stack in shared memory: no sort, no serial push, no idle lane in a node round;
//   * a warp takes the (up to) 32 rays of its lanes, publishes them in shared memory and hands them to its 32 / GW_LANES
//     groups one after the other — a group that finishes takes the next ray, so a long walk delays one group, not a warp;
//   * primitives are not tested where they are found: a hit primitive child goes into a warp-wide queue (primitive, ray
//     slot), and when the queue holds enough entries every lane takes ONE entry — whichever ray it belongs to — loads that
//     ray and the primitive's 96-byte record and runs the test, so primitive tests run at full warps too.  Results meet in
//     shared memory: a 64-bit atomicMin of (distance bits : primitive index) per ray for closest hits — which is the reference
//     result's tie rule, lowest index among equal distances —, a flag for shadow rays, a candidate list for bounce rays.
//     The price is that a group culls with a bound that is a few rounds old (more node visits, same result).
// Candidates, acceptance rules and therefore results are those of unorderedWalk() (trace.cuh): a primitive is a candidate if
// its test hits beyond geometryEpsilon on the right side of the origin and its REFERENCE leaf box passes the reference's
// slab arithmetic; closest-hit rays take the minimum, bounce rays (|direction| < 1) gather the candidates within
// GATHER_WINDOW x the closest distance and replay the reference's accept / reject sequence in array order, shadow rays stop
// at the first blocker.  A ray whose stack or candidate list overflows is handed back for the ordered walk (prim = -2).
// Every lane of the warp must call groupWalk() (need = false for a lane without a ray).
#pragma once

#ifndef GW_LANES
#define GW_LANES UW_WIDTH // lanes per ray = children per node (4, or 8 with -DUW_WIDTH=8)
#endif
#define GW_GROUPS (32 / GW_LANES)
#define GW_WARPS (WALK_THREADS / 32)
#ifndef GW_STACK
#define GW_STACK 32 // stack entries per group
#endif
#ifndef GW_QT
#define GW_QT 24 // primitive tests queued before the warp runs them ...
#endif
#ifndef GW_QPER
#define GW_QPER 3 // ... or this many per group still walking, whichever is less
#endif
#define GW_QCAP 64 // a flush leaves < 32 entries and a round adds <= 32
#ifndef GW_GATHER_CAP
#define GW_GATHER_CAP 32 // candidates per bounce ray (global scratch, 16 bytes each)
#endif
#define GW_DONE ((int)0x80000000) // the group's walk is over (also the stack's sentinel)
#define GW_IDLE ((int)0x80000001) // no ray left for this group
#define GW_POP 0x7fffffff
#define GW_NOPRIM 0x7fffffffu
#if GW_LANES == 4
#define GW_LEADERS 0x11111111u
#else
#define GW_LEADERS 0x01010101u
#endif
#define GW_STACK_STRIDE (GW_GROUPS * 8) // bytes between two entries of one group: entries of the groups are interleaved

// per warp, in shared memory
__shared__ float4 s_ray[GW_WARPS][32 * 3];          // (origin, minDistance0) (direction, mode) (material, lamp, object, -)
__shared__ unsigned long long s_res[GW_WARPS][32];  // high word: best distance (float bits); low word: primitive / count / blocked
__shared__ float4 s_hit[GW_WARPS][32];              // closest hit: point, flags
__shared__ int2 s_stack[GW_WARPS][GW_STACK * GW_GROUPS];
__shared__ int2 s_queue[GW_WARPS][GW_QCAP];         // (primitive ref, ray slot)
__shared__ unsigned int s_ovf[GW_WARPS];            // ray slots that overflowed

// One batch of queued primitive tests: lane k < nProc takes queue entry first + k.  Every lane of the warp calls.
__device__ __noinline__ void gwLeafBatch(const int first, const int nProc, float4* const scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float eps = cSI.geometryEpsilon;
    const bool extended = cSI.extendedGeometry != 0;
    const float4* __restrict__ recs = cS.primRecs;
    __syncwarp();
    bool accepted = false;
    unsigned long long myKey = 0ull;
    float4 myHit = f4(0.f, 0.f, 0.f, 0.f);
    int eslot = 0;
    DBG_ADD(6, lane == 0 ? 1 : 0);
    if (lane < nProc)
    {
        DBG_ADD(7, 1);
        const int2 e = s_queue[warp][first + lane];
        eslot = e.y;
        const bool behind = ((~e.x) & 0x40000000) != 0; // from the point-query tree
        const int idx = (~e.x) & 0x3FFFFFFF;
        const float4 A = s_ray[warp][3 * eslot], B = s_ray[warp][3 * eslot + 1], C = s_ray[warp][3 * eslot + 2];
        const int emode = __float_as_int(B.w);
        const float eMinD0 = A.w;
        const float4* item = recs + (size_t)PRIM_REC_F4 * idx;
        float4 a0, a1, a2, a3, a4, a5;
        ldNode256(item, a0, a1); ldNode256(item + 2, a2, a3); ldNode256(item + 4, a4, a5);
        const int meta = __float_as_int(a3.w);
        const int fast = PM_FAST(meta);
        bool test;
        if (emode == UW_SHADOW)
        {
            const int origIndex = __float_as_int(a5.w);
            const int type = extended ? PM_TYPE(meta) : B200_PT_TRIANGLE;
            // objectId is a compacted index compared with an original id — as the reference does (:829)
            test = fast == 0 && origIndex != __float_as_int(C.y) && origIndex != __float_as_int(C.z) && type != B200_PT_CAMERA &&
                   type != B200_PT_ENVIRONMENT && !(type == B200_PT_TRIANGLE && cSI.doubleSidedTriangles);
        }
        else
            test = fast == 0 || (fast == 1 && __float_as_int(C.x) != PM_MATERIAL(meta));
        if (test)
        {
            Ray r;
            makeRay(r, f3(A.x, A.y, A.z), f3(B.x, B.y, B.z));
            float3 I;
            int flags;
            float planeShadow, leafT;
            // the reference only tests a primitive whose leaf box passes its slab test (:690); checked for hits only
            // (UW_CLOSEST: t_min(leaf) <= entry distance <= hit distance < closest-so-far whenever the hit would be accepted,
            //  so only the geometric part of the leaf test can reject it)
            if (primitiveTestRegs(a0, a1, a2, a3, idx, meta, r, I, flags, planeShadow))
            {
                const float distance = length(I - r.o);
                // hits behind the origin (cylinders/cones only) come from the point query
                if (distance > eps && ((dot(I - r.o, r.d) < 0.f) == behind) &&
                    slabT(a4, a5, r, (emode == UW_CLOSEST) ? 3.0e38f : eMinD0, leafT))
                {
                    unsigned int* const resw = reinterpret_cast<unsigned int*>(&s_res[warp][eslot]); // [0] low word, [1] high word
                    if (emode == UW_SHADOW)
                    {
                        if (distance < sqrtf(dot(r.d, r.d))) resw[0] = 1u; // before the lamp (length(O_L), :877)
                    }
                    else if (emode == UW_CLOSEST)
                    {
                        if (distance < eMinD0)
                        {
                            myKey = ((unsigned long long)__float_as_uint(distance) << 32) | (unsigned int)idx;
                            const unsigned long long old = atomicMin(&s_res[warp][eslot], myKey);
                            accepted = myKey < old;
                            myHit = f4(I.x, I.y, I.z, __int_as_float(flags));
                        }
                    }
                    else
                    {
                        const float best = __uint_as_float(*reinterpret_cast<volatile unsigned int*>(resw + 1));
                        if (distance < eMinD0 && distance <= fminf(eMinD0, GATHER_WINDOW * best))
                        {
                            atomicMin(resw + 1, __float_as_uint(distance));
                            const unsigned int pos = atomicAdd(resw, 1u);
                            if (pos < GW_GATHER_CAP)
                                __stcg(scratch + (size_t)eslot * GW_GATHER_CAP + pos, f4(__int_as_float(idx), distance, leafT, a4.w));
                            else
                                atomicOr(&s_ovf[warp], 1u << eslot);
                        }
                    }
                }
            }
        }
    }
    __syncwarp();
    // the lane that holds a ray's minimum after this batch leaves the hit point (equal keys are the same primitive twice)
    if (accepted && s_res[warp][eslot] == myKey) s_hit[warp][eslot] = myHit;
}

__device__ __noinline__ WalkOut groupWalk(const bool needIn, const int mode, const float3 rayOrigin, const float3 rayDir, const int iteration,
                                          const int currentMaterialId, const int lightId, const int objectId)
{
    WalkOut out;
    out.hit.prim = -1; out.hit.p = f3(0.f, 0.f, 0.f); out.hit.flags = 0; out.shadow = 0.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / GW_LANES, j = lane % GW_LANES;
    const unsigned int ltMask = (1u << lane) - 1u;
    const float minDistance0 = (iteration < 2) ? cSI.viewDistance : cSI.viewDistance / (iteration + 1);
    const float shadowLimit = cSI.shadowIntensity;
    const bool need = needIn && !(mode == UW_SHADOW && !(0.f < shadowLimit));
    unsigned int remaining = __ballot_sync(FULL_MASK, need); // rays not yet handed to a group
    if (remaining == 0) return out;
    if (need)
    {
        s_ray[warp][3 * lane] = f4(rayOrigin.x, rayOrigin.y, rayOrigin.z, minDistance0);
        s_ray[warp][3 * lane + 1] = f4(rayDir.x, rayDir.y, rayDir.z, __int_as_float(mode));
        s_ray[warp][3 * lane + 2] = f4(__int_as_float(currentMaterialId), __int_as_float(lightId), __int_as_float(objectId), 0.f);
        s_res[warp][lane] = ((unsigned long long)__float_as_uint(minDistance0) << 32) | (mode == UW_CLOSEST ? GW_NOPRIM : 0u);
    }
    if (lane == 0) s_ovf[warp] = 0u;
    const float4* __restrict__ nodes = cS.ugnodes;
    const float4* __restrict__ recs = cS.primRecs;
    const int nbMain = cS.nbUWide;
    const bool pointQuery = cS.nbUX > 0;
    float4* const scratch = cP.gatherScratch + ((size_t)(blockIdx.x * GW_WARPS + warp) * 32) * GW_GATHER_CAP;
    const unsigned int stackBase = (unsigned int)__cvta_generic_to_shared(&s_stack[warp][g]);
    const unsigned int spLimit = stackBase + (GW_STACK - GW_LANES) * GW_STACK_STRIDE;

    // the group's ray
    int cur = GW_DONE, slot = 0, gmode = UW_CLOSEST;
    unsigned int sp = stackBase;
    NodeRay q;
    q.ix = q.iy = q.iz = 1.f; q.nox = q.noy = q.noz = 0.f;
    float gMinD0 = 0.f, invLen = 0.f, cullT = 0.f;
    int qn = 0; // queued primitive tests (warp-uniform)

#ifndef GW_MAX_ROUNDS
#define GW_MAX_ROUNDS (1 << 22) // no walk takes this many rounds: a guard against spinning on a corrupt tree, not a tuning knob
#endif
    int rounds = 0;
    DBG_ADD(2, need ? 1 : 0);
    while (true)
    {
        __syncwarp();
        if (++rounds > GW_MAX_ROUNDS) { if (lane == 0) s_ovf[warp] = 0xffffffffu; break; }
        DBG_ADD(4, lane == 0 ? 1 : 0);
        // ---- groups without a walk take the next rays
        const unsigned int dm = __ballot_sync(FULL_MASK, cur == GW_DONE) & GW_LEADERS;
        if (dm != 0u)
        {
            if (remaining != 0u)
            {
                const int rank = __popc(dm & ((1u << (g * GW_LANES)) - 1u));
                const unsigned int take = __fns(remaining, 0, rank + 1);
                if (cur == GW_DONE && take != 0xffffffffu)
                {
                    slot = (int)take;
                    const float4 A = s_ray[warp][3 * slot], B = s_ray[warp][3 * slot + 1];
                    gmode = __float_as_int(B.w);
                    gMinD0 = A.w;
                    Ray r;
                    makeRay(r, f3(A.x, A.y, A.z), f3(B.x, B.y, B.z));
                    nodeRay(q, r);
                    invLen = rsqrtf(dot(r.d, r.d)) * 1.0001f; // world distance -> t, with slack so culling stays conservative
                    // a shadow blocker lies before the lamp (t ~ 1); the others start from the reference's own t_min < closest-so-far
                    cullT = (gmode == UW_SHADOW) ? fminf(gMinD0, UW_SHADOW_TLIMIT) : fminf(gMinD0, gMinD0 * invLen);
                    sp = stackBase;
                    stackPut(sp, GW_DONE, 0); sp += GW_STACK_STRIDE; // sentinel: never behind the bound
                    if (pointQuery) { stackPut(sp, nbMain, 0); sp += GW_STACK_STRIDE; }
                    cur = 0;
                }
                int nTake = min(__popc(dm), __popc(remaining));
                for (int k = 0; k < nTake; ++k) remaining &= remaining - 1u;
            }
            if (cur == GW_DONE) cur = GW_IDLE;
        }
        const unsigned int am = __ballot_sync(FULL_MASK, cur != GW_IDLE);
        if (am == 0u && qn == 0) break;

        // ---- the bound, from what the primitive tests have found so far
        if (cur != GW_IDLE)
        {
            const unsigned long long res = s_res[warp][slot];
            const float best = __uint_as_float((unsigned int)(res >> 32));
            if (gmode == UW_SHADOW) { if ((unsigned int)res != 0u) cur = GW_DONE; }
            else cullT = fminf(gMinD0, ((gmode == UW_GATHER) ? fminf(gMinD0, GATHER_WINDOW * best) : best) * invLen);
        }
        // ---- pop (twice: an entry that has fallen behind the bound costs no round)
#pragma unroll
        for (int t = 0; t < 2; ++t)
            if (cur == GW_POP)
            {
                sp -= GW_STACK_STRIDE;
                const int2 e = stackGet(sp);
                cur = (__int_as_float(e.y) > cullT) ? GW_POP : e.x;
            }
        // ---- node round: one child per lane
        const bool isNode = cur >= 0 && cur != GW_POP;
        int key = KEY_MISS, ref = GW_DONE;
        if (isNode)
        {
            DBG_ADD(5, j == 0 ? 1 : 0);
            float4 a, b;
            ldNode256(nodes + ((size_t)cur * GW_LANES + j) * 2, a, b); // (lo.xyz, hi.x) (hi.yz, ref, -)
            ref = __float_as_int(b.z);
            const float tLimit = (cur >= nbMain) ? 0.f : cullT; // the point-query tree keeps the boxes that contain the origin
            const bool sx = q.ix < 0.f, sy = q.iy < 0.f, sz = q.iz < 0.f;
            const float tnx = __fmaf_rn(sx ? a.w : a.x, q.ix, q.nox), tfx = __fmaf_rn(sx ? a.x : a.w, q.ix, q.nox);
            const float tny = __fmaf_rn(sy ? b.x : a.y, q.iy, q.noy), tfy = __fmaf_rn(sy ? a.y : b.x, q.iy, q.noy);
            const float tnz = __fmaf_rn(sz ? b.y : a.z, q.iz, q.noz), tfz = __fmaf_rn(sz ? a.z : b.y, q.iz, q.noz);
            const float tmin = fmaxf(fmaxf(fmaxf(tnx, tny), tnz), 0.f), tmax = fminf(fminf(fminf(tfx, tfy), tfz), tLimit);
            key = (tmin <= tmax && ref != GW_DONE) ? ((__float_as_int(tmin) & ~(GW_LANES - 1)) | j) : KEY_MISS;
        }
        const bool inner = ref >= 0;
        const int kI = (key != KEY_MISS && inner) ? key : KEY_MISS;
        int kmin = min(kI, __shfl_xor_sync(FULL_MASK, kI, 1));
        kmin = min(kmin, __shfl_xor_sync(FULL_MASK, kmin, 2));
#if GW_LANES == 8
        kmin = min(kmin, __shfl_xor_sync(FULL_MASK, kmin, 4));
#endif
        const int nextRef = __shfl_sync(FULL_MASK, ref, (lane & ~(GW_LANES - 1)) | (kmin & (GW_LANES - 1)));
        const bool pushMe = kI != KEY_MISS && kI != kmin; // keys are distinct (child number in the low bits)
        const bool leafHit = key != KEY_MISS && !inner;
        const unsigned int pm = __ballot_sync(FULL_MASK, pushMe), lm = __ballot_sync(FULL_MASK, leafHit);
        if (isNode)
        {
            if (sp > spLimit)
            {
                atomicOr(&s_ovf[warp], 1u << slot); // degenerate tree: the ordered walk takes this ray
                cur = GW_DONE;
            }
            else
            {
                const unsigned int gpm = (pm >> (g * GW_LANES)) & ((1u << GW_LANES) - 1u);
                if (pushMe) stackPut(sp + __popc(gpm & ((1u << j) - 1u)) * GW_STACK_STRIDE, ref, key);
                sp += __popc(gpm) * GW_STACK_STRIDE;
                cur = (kmin == KEY_MISS) ? GW_POP : nextRef;
            }
        }
        if (leafHit) s_queue[warp][qn + __popc(lm & ltMask)] = make_int2(ref, slot);
        qn += __popc(lm);

        // ---- primitive tests: one queue entry per lane, whichever ray it belongs to
        const int nAct = __popc(am & GW_LEADERS);
        if (qn > 0 && qn >= min(GW_QT, GW_QPER * nAct))
        {
            const int nProc = min(qn, 32);
            qn -= nProc;
            gwLeafBatch(qn, nProc, scratch);
        }
    }
    __syncwarp();
    if (need)
    {
        const unsigned long long res = s_res[warp][lane];
        const unsigned int low = (unsigned int)res;
        if ((s_ovf[warp] >> lane) & 1u) { out.hit.prim = -2; out.shadow = -1.f; } // caller runs the ordered walk
        else if (mode == UW_SHADOW) out.shadow = (low != 0u) ? fmaxf(0.f, fminf(shadowLimit, shadowLimit)) : 0.f;
        else if (mode == UW_CLOSEST)
        {
            if (low != GW_NOPRIM)
            {
                const float4 h = s_hit[warp][lane];
                out.hit.prim = (int)low; out.hit.p = f3(h.x, h.y, h.z); out.hit.flags = __float_as_int(h.w);
            }
        }
        else
        {
            // replay in array order (selection by ascending index: the list is short; a cylinder listed twice by the point query
            // is taken once).  Primitives of one leaf are contiguous and share the leaf's fate, decided when the leaf is reached
            // (before any of its primitives): t_min(leaf) < closest-so-far.
            const int n = (int)low;
            const float window = fminf(minDistance0, GATHER_WINDOW * __uint_as_float((unsigned int)(res >> 32)));
            const float4* const list = scratch + (size_t)lane * GW_GATHER_CAP;
            float m = minDistance0;
            bool leafPass = false;
            int prevLeaf = -1, winner = -1, last = -1;
            for (int pass = 0; pass < n; ++pass)
            {
                int bi = 0x7fffffff;
                float4 bc = f4(0.f, 0.f, 0.f, 0.f);
                for (int k = 0; k < n; ++k)
                {
                    const float4 c = __ldcg(list + k);
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
            if (winner >= 0)
            {
                Ray r;
                makeRay(r, rayOrigin, rayDir);
                const float4* item = recs + (size_t)PRIM_REC_F4 * winner;
                float4 a0, a1, a2, a3;
                ldNode256(item, a0, a1); ldNode256(item + 2, a2, a3);
                float3 I;
                int flags;
                float planeShadow;
                primitiveTestRegs(a0, a1, a2, a3, winner, __float_as_int(a3.w), r, I, flags, planeShadow); // same hit point as when it was gathered
                out.hit.prim = winner; out.hit.p = I; out.hit.flags = flags;
            }
        }
    }
    __syncwarp(); // the next call reuses the shared arrays
    return out;
}

// closest hit of a ray class whose result does not depend on the visiting order; every lane of the warp calls
SB_DEV Hit closestHitGroup(const float3 origin, const float3 target, const int iteration, const int currentMaterialId, const bool need)
{
    const float3 d = target - origin;
    const int mode = (dot(d, d) >= 1.0002f) ? UW_CLOSEST : UW_GATHER;
    Hit hit = groupWalk(need, mode, origin, d, iteration, currentMaterialId, 0, 0).hit;
    if (need && hit.prim == -2) hit = closestHitWide(origin, target, iteration, currentMaterialId);
    return hit;
}
