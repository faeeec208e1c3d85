"""Analysis, not a test (CPU only): how much of a warp's walk time is the longest of its 32 walks, and what grouping the rays
differently would change.

The oracle logs every closest-hit ray of a frame of config 2 (origin, target, hit distance); oracle_walk_visits counts, per ray,
the nodes an ideal front-to-back walk visits in the engine's own 4-wide tree (sol-r_b200/csrc/engine.cu, built host-only).  A warp
runs as long as its longest walk, so with rays grouped 32 at a time   utilisation = sum(visits) / (32 * sum over warps of max(visits)).
Groupings compared per pass: (a) tile order — what the staged kernels' compacted queues hold (paths in the order of the 8x4-pixel
tile they started from); (b) the same rays sorted by direction octant and the Z-order cell of their origin; (c) sorted by the
visit count itself (the unreachable optimum, for scale).

Bounce rays (passes >= 1) are walked by the engine in "gather" mode, which keeps every candidate within 1.5 x the closest distance
(DESIGN.md 3): their walks are modelled with the bound at 1.5 x the hit distance.

usage: python tests/analysis_ray_order.py [width height]     (default 960 540)
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402

from _solr_b200_import import solr_b200  # noqa: E402,F401
import oracle  # noqa: E402
from solr_b200 import engine, host, scenes, wire  # noqa: E402


def spread(v):
    v = v.astype(np.uint64)
    v = (v | (v << np.uint64(16))) & np.uint64(0x030000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x0300F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x030C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x09249249)
    return v


def utilisation(visits):
    n = len(visits) // 32 * 32
    if n == 0:
        return float("nan"), 0
    v = visits[:n].reshape(-1, 32)
    return float(v.sum()) / float(32 * v.max(axis=1).sum()), int(v.max(axis=1).sum())


def main():
    W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (960, 540)
    sc = scenes.config2()
    si = wire.default_scene_info(W, H, graphics_level=wire.GL_FULL, nb_ray_iterations=3)
    h = host.SceneHost(si)
    sc.replay(h)
    a = h.arrays()
    bounds = a["bounds"]
    h.close()
    lib = oracle.load()
    lib.oracle_ray_log.argtypes = [C.c_void_p, C.c_ulonglong]
    lib.oracle_ray_log.restype = C.c_ulonglong
    lib.oracle_walk_visits.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_ulonglong, C.c_void_p]
    cap = W * H * 4
    log = np.zeros((cap, 9), np.float32)
    lib.oracle_ray_log(log.ctypes.data_as(C.c_void_p), cap)
    o = oracle.Oracle(a, W, H, randoms=np.zeros(wire.REF_MAX_BITMAP_SIZE, np.float32))
    o.render(si, sc.eye, sc.target, sc.angles, threads=os.cpu_count())
    n = int(lib.oracle_ray_log(None, 0))
    rays = log[:min(n, cap)].copy()
    bounce = (rays[:, 1] >= 1) & (rays[:, 8] > 0)
    rays[bounce, 8] *= 1.5   # the gather window of the bounce-ray walks
    nodes, _, nb_main, _ = engine.build_walk_trees(a)
    visits = np.zeros(len(rays), np.uint32)
    flat = np.ascontiguousarray(nodes.reshape(-1))
    CUT = 24
    stack_at, max_stack = np.zeros(len(rays), np.uint32), np.zeros(len(rays), np.uint32)
    lib.oracle_walk_visits_ex.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.oracle_walk_visits_ex(flat.ctypes.data_as(C.c_void_p), nb_main, rays.ctypes.data_as(C.c_void_p), len(rays), visits.ctypes.data_as(C.c_void_p),
                              CUT, stack_at.ctypes.data_as(C.c_void_p), max_stack.ctypes.data_as(C.c_void_p))
    print("config 2 at %dx%d: %d closest-hit rays, %d nodes in the tree" % (W, H, len(rays), nb_main))
    pixel = rays[:, 0].astype(np.int64)
    tile = (pixel // W // 4) * ((W + 7) // 8) + (pixel % W) // 8
    for it in sorted(set(rays[:, 1].astype(int))):
        m = rays[:, 1].astype(int) == it
        r, v, t, px = rays[m], visits[m].astype(np.int64), tile[m], pixel[m]
        order_tile = np.lexsort((px, t))
        d = r[:, 5:8] - r[:, 2:5]
        octant = (d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << np.uint64(1)) | ((d[:, 2] < 0).astype(np.uint64) << np.uint64(2))
        lo, hi = bounds[:3], bounds[3:]
        cell = np.clip(((r[:, 2:5] - lo) / np.maximum(hi - lo, 1e-6) * 64).astype(np.int64), 0, 63)
        morton = spread(cell[:, 0]) | (spread(cell[:, 1]) << np.uint64(1)) | (spread(cell[:, 2]) << np.uint64(2))
        order_sorted = np.argsort((octant << np.uint64(32)) | morton, kind="stable")
        order_cell_only = np.argsort(morton, kind="stable")
        order_best = np.argsort(v, kind="stable")
        hit = r[:, 8] > 0
        print("pass %d: %7d rays, %4.1f %% hit, visits mean %.1f  max %d" % (it, len(v), 100.0 * hit.mean(), v.mean(), v.max()))
        base = None
        for name, order in (("tile order (as queued)", order_tile), ("origin cell (Z-order)", order_cell_only),
                            ("octant + origin cell", order_sorted), ("by visit count (optimum)", order_best)):
            u, cost = utilisation(v[order])
            base = base or cost
            print("    %-26s utilisation %5.1f %%   warp-rounds %9d  (%.2f x)" % (name, 100.0 * u, cost, cost / base))
        print("    visits: hits mean %.1f, misses mean %.1f; share of rays with more than 32 / 48 / 64 visits: %.1f / %.1f / %.1f %%" % (
            v[hit].mean() if hit.any() else 0.0, v[~hit].mean() if (~hit).any() else 0.0, 100.0 * (v > 32).mean(), 100.0 * (v > 48).mean(), 100.0 * (v > 64).mean()))
        sa, ms = stack_at[m][v > CUT], max_stack[m]
        if len(sa):
            print("    stack entries a walk cut after %d visits would park: mean %.1f, 95 %% <= %d, max %d (deepest stack of any walk: %d)" % (
                CUT, sa.mean(), int(np.percentile(sa, 95)), sa.max(), ms.max()))
        # (d) walks cut off after R node rounds: the unfinished rays are parked with their stacks, compacted and continued 32 at a time
        vt = v[order_tile]
        for R in (16, 24, 32, 48):
            rest, cost, phases = vt.copy(), 0, 0
            while len(rest):
                nfull = (len(rest) + 31) // 32 * 32
                pad = np.concatenate([rest, np.zeros(nfull - len(rest), rest.dtype)]).reshape(-1, 32)
                cost += int(np.minimum(pad.max(axis=1), R).sum())
                rest = rest[rest > R] - R
                phases += 1
            print("    cut after %2d rounds, continue compacted: warp-rounds %9d  (%.2f x) in %d phases" % (R, cost, cost / base, phases))
        # (f) a pool of R rays per warp, lanes fetch the next ray when they finish theirs (no state parked, no second launch)
        for R in (32, 64, 128, 256):
            c1 = pool_rounds(vt, R)
            c4 = pool_rounds(vt, R, refill_at=4)
            print("    pool of %3d rays per warp, per-lane refill: warp-rounds %9d  (%.2f x); refill when 4 lanes idle: %9d  (%.2f x)" % (
                R, c1, c1 / base, c4, c4 / base))
        # (e) no state parked: a walk that exceeds its budget is abandoned and the ray walks again from the root in a later group of
        # long rays (budgets R1 < R2 < unlimited)
        def grouped(x, cap):
            nfull = (len(x) + 31) // 32 * 32
            pad = np.concatenate([x, np.zeros(nfull - len(x), x.dtype)]).reshape(-1, 32)
            return int(np.minimum(pad.max(axis=1), cap).sum()) if len(x) else 0
        for budgets in ((24,), (32,), (40,), (24, 48), (32, 64)):
            rest, cost = vt, 0
            for R in budgets:
                cost += grouped(rest, R)
                rest = rest[rest > R]
            cost += grouped(rest, 1 << 30)
            print("    abandon after %-8s rounds, walk again compacted: warp-rounds %9d  (%.2f x)" % ("/".join(map(str, budgets)), cost, cost / base))


def pool_rounds(v, R, lanes=32, refill_at=1):
    """Warp-rounds when a warp takes R rays at a time and each of its lanes fetches the next ray of the pool as soon as it has
    finished its own (greedy list scheduling in queue order); the warp runs until the pool's last ray is done.
    refill_at: idle lanes wait until this many are idle (the refill branch then serves them together)."""
    import heapq
    total = 0
    for b in range(0, len(v), R):
        pool = list(v[b:b + R])
        if refill_at <= 1:
            heap = [0] * lanes
            for x in pool:
                t = heapq.heappop(heap)
                heapq.heappush(heap, t + int(x))
            total += max(heap)
        else:
            # event simulation with grouped refills
            busy_until = [0] * lanes
            nxt, now = 0, 0
            while True:
                idle = [i for i in range(lanes) if busy_until[i] <= now]
                if nxt < len(pool) and (len(idle) >= refill_at or len(idle) == lanes):
                    for i in idle:
                        if nxt < len(pool):
                            busy_until[i] = now + int(pool[nxt]); nxt += 1
                later = [t for t in busy_until if t > now]
                if not later:
                    if nxt >= len(pool):
                        break
                    continue
                now = min(later)
            total += now
    return total


if __name__ == "__main__":
    main()
