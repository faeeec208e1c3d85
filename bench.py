#!/usr/bin/env python
"""bench.py — Mrays/s of Sol-R's ray-propagation hot path on B200 (contract: see the task's bench section).

A *step* is one full frame of the workload through the hot path: per-pixel ray generation, box-list walk,
primitive tests, shading with shadow rays, reflection/refraction bounce loop, accumulation, RGB8 pack.
A *ray* is one box-list walk (closest-hit or shadow), SURVEY.md §8(d), counted in-kernel.

  workload  the headline line is BASELINE.json configs[1] (sol-r_b200/workloads.py "config2"): the molecule scene (105 k atom
          spheres + bond cylinders + ground + light = 216 k primitives), 1920x1080, 1 spp, shadows, 3 bounces — synthetic
          (the reference's 486-atom PDB file does not travel; sol-r_b200/scenes.py generates the lattice).  The other
          configurations ride along in the same JSON line under "workloads" (fewer frames each): config4 (1 M spheres,
          3840x2160, accumulated samples — the size north_star's 8-GPU efficiency target names) at every N; config3
          (1 M triangles, 5 bounces) and config5 (anaglyph 4K, 16-frame progressive) at N = 1.  --workload X makes X the headline.

  value   whole-job Mrays/s with scene and frame state resident in HBM: K x b200_render, CUDA events on the
          render stream around each frame, L2 flushed (256 MiB memset) between frames, max over ranks.
          N>1: ONE frame is split across the N GPUs as interleaved 8x4-pixel tiles (strong scaling); the exchange is fused
          into the ray kernels — every rank stores its finished pixels straight into rank 0's device bitmap through NVLink
          peer memory (partition.PeerFrame) — and a one-element NCCL all-reduce inside the timed region orders the frame.
          After the timed runs rank 0 renders the whole frame alone and checks the merged frame against it, byte for byte.
  e2e     the same metric through the host-side drop-in (SceneHost.render_begin + render_end = the calls CudaKernel makes) with
          the REFERENCE'S protocol: per-frame parameter upload, kernels, device->host copy of RGB8 AND of the 16-byte-per-pixel
          id buffer into caller-owned host memory (CudaKernel.cpp:304-313), wall clock.  The figure with the id buffer fetched on
          demand (getPrimitiveAt is its only reader) is reported beside it as e2e.lazy_ids.
  roofline  FP32 CUDA-core roofline (north_star: compute-bound, no tensor cores): algorithmic flops counted by
          the oracle in REFERENCE traversal order (SURVEY §8(d) constants) / device time / peak, peak = 148 SM x 128 lanes x
          2 x sm_max_mhz; the FFMA rate and clock a microbenchmark sustains on this GPU are reported beside it
          (b200_measure_fp32_peak).  DRAM traffic per frame comes from the committed ncu capture, stamped with its commit.
  parity  (N = 1, headline) after the timed region the frame is rendered once more and compared with the reference's own CUDA
          engine on the same GPU (oracle/_ref/libsolr_ref_cuda.so, when it travelled): differing hit ids, pixels beyond 2/255.
  cpu_baseline / --impl reference: the reference's own code on the host cores (oracle/_ref/libsolr_ref_cpu.so = its CUDA source
          compiled for the host with OpenMP over blocks — NOT its OpenCL engine: no OpenCL CPU device exists on the box,
          profiles/r02_opencl_probe.txt) when that library travelled, else the oracle port; sample = one whole frame of the same
          scene / camera per step.  The reference arm loads no library of the product.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

SM_COUNT, LANES_PER_SM = 148, 128
# per-workload frames when a workload rides along as a sub-record: (warm-up frames, timed frames)
SUB_FRAMES = {"config2": (3, 12), "config3": (3, 8), "config4": (4, 12), "config5": (16, 32)}


def peaks():
    p = {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update(hbm_gbs=float(m["hbm_gbs"]), sm_max_mhz=float(m.get("sm_max_mhz", 1965.0)), source="measured")
    except Exception:
        pass
    p["fp32_tflops"] = SM_COUNT * LANES_PER_SM * 2 * p["sm_max_mhz"] * 1e6 / 1e12
    return p


def recorded_flops():
    try:
        with open(os.path.join(ROOT, "profiles", "algorithmic_flops.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


_JSON_OUT = None


def claim_stdout():
    """stdout carries the one JSON line and nothing else: native libraries in the process log there too (the reference library's
    banner, NCCL's version line), so file descriptor 1 is pointed at stderr and the line is written to the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def host_cores():
    """Host threads for the CPU legs: every core this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its
    workers, which would silently run the reference arm on one core at N > 1, so the OpenMP runtime is told explicitly."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1", mode=ctypes.RTLD_GLOBAL).omp_set_num_threads(n)
    except Exception:
        pass
    return n


# ------------------------------------------------------------------------------------------------------------------
# CPU legs: the reference's own code (or the oracle port) on the host cores.  No native library of the product is loaded here when
# the reference library travelled: the reference's own container builds the scene.  Without it the product's host container
# builds the same arrays for the port, and the record says so.
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(key, steps, warmup):
    import oracle
    import refh
    from solr_b200 import wire, workloads   # pure-Python struct mirrors and scene recipes: no native library of the product
    wl = workloads.WORKLOADS[key]
    W, H = wl["size"]
    if (W, H) != (1920, 1080):
        raise SystemExit("bench.py: the CPU legs exist for the 1920x1080 workloads (the reference's frame limit)")
    si = workloads.scene_info(key)
    sc = wl["scene"]()
    rnd = np.zeros(wire.REF_MAX_BITMAP_SIZE, np.float32)
    cores = host_cores()
    builder = "reference container (libsolr_ref_cpu.so)"
    r = None
    if refh.available("cpu"):
        r = refh.RefScene(si, "cpu")
        sc.replay(r)
        a = r.arrays()
    else:
        from solr_b200 import host   # the reference library did not travel: the product's container builds the same arrays
        builder = "product host container (reference library absent)"
        h = host.SceneHost(si)
        sc.replay(h)
        a = h.arrays()
        h.close()
    o = oracle.Oracle(a, W, H, randoms=rnd)
    t0 = time.perf_counter()
    o.render(si, sc.eye, sc.target, sc.angles, threads=cores)
    t_port = time.perf_counter() - t0
    counters = o.counters.as_dict()
    rays = counters["rays"]
    kind, times = "port", []
    if r is not None:
        kind = "reference"
        for k in range(warmup + steps):
            t0 = time.perf_counter()
            r.render(si, sc.eye, sc.target, sc.angles, randoms=rnd, block=(16, 8), want_post=False)
            if k >= warmup:
                times.append(time.perf_counter() - t0)
        r.close()
    else:
        for k in range(warmup + steps):
            o2 = oracle.Oracle(a, W, H, randoms=rnd)
            t0 = time.perf_counter()
            o2.render(si, sc.eye, sc.target, sc.angles, threads=cores)
            if k >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"kind": kind, "cores": cores, "rays_per_frame": rays, "sec_per_frame": sec, "mrays_s": rays / sec / 1e6,
            "port_sec_per_frame": t_port, "flops_per_frame": o.flops(), "counters": counters, "scene_built_by": builder,
            "sample": "whole %dx%d frame of the same scene and camera per step, %d timed frames; %s" % (
                W, H, len(times), "the reference's CUDA source compiled for the host (OpenMP), not its OpenCL engine" if kind == "reference"
                else "oracle port")}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from solr_b200 import workloads
    r = cpu_reference_run(args.workload, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "Mrays/s", "value": r["mrays_s"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["sec_per_frame"] * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workloads.WORKLOADS[args.workload]["name"], "bench_sample": r["sample"], "scene_built_by": r["scene_built_by"]},
            "cpu_baseline": {"value": r["mrays_s"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["mrays_s"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------------------------
# engine arm
# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def measure_workload(key, steps, warmup, ctx, sample_clocks=False, want_e2e=True):
    """One workload through the drop-in host path and the C ABI on this process's GPU: device-resident timing, end-to-end timing,
    and (N > 1) the merged-frame check.  Returns a dict (complete on rank 0)."""
    import torch
    import torch.distributed as dist
    from solr_b200 import engine, host, partition, wire, workloads
    wl = workloads.WORKLOADS[key]
    W, H = wl["size"]
    world, rank, local_rank, stream, lib = ctx.world, ctx.rank, ctx.local_rank, ctx.stream, ctx.lib
    sc = wl["scene"]()
    si = workloads.scene_info(key)
    limits = wl["limits"]
    table = max(wire.REF_MAX_BITMAP_SIZE, W * H)
    rnd = np.zeros(table, np.float32)
    iterations = wl["iterations"]
    frame_no = [-1]

    def next_iteration():
        frame_no[0] += 1
        return iterations[frame_no[0] % len(iterations)]

    # ---- the drop-in host path (SceneHost -> C ABI) owns the engine in this process --------------------
    t0 = time.perf_counter()
    h = host.SceneHost(si, limits=limits, rank=rank, world=world, device=local_rank, capacity=wl["capacity"])
    sc.replay(h)
    host_build_s = time.perf_counter() - t0
    h.set_randoms(rnd, 0)
    h.set_camera(sc.eye, sc.target, sc.angles)
    h.init_buffers()
    lib.b200_set_stream(stream.cuda_stream)
    si_live = h.scene_info
    si_live.maxPathTracingIterations = 1 << 30   # keep m_refresh true: every render_begin renders a frame

    def frame_e2e_once(iteration=0):
        si_live.pathTracingIteration = iteration
        h.set_scene_info(si_live)
        h.render_begin(0.0)
        h.render_end()

    t0 = time.perf_counter()
    frame_e2e_once()   # uploads the scene (dirty flags), first frame
    torch.cuda.synchronize()
    first_frame_s = time.perf_counter() - t0
    eng = engine.Engine.__new__(engine.Engine)   # thin view on the already-initialised library for counters etc.
    eng.lib = lib
    eng.counters(reset=True)
    a = h.arrays()
    objects = wire.Int4(a["nbBoxes"], a["nbPrimitives"], a["nbLamps"], a["lightInformationSize"])
    occ = wire.Int2(1, 1)
    eye, target, angles = wire.Float3(*sc.eye), wire.Float3(*sc.target), wire.Float4(*sc.angles)
    pp = wire.PostProcessingInfo()
    si0 = workloads.scene_info(key)
    si0.maxPathTracingIterations = 1 << 30
    peer = partition.PeerFrame(lib, rank, world, stream=stream) if world > 1 else None   # ranks > 0 now write into rank 0's frame

    def frame_device():
        it = next_iteration()
        si0.pathTracingIteration = it
        lib.b200_render(occ, wire.Int4(8, 4, 1, 0), si0, objects, pp, eye, target, angles)
        if world > 1:
            peer.fence()   # every rank's kernels, and with them their stores into rank 0's frame, are done
        return it

    # ---- device-resident timing --------------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if sample_clocks else None
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            ctx.flush.zero_()
            frame_device()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.counters(reset=True)
        launches0 = eng.kernel_launches()
        if sampler:
            sampler.start()
        evs = []
        for _ in range(steps):
            ctx.flush.zero_()                                  # L2 flush, outside the timed region
            if world > 1:
                dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            it = frame_device()
            e.record(stream)
            evs.append((s, e, it))
        stream.synchronize()
        torch.cuda.synchronize()
    clocks = sampler.summary() if sampler else None
    ms = [s.elapsed_time(e) for s, e, _ in evs]
    per_it = {}
    for (s, e, it), m in zip(evs, ms):
        per_it.setdefault(it, []).append(m)
    launches = eng.kernel_launches() - launches0
    rays_local, px_local = eng.counters(reset=True)
    t = torch.tensor([sum(ms)] + [sum(per_it[i]) / len(per_it[i]) for i in sorted(per_it)], dtype=torch.float64, device="cuda")
    r = torch.tensor([float(rays_local), float(px_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    total_ms = float(t[0].item())
    rays_total, px_total = float(r[0].item()), float(r[1].item())
    rec = {"workload": wl["name"], "ms_per_step": total_ms / steps, "value": rays_total / (total_ms * 1e-3) / 1e6, "steps": steps, "warmup": warmup,
           "rays_per_frame": rays_total / steps, "mpixels_per_s": px_total / (total_ms * 1e-3) / 1e6,
           "ms_per_iteration": {str(i): round(float(v), 4) for i, v in zip(sorted(per_it), t[1:].tolist())},
           "gpu_launches": int(launches), "kernel_ms_last_frame": float(lib.b200_last_render_ms()), "clocks": clocks,
           "primitives": int(a["nbPrimitives"]), "boxes": int(a["nbBoxes"]), "host_scene_build_s": round(host_build_s, 2),
           "first_frame_with_scene_upload_s": round(first_frame_s, 2), "size": [W, H]}

    # ---- end to end through the host drop-in -------------------------------------------------------------
    # Every rank runs the host drop-in for its tiles (render_begin: per-frame parameter upload + launch); finished pixels land in
    # rank 0's frame over NVLink (the path's one exchange step) and rank 0 reads the merged frame back to host memory
    # (render_end = d2h_bitmap).  Wall clock around K frames, max over ranks.
    meet = [peer]   # what orders the ranks around a frame: the peer frame's fence, then the shared host frame's

    def frame_e2e_all(iteration=None):
        si_live.pathTracingIteration = next_iteration() if iteration is None else iteration
        h.set_scene_info(si_live)
        if world > 1:
            with torch.cuda.stream(stream):
                meet[0].fence()   # rank 0 has read the previous frame
        h.render_begin(0.0)
        if world > 1:
            with torch.cuda.stream(stream):
                meet[0].fence()   # the frame is complete: in rank 0's device bitmap, or in the shared host frame
        if rank == 0:
            h.render_end()
        else:
            stream.synchronize()

    def measure_e2e():
        frame_no[0] = -1
        for _ in range(max(3, len(iterations))):
            frame_e2e_all()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        eng.counters(reset=True)
        t0 = time.perf_counter()
        for _ in range(steps):
            frame_e2e_all()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rays_e2e, _ = eng.counters(reset=True)
        te = torch.tensor([dt], dtype=torch.float64, device="cuda")
        re_ = torch.tensor([float(rays_e2e)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(re_, op=dist.ReduceOp.SUM)
        return float(re_.item()) / float(te.item()) / 1e6, float(te.item()) / steps * 1e3

    if want_e2e:
        # the reference's own protocol first — pixels and 16 bytes of ids per pixel, every frame (CudaKernel.cpp:304-313) —, then the
        # drop-in's option: the id buffer stays on the device until getPrimitiveAt asks (its only host-side reader, GPUKernel.cpp:729-739)
        h.set_lazy_ids(False)
        if world > 1:
            # first the frame merged in the root GPU's memory over NVLink and read back by the root (its own ids only) ...
            lib.b200_set_option(12, 0)
            copied_value, copied_ms = measure_e2e()
            h.set_lazy_ids(True)
            lazy_value, lazy_ms = measure_e2e()
            h.set_lazy_ids(False)
            lib.b200_set_option(12, 1)
            # ... then ONE frame in shared host memory that every GPU's kernels fill with their own tiles, ids included, over their
            # own PCIe links (partition.SharedHostFrame): no exchange between GPUs, no read-back on the root
            meet[0] = partition.SharedHostFrame(h, lib, rank, world, stream=stream)
        streamed0 = int(lib.b200_frames_streamed())
        eager_value, eager_ms = measure_e2e()
        streamed = int(lib.b200_frames_streamed()) - streamed0
        if world == 1:
            lib.b200_set_option(12, 0)   # the same protocol with the frame and the ids copied after the kernels (no streamed output)
            copied_value, copied_ms = measure_e2e()
            lib.b200_set_option(12, 1)
            h.set_lazy_ids(True)
            lazy_value, lazy_ms = measure_e2e()
        rec["e2e"] = {"value": eager_value, "unit": "Mrays/s", "ms_per_frame": eager_ms,
                      "output": (("streamed: every GPU's ray kernels write the tiles they own, pixels and ids, into one frame in shared pinned host memory (%d of %d timed + warm-up frames on rank 0)"
                                  if world > 1 else
                                  "streamed: the ray kernels write the caller's pinned frame and id buffers tile by tile as tiles finish (%d of %d timed + warm-up frames)")
                                 % (streamed, steps + max(3, len(iterations)))) if streamed else "copied after the kernels",
                      "copied_output": {"value": copied_value, "ms_per_frame": copied_ms,
                                        "note": "option 12 = 0: cudaMemcpyAsync of both buffers in render_end" +
                                                (" on rank 0, from the frame the other GPUs' kernels filled over NVLink (the id buffer: rank 0's own tiles only)" if world > 1 else "")},
                      "h2d_bytes_per_step": int(lib.b200_frame_parameter_bytes()) * world,   # scene-info + camera + pointers block, per frame and rank
                      "d2h_bytes_per_step": W * H * 3 + W * H * 16,   # RGB8 + PrimitiveXYIdBuffer into caller-owned memory (rank 0)
                      "protocol": "the reference's render_end: bitmap and id buffer read back every frame (CudaKernel.cpp:304-313)",
                      "lazy_ids": {"value": lazy_value, "ms_per_frame": lazy_ms, "d2h_bytes_per_step": W * H * 3,
                                   "note": "id buffer fetched on demand (getPrimitiveAt is its only host-side reader)"}}
    frame_e2e_all(0)   # a single-sample frame on every rank: what frame_check compares

    if world > 1:
        # the merged frame rank 0 just read back against the whole frame rendered by rank 0 alone
        dist.barrier()
        peer.close()
        if rank == 0:
            merged, merged_ids = np.array(h.bitmap(), copy=True), np.array(h.primitive_ids(), copy=True)
            lib.b200_set_partition(0, 1)
            frame_e2e_once(0)
            whole, whole_ids = np.array(h.bitmap(), copy=True), np.array(h.primitive_ids(), copy=True)
            differing = int(np.count_nonzero(merged != whole))
            rec["frame_check"] = {"bytes_differing_from_the_one_gpu_frame": differing, "bytes": int(whole.size)}
            if want_e2e:   # the shared host frame carries every rank's ids too
                rec["frame_check"]["id_words_differing_from_the_one_gpu_frame"] = int(np.count_nonzero(merged_ids != whole_ids))
                differing += rec["frame_check"]["id_words_differing_from_the_one_gpu_frame"]
            if differing:
                raise SystemExit("bench.py: %s: the frame merged over %d GPUs differs from the one-GPU frame in %d bytes" % (key, world, differing))
        dist.barrier()
    rec["_frame"] = (np.array(h.bitmap(), copy=True), np.array(h.primitive_ids(), copy=True)) if (rank == 0 and world == 1) else None
    h.close()
    return rec



def measure_sample_split(key, ctx, groups=3):
    """north_star's second split (configs 4 / 5: "sample accumulation optionally split by GPU", reduced with NCCL over NVLink): every
    rank renders the WHOLE frame for its share of the accumulation iterations (partition.SampleSplit), the partial colour sums are
    sum-reduced onto the root and packed.  Past NB_MAX_ITERATIONS a frame only adds a sample (CudaRayTracer.cu:550-562), so the
    samples of a progressive sequence are exchangeable.  Timed: `groups` rounds of one accumulation iteration per rank + the reduce
    and the pack, wall clock between device synchronisations, max over ranks.  Complete on rank 0."""
    import torch
    import torch.distributed as dist
    from solr_b200 import engine, host, partition, wire, workloads
    wl = workloads.WORKLOADS[key]
    W, H = wl["size"]
    world, rank, local_rank, stream, lib = ctx.world, ctx.rank, ctx.local_rank, ctx.stream, ctx.lib
    sc = wl["scene"]()
    si = workloads.scene_info(key)
    si.maxPathTracingIterations = 1 << 30
    h = host.SceneHost(si, limits=wl["limits"], capacity=wl["capacity"])
    sc.replay(h)
    a = h.arrays()
    h.close()
    e = engine.Engine(si, device=local_rank, limits=wl["limits"])      # whole frame on every rank
    e.upload(a, randoms=np.zeros(max(wire.REF_MAX_BITMAP_SIZE, W * H), np.float32))
    first = 11   # NB_MAX_ITERATIONS + 1: the first iteration that only adds a sample
    for it in range(0, first):                                         # the deepening passes and the first sample: everybody
        si.pathTracingIteration = it
        e.render(si, sc.eye, sc.target, sc.angles)
    e.synchronize()
    split = partition.SampleSplit(lib, rank, world, W, H)
    split.begin()
    e.counters(reset=True)
    last = first + groups * world - 1
    mine = split.iterations(first, last)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for it in mine:
        si.pathTracingIteration = it
        e.render(si, sc.eye, sc.target, sc.angles)
    e.synchronize()
    t1 = time.perf_counter()
    split.finish(last)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    rays, _ = e.counters(reset=True)
    t = torch.tensor([t2 - t0, t1 - t0], dtype=torch.float64, device="cuda")
    r = torch.tensor([float(rays)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(r, op=dist.ReduceOp.SUM)
    e.close()
    dt, render_s = float(t[0].item()), float(t[1].item())
    samples = groups * world
    return {"split": "samples: every GPU renders the whole frame for iterations = rank mod N, colour sums reduced onto rank 0 over NCCL and packed",
            "samples": samples, "ms_per_sample": dt / samples * 1e3, "value": float(r.item()) / dt / 1e6, "unit": "Mrays/s",
            "ms_render_per_rank": render_s * 1e3, "ms_reduce_and_pack": (dt - render_s) * 1e3,
            "reduce_bytes_per_rank": W * H * 16, "iterations": [first, last]}


def measure_scene_paths(ctx):
    """The scene side of the path, on config 2, outside every timed region above:
      N = 1  one animation step of the drop-in host path — rotatePrimitives + compactBoxes(false) (MoleculeScene.cpp:75-81), then a
             frame, which re-uploads the scene — with the walk trees built on host threads and on the GPU (option key 10);
      N > 1  scene replication — every rank uploading the scene itself against the root uploading it and the other ranks receiving
             the device arrays over NCCL (partition.broadcast_scene), max over ranks."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from solr_b200 import engine, host, partition, wire, workloads
    key = "config2"
    wl = workloads.WORKLOADS[key]
    W, H = wl["size"]
    world, rank, lib = ctx.world, ctx.rank, ctx.lib
    sc = wl["scene"]()
    si = workloads.scene_info(key)
    rnd = np.zeros(max(wire.REF_MAX_BITMAP_SIZE, W * H), np.float32)

    def upload_stats():
        ms, n0, n1, gpu = C.c_float(), C.c_int(), C.c_int(), C.c_int()
        lib.b200_scene_upload_stats(C.byref(ms), C.byref(n0), C.byref(n1), C.byref(gpu))
        return ms.value, n0.value + n1.value, gpu.value

    if world == 1:
        out = {"workload": wl["name"], "step": "rotatePrimitives + compactBoxes(false) + one frame (scene re-uploaded), 3 steps each"}
        for name, opt in (("host_built_trees", 0), ("gpu_built_trees", 1)):
            lib.b200_set_option(10, opt)
            h = host.SceneHost(si, limits=wl["limits"], rank=0, world=1, device=ctx.local_rank, capacity=wl["capacity"])
            sc.replay(h)
            h.set_randoms(rnd, 0)
            h.set_camera(sc.eye, sc.target, sc.angles)
            h.init_buffers()
            h.render_begin(0.0); h.render_end()
            torch.cuda.synchronize()
            flat, frame, upl, render = [], [], [], []
            for k in range(3):
                t0 = time.perf_counter()
                h.rotate_primitives((0.0, 0.0, 0.0), (0.0, 0.05, 0.0))
                h.compact_boxes(False)
                t1 = time.perf_counter()
                h.render_begin(0.0); h.render_end()
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                flat.append((t1 - t0) * 1e3); frame.append((t2 - t1) * 1e3)
                upl.append(upload_stats()[0]); render.append(float(lib.b200_last_render_ms()))
            out[name] = {"host_box_compaction_ms": round(min(flat), 1), "frame_with_scene_upload_ms": round(min(frame), 1),
                         "of_which_b200_h2d_scene_ms": round(min(upl), 1), "of_which_render_kernels_ms": round(min(render), 2),
                         "walk_tree_nodes": upload_stats()[1]}
            h.close()
        lib.b200_set_option(10, 0)
        # the same step applied where the scene lives (b200_rotate_primitives, csrc/animate.cuh): no host compaction, no upload
        h = host.SceneHost(si, limits=wl["limits"], rank=0, world=1, device=ctx.local_rank, capacity=wl["capacity"])
        sc.replay(h)
        h.set_randoms(rnd, 0)
        h.set_camera(sc.eye, sc.target, sc.angles)
        h.init_buffers()
        h.render_begin(0.0); h.render_end()
        torch.cuda.synchronize()
        h.set_device_animation(True)
        step, frame, render = [], [], []
        for k in range(5):
            t0 = time.perf_counter()
            h.rotate_primitives((0.0, 0.0, 0.0), (0.0, 0.05, 0.0))
            h.compact_boxes(False)
            t1 = time.perf_counter()
            h.render_begin(0.0); h.render_end()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            step.append((t1 - t0) * 1e3); frame.append((t2 - t1) * 1e3); render.append(float(lib.b200_last_render_ms()))
        out["on_the_device"] = {"step_ms": round(min(step), 2), "of_which_device_work_ms": round(float(lib.b200_last_animation_ms()), 2),
                                "frame_ms": round(min(frame), 2), "of_which_render_kernels_ms": round(min(render), 2),
                                "walk_tree_nodes": upload_stats()[1]}
        h.close()
        return out

    # N > 1
    h = host.SceneHost(si, limits=wl["limits"])
    sc.replay(h)
    a = h.arrays()
    h.close()
    out = {"workload": wl["name"]}
    for name in ("every_rank_uploads", "root_uploads_others_receive_over_nccl"):
        e = engine.Engine(si, device=ctx.local_rank, limits=wl["limits"], rank=rank, world=world)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if name == "every_rank_uploads":
            e.upload(a, randoms=rnd)
            received = 0
        else:
            mine = a if rank == 0 else dict(a, boxes=np.zeros(0, np.uint8), primitives=np.zeros(0, np.uint8))
            received = partition.broadcast_scene(e, mine, rank, world, src=0, randoms=rnd)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out[name] = {"ms_max_over_ranks": round(float(dt.item()) * 1e3, 1), "bytes_received_per_rank": int(received)}
        e.close()
    return out


def parity_against_reference_cuda(key, frame):
    """The engine's frame (iteration 0) against the reference's own CUDA engine on the same GPU.  Checker only: runs after every
    timed region and in a child process — the reference's finalize_scene resets the device (CudaRayTracer.cu:1530), which would
    take this process's CUDA context with it."""
    import tempfile
    import refh
    from solr_b200 import workloads
    if frame is None or not refh.available("cuda"):
        return {"checked": False, "why": "reference CUDA build (oracle/_ref/libsolr_ref_cuda.so) not present"}
    wl = workloads.WORKLOADS[key]
    if wl["size"] != (1920, 1080) or wl["capacity"] is not None:
        return {"checked": False, "why": "beyond the reference's frame / box limits: tests/test_gpu_parity_at_size.py compares this size with the oracle"}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "frame.npz")
        np.savez(path, bm=frame[0], ids=frame[1])
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", key, "--parity-child", path],
                                 capture_output=True, text=True, timeout=600)
            return json.loads(out.stdout.strip().splitlines()[-1])
        except Exception as ex:   # the checker failing must not lose the measurement
            return {"checked": False, "why": "parity child failed: %r" % (ex,)}


def parity_child(key, path):
    import refh
    from solr_b200 import wire, workloads
    wl = workloads.WORKLOADS[key]
    z = np.load(path)
    bm, ids = z["bm"], z["ids"]
    si = workloads.scene_info(key)
    sc = wl["scene"]()
    rg = refh.RefScene(si, "cuda")
    sc.replay(rg)
    si.pathTracingIteration = 0
    gbm, gids, _ = rg.render(si, sc.eye, sc.target, sc.angles, randoms=np.zeros(wire.REF_MAX_BITMAP_SIZE, np.float32), block=(16, 8), want_post=False)
    rg.close()
    n = ids.shape[0] * ids.shape[1]
    return {"checked": True, "against": "reference CUDA engine (sm_100 build), same GPU, same scene and camera", "pixels": n,
            "ids_diff": int((ids[..., 0] != gids[..., 0]).sum()),
            "rgb_bad": int((np.abs(bm.astype(int) - gbm.astype(int)).max(-1) > 2).sum()), "rgb_bar": "2/255 per channel"}


def traffic_record():
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary_latest.json")) as f:
            j = json.load(f)
        return j.get("dram_bytes_per_frame", j.get("dram_bytes_per_launch")), {
            "file": "profiles/ncu_summary_latest.json", "captured_at_commit": j.get("git_head"), "captured": j.get("captured"),
            "note": "ncu --set full of one frame's ray kernels, dram__bytes_read.sum + dram__bytes_write.sum summed over the launches; "
                    "a recorded capture, not taken in this run"}
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4", "config5"])
    ap.add_argument("--no-sub", action="store_true", help="headline workload only")
    ap.add_argument("--parity-child", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    claim_stdout()

    from _solr_b200_import import solr_b200  # noqa: F401
    if args.parity_child:
        emit(parity_child(args.workload, args.parity_child))
        return
    if args.impl == "reference":
        import oracle
        oracle.load()   # builds oracle/libsolr_oracle.so if it is missing; nothing of the product is built or loaded
        run_reference_arm(args)
        return

    import __graft_entry__ as graft
    graft.build(quiet=True)
    import torch
    import torch.distributed as dist
    from solr_b200 import engine, workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this engine has no CPU fallback")
    ctx = Ctx()
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(ctx.local_rank)
    if ctx.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local_rank))
    assert ctx.world == args.gpus or ctx.world == 1, "launch with torchrun --nproc-per-node == --gpus"
    world, rank = ctx.world, ctx.rank
    ctx.stream = torch.cuda.Stream()
    ctx.lib = engine.load()
    ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    pk = peaks()
    flops_table = recorded_flops()

    key = args.workload
    head = measure_workload(key, args.steps, args.warmup, ctx, sample_clocks=True)
    frame = head.pop("_frame")

    subs = {}
    if not args.no_sub:
        sub_keys = [k for k in (["config4", "config5"] if world > 1 else ["config4", "config3", "config5"]) if k != key]
        for k in sub_keys:
            w, s = SUB_FRAMES[k]
            rec = measure_workload(k, s, w, ctx, sample_clocks=False)
            rec.pop("_frame")
            subs[k] = rec
        for k in ("config4", "config5"):
            if world > 1 and k in subs:
                # the other way to share configs 4 / 5 out: whole frames of different samples per GPU instead of tiles of one frame
                ss = measure_sample_split(k, ctx)
                ss["tile_split_ms_per_sample_same_iterations"] = subs[k]["ms_per_iteration"].get("11")
                subs[k]["sample_split"] = ss

    scene_paths = None
    if not args.no_sub:
        scene_paths = measure_scene_paths(ctx)

    # FP32 FMA microbenchmark on this GPU (dependent FFMA chains on every resident lane) and the clock it sustains
    fp32 = None
    if rank == 0:
        import ctypes as C
        tf, mhz = C.c_float(), C.c_float()
        if ctx.lib.b200_measure_fp32_peak(C.byref(tf), C.byref(mhz)) == 0:
            fp32 = {"ffma_tflops": round(tf.value, 2), "sm_mhz_under_that_load": round(mhz.value, 0)}

    if rank == 0:
        W, H = workloads.WORKLOADS[key]["size"]
        line = {"metric": "Mrays/s", "value": head["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": head["workload"], "rays_per_frame": head["rays_per_frame"], "mpixels_per_s": head["mpixels_per_s"],
                           "primitives": head["primitives"], "boxes": head["boxes"],
                           "l2": "flushed between timed frames (256 MiB memset outside the timed region)",
                           "parallelism": "one frame, interleaved 8x4 tiles over %d GPU(s)%s" % (
                               world, ", finished pixels stored into rank 0's frame over NVLink peer memory by the ray kernels, "
                               "one-element NCCL all-reduce as frame fence" if world > 1 else "")},
                "clocks": head["clocks"], "gpu_launches": head["gpu_launches"], "kernel_ms_last_frame": head["kernel_ms_last_frame"],
                "ms_per_iteration": head["ms_per_iteration"], "e2e": head["e2e"]}
        if "frame_check" in head:
            line["frame_check"] = head["frame_check"]

        def roofline(k, rec, flops_frame, source):
            achieved = flops_frame / (rec["ms_per_step"] * 1e-3) / 1e12 / world
            w_, h_ = rec["size"]
            return {"bound": "fp32", "achieved": achieved, "peak": pk["fp32_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["fp32_tflops"],
                    "peak_source": "148 SM x 128 lanes x 2 x sm_max_mhz (%s MEASURED_PEAKS.json), per GPU" % pk["source"],
                    "algorithmic_gflop_per_frame": flops_frame / 1e9, "algorithmic_flops_source": source,
                    "hbm": {"peak_gbs": pk["hbm_gbs"], "mandatory_bytes_per_frame": w_ * h_ * (32 + 16 + 3) * 2}}

        rec_flops = flops_table.get(key, {}).get("gflop_per_frame_mean_over_bench_iterations")
        flops_frame = rec_flops * 1e9 if rec_flops else None
        source = "oracle count in reference traversal order on sampled rows, recorded (profiles/algorithmic_flops.json)"
        if world == 1 and not args.no_cpu_baseline and (W, H) == (1920, 1080):
            cb = cpu_reference_run(key, 1, 0)
            flops_frame = cb["flops_per_frame"]
            source = "oracle count in reference traversal order, live on the whole %dx%d frame" % (W, H)
            line["cpu_baseline"] = {"value": cb["mrays_s"], "unit": "Mrays/s", "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"]}
        if flops_frame:
            # the frame is a handful of kernels of one code base (k_stage_primary, k_stage_pass per bounce, ...): the roofline is taken
            # over the timed region they fill, per GPU
            line["roofline"] = roofline(key, head, flops_frame, source)
            traffic, stamp = traffic_record()
            line["roofline"]["traffic"] = traffic if key == "config2" else None
            line["roofline"]["traffic_source"] = stamp if key == "config2" else None
            if fp32:
                line["roofline"]["measured_fp32"] = fp32
                line["roofline"]["frac_of_measured_ffma_rate"] = line["roofline"]["achieved"] / fp32["ffma_tflops"]
        if world == 1:
            line["parity"] = parity_against_reference_cuda(key, frame)
        for k, rec in subs.items():
            f = flops_table.get(k, {}).get("gflop_per_frame_mean_over_bench_iterations")
            rec["unit"] = "Mrays/s"
            if f:
                rec["roofline"] = roofline(k, rec, f * 1e9, "oracle count in reference traversal order on sampled rows, recorded (profiles/algorithmic_flops.json)")
        if subs:
            line["workloads"] = subs
        if scene_paths:
            line["scene_paths"] = scene_paths
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
