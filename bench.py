#!/usr/bin/env python
"""bench.py — Mrays/s of Sol-R's ray-propagation hot path on B200 (contract: see the task's bench section).

A *step* is one full frame of the workload through the hot path: per-pixel ray generation, box-list walk,
primitive tests, shading with shadow rays, reflection/refraction bounce loop, accumulation, RGB8 pack.
A *ray* is one box-list walk (closest-hit or shadow), SURVEY.md §8(d).

  workload (N=1 and N>1): BASELINE.json configs[1] — the molecule scene (105 k atom spheres + bond cylinders +
  ground + light = 216 k primitives), 1920x1080, 1 spp, shadows, 3 bounces (glFull, nbRayIterations=3),
  synthetic (the reference's 486-atom PDB file does not travel; sol-r_b200/scenes.py generates the lattice).

  value   whole-job Mrays/s with scene and frame state resident in HBM: K x b200_render, CUDA events on the
          render stream around each frame, L2 flushed (256 MiB memset) between frames, max over ranks.
          N>1: ONE frame is split across the N GPUs as interleaved 8x4-pixel tiles (strong scaling); the exchange is fused
          into the ray kernels — every rank stores its finished pixels straight into rank 0's device bitmap through NVLink
          peer memory (partition.PeerFrame) — and a one-element NCCL all-reduce inside the timed region orders the frame.
          After the timed runs rank 0 renders the whole frame alone and checks the merged frame against it, byte for byte.
  e2e     the same metric through the host-side drop-in (SceneHost.render_begin + render_end = the calls
          CudaKernel makes): per-frame parameter upload, kernels, device->host copy of RGB8 into caller-owned
          host memory, wall clock.  The id buffer (16 B per pixel) is read back on demand (getPrimitiveAt), as the
          drop-in does by default; the figure with the reference's every-frame id read-back is reported beside it.
  roofline  FP32 CUDA-core roofline (north_star: compute-bound, no tensor cores): algorithmic flops counted by
          the oracle in REFERENCE traversal order (SURVEY §8(d) constants) / device time / (148 SM x 128 lanes x
          2 x sm_max_mhz).  HBM traffic is reported beside it.
  cpu_baseline / --impl reference: the reference's own code on the host cores (oracle/_ref/libsolr_ref_cpu.so =
          its CUDA source compiled for the host, OpenMP over blocks) when that library travelled, else the
          oracle port; sample = one whole 1920x1080 frame of the same scene/camera per step (about 1.1-1.5 s on 16 cores).
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

W, H, NB_RAY_ITERATIONS = 1920, 1080, 3
SAMPLE_W, SAMPLE_H = 1920, 1080   # the CPU legs render the whole frame of the workload (about 1.5 s per frame on 16 cores)
WORKLOAD = "config2_molecule_216k_primitives_1920x1080_glFull_3_bounces"
# algorithmic work of one frame of this workload (SURVEY.md 8(d) flop weights x the oracle's counts in the reference's traversal
# order, full 1920x1080 frame; the N=1 run re-counts it live on the CPU sample)
ALGORITHMIC_GFLOP_PER_FRAME = 37.2259
SM_COUNT, LANES_PER_SM = 148, 128


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


# --workload: config2 is the configuration BASELINE.json's metric is quoted on (the default, and what the driver runs);
# config4 (1 M spheres, 3840x2160, iterations 10..13 = 4 accumulated samples, one frame per step) is the size north_star's
# 8-GPU efficiency target names.  The CPU legs and the roofline's flop count exist for config2 only.
WORKLOADS = {
    "config2": dict(name=WORKLOAD, size=(1920, 1080), nit=3, scene="config2", iterations=[0], limits=None, capacity=None),
    "config4": dict(name="config4_1M_spheres_3840x2160_glFull_3_bounces_accumulated_samples_iterations_10_to_13",
                    size=(3840, 2160), nit=3, scene="config4", iterations=[10, 11, 12, 13], limits=(3840, 2160),
                    capacity=(16_000_000, 4_000_000)),
}
SELECTED = WORKLOADS["config2"]


def scene_and_info(width, height):
    from solr_b200 import scenes, wire
    sc = getattr(scenes, SELECTED["scene"])()
    si = wire.default_scene_info(width, height, graphics_level=wire.GL_FULL, nb_ray_iterations=NB_RAY_ITERATIONS)
    return sc, si


def cpu_reference_run(steps, warmup, want_counts=True):
    """The reference arm / cpu_baseline: the path on host cores at SAMPLE_W x SAMPLE_H."""
    return _cpu_reference_run(steps, warmup)


def _cpu_reference_run(steps, warmup):
    import oracle
    import refh
    from solr_b200 import host, wire
    sc, si = scene_and_info(SAMPLE_W, SAMPLE_H)
    h = host.SceneHost(si)
    sc.replay(h)
    a = h.arrays()
    h.close()
    rnd = np.zeros(wire.REF_MAX_BITMAP_SIZE, np.float32)
    cores = host_cores()
    o = oracle.Oracle(a, SAMPLE_W, SAMPLE_H, randoms=rnd)
    t0 = time.perf_counter()
    o.render(si, sc.eye, sc.target, sc.angles, threads=cores)
    t_port = time.perf_counter() - t0
    counters = o.counters.as_dict()
    rays = counters["rays"]
    kind, times = "port", []
    if refh.available("cpu"):
        kind = "reference"
        r = refh.RefScene(si, "cpu")
        sc.replay(r)
        for k in range(warmup + steps):
            t0 = time.perf_counter()
            r.render(si, sc.eye, sc.target, sc.angles, randoms=rnd, block=(16, 8), want_post=False)
            if k >= warmup:
                times.append(time.perf_counter() - t0)
    else:
        for k in range(warmup + steps):
            o2 = oracle.Oracle(a, SAMPLE_W, SAMPLE_H, randoms=rnd)
            t0 = time.perf_counter()
            o2.render(si, sc.eye, sc.target, sc.angles, threads=cores)
            if k >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"kind": kind, "cores": cores, "rays_per_frame": rays, "sec_per_frame": sec, "mrays_s": rays / sec / 1e6,
            "port_sec_per_frame": t_port, "flops_per_frame": o.flops(), "counters": counters,
            "sample": "%dx%d frame of the same scene and camera (%s of the pixels), %d timed frames" % (
                SAMPLE_W, SAMPLE_H, "all" if SAMPLE_W * SAMPLE_H == W * H else "1/%d" % round(W * H / (SAMPLE_W * SAMPLE_H)), len(times))}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "Mrays/s", "value": r["mrays_s"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["sec_per_frame"] * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "bench_sample": r["sample"]},
            "cpu_baseline": {"value": r["mrays_s"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["mrays_s"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    global W, H, NB_RAY_ITERATIONS, WORKLOAD, SELECTED
    SELECTED = WORKLOADS[args.workload]
    (W, H), NB_RAY_ITERATIONS, WORKLOAD = SELECTED["size"], SELECTED["nit"], SELECTED["name"]
    if args.workload != "config2":
        args.no_cpu_baseline = True
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    claim_stdout()

    from _solr_b200_import import solr_b200  # noqa: F401
    import __graft_entry__ as graft
    graft.build(quiet=True)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from solr_b200 import engine, host, partition, wire

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this engine has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    pk = peaks()
    sc, si = scene_and_info(W, H)
    rnd = np.zeros(max(wire.REF_MAX_BITMAP_SIZE, W * H), np.float32)
    iterations = SELECTED["iterations"]
    frame_no = [0]

    def next_iteration():
        frame_no[0] += 1
        return iterations[frame_no[0] % len(iterations)]

    # ---- the drop-in host path (SceneHost -> C ABI) owns the engine in this process --------------------
    h = host.SceneHost(si, limits=SELECTED["limits"], rank=rank, world=world, device=local_rank, capacity=SELECTED["capacity"])
    sc.replay(h)
    h.set_randoms(rnd, 0)
    h.set_camera(sc.eye, sc.target, sc.angles)
    stream = torch.cuda.Stream()
    lib = engine.load()
    h.init_buffers()
    lib.b200_set_stream(stream.cuda_stream)
    si_live = h.scene_info
    si_live.maxPathTracingIterations = 1 << 30   # keep m_refresh true: every render_begin renders a frame

    def frame_e2e():
        si_live.pathTracingIteration = 0
        h.set_scene_info(si_live)
        h.render_begin(0.0)
        h.render_end()

    frame_e2e()   # uploads the scene (dirty flags), first frame
    torch.cuda.synchronize()
    eng = engine.Engine.__new__(engine.Engine)   # thin view on the already-initialised library for counters etc.
    eng.lib = lib
    eng.counters(reset=True)
    a = h.arrays()
    objects = wire.Int4(a["nbBoxes"], a["nbPrimitives"], a["nbLamps"], a["lightInformationSize"])
    occ = wire.Int2(1, 1)
    eye, target, angles = wire.Float3(*sc.eye), wire.Float3(*sc.target), wire.Float4(*sc.angles)
    pp = wire.PostProcessingInfo()
    si0 = wire.default_scene_info(W, H, graphics_level=wire.GL_FULL, nb_ray_iterations=NB_RAY_ITERATIONS)
    bitmap_t, ids_t = partition.device_tensors(eng, W, H)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    peer = partition.PeerFrame(lib, rank, world) if world > 1 else None   # ranks > 0 now write into rank 0's frame

    def frame_device():
        si0.pathTracingIteration = next_iteration()
        lib.b200_render(occ, wire.Int4(8, 4, 1, 0), si0, objects, pp, eye, target, angles)
        if world > 1:
            peer.fence()   # every rank's kernels, and with them their stores into rank 0's frame, are done

    # ---- device-resident timing --------------------------------------------------------------------------
    launches0 = eng.kernel_launches()
    sampler = ClockSampler(local_rank)
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            flush.zero_()
            frame_device()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.counters(reset=True)
        launches0 = eng.kernel_launches()
        sampler.start()
        evs = []
        for _ in range(args.steps):
            flush.zero_()                                  # L2 flush, outside the timed region
            if world > 1:
                dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            frame_device()
            e.record(stream)
            evs.append((s, e))
        stream.synchronize()
        torch.cuda.synchronize()
    clocks = sampler.summary()
    ms = [s.elapsed_time(e) for s, e in evs]
    launches = eng.kernel_launches() - launches0
    rays_local, px_local = eng.counters(reset=True)
    t = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
    r = torch.tensor([float(rays_local), float(px_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    rays_total, px_total = float(r[0].item()), float(r[1].item())
    ms_per_step = total_ms / args.steps
    value = rays_total / (total_ms * 1e-3) / 1e6
    kernel_ms = float(lib.b200_last_render_ms())

    # ---- end to end through the host drop-in -------------------------------------------------------------
    # Every rank runs the host drop-in for its tiles (render_begin: per-frame parameter upload + launch); the partial RGB8
    # frames are summed onto rank 0 over NVLink (the path's one exchange step) and rank 0 reads the merged frame and its
    # ids back to host memory (render_end = d2h_bitmap).  Wall clock around K frames, max over ranks.
    def frame_e2e_all(iteration=None):
        si_live.pathTracingIteration = next_iteration() if iteration is None else iteration
        h.set_scene_info(si_live)
        if world > 1:
            with torch.cuda.stream(stream):
                peer.fence()   # rank 0 has read the previous frame back
        h.render_begin(0.0)
        if world > 1:
            with torch.cuda.stream(stream):
                peer.fence()   # the frame is complete in rank 0's device bitmap
        if rank == 0:
            h.render_end()
        else:
            stream.synchronize()

    def measure_e2e():
        for _ in range(3):
            frame_e2e_all()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        eng.counters(reset=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            frame_e2e_all()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rays_e2e, _ = eng.counters(reset=True)
        te = torch.tensor([dt], dtype=torch.float64, device="cuda")
        re_ = torch.tensor([float(rays_e2e)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(re_, op=dist.ReduceOp.SUM)
        return float(re_.item()) / float(te.item()) / 1e6, float(te.item()) / args.steps * 1e3

    # the drop-in's default: render_end reads the pixels back, the id buffer stays on the device until getPrimitiveAt asks (its only
    # host-side reader, GPUKernel.cpp:729-739); then the reference's own protocol — pixels and 16 bytes of ids per pixel, every frame
    e2e_value, e2e_ms = measure_e2e()
    h.set_lazy_ids(False)
    eager_value, eager_ms = measure_e2e()
    h.set_lazy_ids(True)
    frame_e2e_all(0)   # a single-sample frame on every rank: what frame_check compares
    e2e = {"value": e2e_value, "unit": "Mrays/s", "ms_per_frame": e2e_ms,
           "h2d_bytes_per_step": int(lib.b200_frame_parameter_bytes()) * world,   # scene-info + camera + pointers block, per frame and rank
           "d2h_bytes_per_step": W * H * 3,   # RGB8 into caller-owned memory (rank 0)
           "with_id_buffer_every_frame": {"value": eager_value, "ms_per_frame": eager_ms, "d2h_bytes_per_step": W * H * 3 + W * H * 16,
                                          "note": "the reference's render_end protocol (CudaKernel.cpp:304-313)"}}

    frame_check = None
    if world > 1:
        # the merged frame rank 0 just read back against the whole frame rendered by rank 0 alone
        dist.barrier()
        peer.close()
        if rank == 0:
            merged = np.array(h.bitmap(), copy=True)
            lib.b200_set_partition(0, 1)
            frame_e2e()
            whole = np.array(h.bitmap(), copy=True)
            differing = int(np.count_nonzero(merged != whole))
            frame_check = {"bytes_differing_from_the_one_gpu_frame": differing, "bytes": int(whole.size)}
            if differing:
                raise SystemExit("bench.py: the frame merged over %d GPUs differs from the one-GPU frame in %d bytes" % (world, differing))
        dist.barrier()
    if rank == 0:
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "rays_per_frame": rays_total / args.steps, "mpixels_per_s": px_total / (total_ms * 1e-3) / 1e6,
                           "l2": "flushed between timed frames (256 MiB memset outside the timed region)",
                           "parallelism": "one frame, interleaved 8x4 tiles over %d GPU(s)%s" % (
                               world, ", finished pixels stored into rank 0's frame over NVLink peer memory by the ray kernels, "
                               "one-element NCCL all-reduce as frame fence" if world > 1 else "")},
                "clocks": clocks, "gpu_launches": int(launches), "kernel_ms_last_frame": kernel_ms}
        line["e2e"] = e2e
        if frame_check is not None:
            line["frame_check"] = frame_check
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_summary_latest.json")) as f:
                j = json.load(f)
                traffic = j.get("dram_bytes_per_frame", j.get("dram_bytes_per_launch"))   # summed over the frame's launches
        except Exception:
            pass
        flops_frame = ALGORITHMIC_GFLOP_PER_FRAME * 1e9
        flops_source = "oracle count in reference traversal order, full frame (recorded)" if args.workload == "config2" else "not counted for this workload"
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_run(1, 0)
            flops_frame = cb["flops_per_frame"] * (W * H) / float(SAMPLE_W * SAMPLE_H)
            flops_source = "oracle count in reference traversal order, live on the %dx%d CPU frame" % (SAMPLE_W, SAMPLE_H)
            line["cpu_baseline"] = {"value": cb["mrays_s"], "unit": "Mrays/s", "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"]}
        # the frame is a handful of kernels of one code base (k_stage_primary, k_stage_pass per bounce, ...): the roofline is
        # taken over the timed region they fill, per GPU
        achieved = flops_frame / (ms_per_step * 1e-3) / 1e12 / world
        line["roofline"] = {"bound": "fp32", "achieved": achieved, "peak": pk["fp32_tflops"], "unit": "TFLOP/s",
                            "frac": achieved / pk["fp32_tflops"], "traffic": traffic,
                            "peak_source": "148 SM x 128 lanes x 2 x sm_max_mhz (%s MEASURED_PEAKS.json), per GPU" % pk["source"],
                            "algorithmic_gflop_per_frame": flops_frame / 1e9, "algorithmic_flops_source": flops_source,
                            "hbm": {"peak_gbs": pk["hbm_gbs"], "mandatory_bytes_per_frame": W * H * (32 + 16 + 3) * 2}}
        if args.workload != "config2":
            line["roofline"].update(achieved=None, frac=None, traffic=None, algorithmic_gflop_per_frame=None)
        emit(line)
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
