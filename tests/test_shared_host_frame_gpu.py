"""GPU test of the shared host frame (SceneHost::shareFrame, b200_stream_target, partition.SharedHostFrame): two processes share
cuda:0, each renders its interleaved tiles through the host drop-in, and each GPU's kernels write their tiles — pixels AND ids —
into ONE frame in shared host memory.  What the root finds there after the fence must equal what one process renders and reads
back alone (the reference's render_end, CudaKernel.cpp:304-313), over progressive frames, for a camera the staged kernels do not
serve (whole-frame form) and with an effect-free anaglyph frame.  gloo is the fence because NCCL refuses two ranks on one
device; bench.py runs the same path with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_scenes as gs
from solr_b200 import engine, host, partition

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _host(case, rank, world):
    sc, si, eye, target, angles, rnd, frames = gs.case_setup(case)
    si.maxPathTracingIterations = 1 << 30
    h = host.SceneHost(si, rank=rank, world=world)
    sc.replay(h)
    h.set_randoms(rnd, 0)
    h.set_camera(eye, target, angles)
    h.init_buffers()
    h.set_lazy_ids(False)
    return h, si


def _frames(h, si, iterations, fence=None):
    seen = []
    for it in iterations:
        si.pathTracingIteration = it
        h.set_scene_info(si)
        h.render_begin(0.0)
        if fence:
            fence()
        h.render_end()
        seen.append((h.bitmap().copy(), h.primitive_ids().copy()))
    return seen


def _worker(rank, world, port, case, iterations, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h, si = _host(case, rank, world)
    lib = engine.load()
    n0 = int(lib.b200_frames_streamed())
    shared = partition.SharedHostFrame(h, lib, rank, world)
    merged = []
    for it in iterations:
        si.pathTracingIteration = it
        h.set_scene_info(si)
        h.render_begin(0.0)
        shared.fence()                       # every rank's kernels, and their writes into the shared frame, are done
        if rank == 0:
            h.render_end()                   # waits for this rank's stream; copies nothing
            merged.append((h.bitmap().copy(), h.primitive_ids().copy()))
        dist.barrier()                       # nobody starts the next frame while the root reads this one
    streamed = int(lib.b200_frames_streamed()) - n0
    h.close()
    if rank == 0:
        h1, si1 = _host(case, 0, 1)
        whole = _frames(h1, si1, iterations)
        h1.close()
        out["bitmap_diff"] = [int(np.count_nonzero(m[0] != w[0])) for m, w in zip(merged, whole)]
        out["ids_diff"] = [int(np.count_nonzero(m[1] != w[1])) for m, w in zip(merged, whole)]
        out["nonzero"] = int(np.count_nonzero(whole[-1][0]))
        out["streamed"] = streamed
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,iterations,tiled", [("spheres_full", [0, 0], True), ("spheres_progressive", list(range(0, 13)), True),
                                                   ("spheres_anaglyph", [0, 1, 10, 11], True), ("spheres_aa", [0, 0], False)])
def test_two_processes_fill_one_host_frame(case, iterations, tiled):
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), case, iterations, out), nprocs=2, join=True)
    assert out["nonzero"] > 0
    assert out["bitmap_diff"] == [0] * len(iterations)
    assert out["ids_diff"] == [0] * len(iterations)
    assert out["streamed"] == (len(iterations) if tiled else 0)   # tile by tile from the ray kernels / whole after them


def test_one_process_shared_frame_equals_its_own_buffers():
    h, si = _host("mixed_full", 0, 1)
    own = _frames(h, si, [0, 0])
    lib = engine.load()
    partition.SharedHostFrame(h, lib, 0, 1)
    shared = _frames(h, si, [0, 0])
    h.close()
    for (b1, i1), (b2, i2) in zip(own, shared):
        assert np.array_equal(b1, b2) and np.array_equal(i1, i2)
