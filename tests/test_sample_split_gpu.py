"""GPU test of the sample split (include/solr_b200.h b200_accumulation_*, sol-r_b200/partition.py SampleSplit): two processes
share cuda:0, both render the shared iterations 0..10 of a progressive sequence over the whole frame, the accumulation iterations
11..K are dealt out between them, the partial sums are reduced onto rank 0 and packed — and the frame must equal the one a
single process accumulates sequentially, up to the order of the float additions (reference: CudaRayTracer.cu:550-562, k_default
:1066-1070).  gloo carries the reduce here because NCCL refuses two ranks on one device; bench.py --split samples runs it over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_scenes as gs
from solr_b200 import engine, host, partition

pytestmark = pytest.mark.gpu
FIRST = 11   # first iteration that only adds a sample (NB_MAX_ITERATIONS + 1)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, case, last, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc, si, eye, target, angles, rnd, frames = gs.case_setup(case)
    si.maxPathTracingIterations = last + 1
    h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
    e = engine.Engine(si)   # whole frame on every process
    e.upload(a, randoms=rnd)
    split = partition.SampleSplit(e.lib, rank, world, si.size.x, si.size.y)
    for it in range(0, FIRST):
        si.pathTracingIteration = it
        e.render(si, eye, target, angles)
    split.begin()
    mine = split.iterations(FIRST, last)
    for it in mine:
        si.pathTracingIteration = it
        e.render(si, eye, target, angles)
    split.finish(last)
    if rank == 0:
        merged = e.readback(si)[0].copy()
        post = e.read_post_buffer(si).copy()
    e.close()
    if rank == 0:
        e = engine.Engine(si)
        e.upload(a, randoms=rnd)
        for it in range(0, last + 1):
            si.pathTracingIteration = it
            e.render(si, eye, target, angles)
        whole = e.readback(si)[0].copy()
        wpost = e.read_post_buffer(si).copy()
        e.close()
        d = np.abs(merged.astype(int) - whole.astype(int)).max(-1)
        out["pixels"] = int(d.size)
        out["beyond_1"] = int((d > 1).sum())
        out["beyond_2"] = int((d > 2).sum())
        out["nonzero"] = int(np.count_nonzero(whole))
        out["sum_rel"] = float(np.abs(post[..., :3] - wpost[..., :3]).max() / max(1e-6, np.abs(wpost[..., :3]).max()))
        out["mine"] = mine
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,last", [("spheres_progressive", 13), ("spheres_progressive", 18)])
def test_sample_split_equals_sequential_accumulation(case, last):
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), case, last, out), nprocs=2, join=True)
    assert out["nonzero"] > 0
    # float sums in a different order; the reference's running maximum for pixels that see an emissive surface is per process
    assert out["beyond_2"] <= 1e-3 * out["pixels"], dict(out)
    assert out["beyond_1"] <= 1e-2 * out["pixels"], dict(out)
