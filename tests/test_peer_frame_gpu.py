"""GPU test of the fused frame exchange (include/solr_b200.h b200_peer_frame_*, sol-r_b200/partition.py PeerFrame):
two processes share cuda:0, each renders its interleaved tiles, rank 1's kernels store their pixels into rank 0's device
bitmap through the inter-process mapping, and rank 0's read-back must equal the frame one process renders alone — over
progressive frames too (pixels that drop out of the deepening passes keep the value they were last given).
gloo is the fence here because NCCL refuses two ranks on one device; bench.py runs the same path with NCCL over NVLink."""
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


def _setup(case):
    sc, si, eye, target, angles, rnd, frames = gs.case_setup(case)
    h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
    return si, eye, target, angles, rnd, a


def _worker(rank, world, port, case, iterations, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    si, eye, target, angles, rnd, a = _setup(case)
    e = engine.Engine(si, rank=rank, world=world)
    e.upload(a, randoms=rnd)
    peer = partition.PeerFrame(e.lib, rank, world)
    merged = []
    for it in iterations:
        si.pathTracingIteration = it
        peer.fence()
        e.render(si, eye, target, angles)
        peer.fence()
        if rank == 0:
            merged.append(e.readback(si)[0].copy())
    peer.fence()
    peer.close()
    e.close()
    if rank == 0:
        whole = []
        e = engine.Engine(si)
        e.upload(a, randoms=rnd)
        for it in iterations:
            si.pathTracingIteration = it
            e.render(si, eye, target, angles)
            whole.append(e.readback(si)[0].copy())
        e.close()
        out["differing"] = [int(np.count_nonzero(m != w)) for m, w in zip(merged, whole)]
        out["nonzero"] = int(np.count_nonzero(whole[-1]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,iterations", [("spheres_full", [0]), ("spheres_progressive", list(range(0, 13)))])
def test_peer_frame_equals_the_one_process_frame(case, iterations):
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), case, iterations, out), nprocs=2, join=True)
    assert out["nonzero"] > 0
    assert out["differing"] == [0] * len(iterations)
