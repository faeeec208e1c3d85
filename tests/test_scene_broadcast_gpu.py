"""GPU test of the scene replication (include/solr_b200.h b200_scene_layout / b200_scene_adopt_layout / b200_scene_device_arrays /
b200_scene_adopt_finish, sol-r_b200/partition.py broadcast_scene): rank 0 uploads the scene, rank 1 — which is handed EMPTY box and
primitive arrays — adopts the layout and receives the device arrays with one broadcast per array; both then rotate the scene ON THE DEVICE and render
their tiles of the frame split, and the merged frame must equal the frame one process renders alone, bit for bit, with host-built and with
GPU-built walk trees.  gloo carries the broadcasts here because NCCL refuses two ranks on one device; bench.py runs the same path
with NCCL over NVLink."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_scenes as gs
from solr_b200 import engine, host, partition

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, case, gpu_trees, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc, si, eye, target, angles, rnd, frames = gs.case_setup(case)
    h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
    full = dict(a)
    if rank != 0:
        # this rank never sees the boxes or the primitives on the host
        a = dict(a); a["boxes"] = np.zeros(0, np.uint8); a["primitives"] = np.zeros(0, np.uint8)
    e = engine.Engine(si, rank=rank, world=world)
    e.set_option(10, gpu_trees)
    received = partition.broadcast_scene(e, a, rank, world, src=0, randoms=rnd)
    # the adopted scene animates on the device like the uploaded one (the maps of csrc/animate.cuh travelled with it)
    e.rotate_primitives((0.0, 0.0, 0.0), (0.1, 0.2, 0.0))
    e.render(si, eye, target, angles)
    bm, ids = e.readback(si)
    t = torch.from_numpy(bm.astype(np.int32))  # pixels this rank does not own are zero: the partial frames merge by summation
    dist.reduce(t, 0)
    stats = e.scene_stats()
    e.set_option(10, 0)
    e.close()
    if rank == 0:
        e = engine.Engine(si)
        e.upload(full, randoms=rnd)
        e.rotate_primitives((0.0, 0.0, 0.0), (0.1, 0.2, 0.0))
        e.render(si, eye, target, angles)
        whole = e.readback(si)[0].copy()
        e.close()
        out["differing"] = int(np.count_nonzero(t.numpy() != whole.astype(np.int32)))
        out["nonzero"] = int(np.count_nonzero(whole))
        out["received"] = received
        out["nodes"] = stats["walk_tree_nodes"]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("gpu_trees", [0, 1])
def test_broadcast_scene_renders_the_same_frame(gpu_trees):
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, "molecule_full", gpu_trees, out), nprocs=2, join=True)
    assert out["nonzero"] > 0 and out["received"] > 0 and out["nodes"] > 0
    assert out["differing"] == 0
