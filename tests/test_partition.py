"""Multi-GPU host logic on CPU: tile ownership is a partition of the frame, and world_size-2 gloo merging
of per-rank partial frames (non-owned pixels zero) reproduces the full frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_scenes as gs
import oracle
from solr_b200 import host, partition


@pytest.mark.parametrize("wh", [(96, 72), (1920, 1080), (1024, 768), (37, 21)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_owner_map_is_a_balanced_partition(wh, world):
    W, H = wh
    m = partition.owner_map(W, H, world)
    assert m.shape == (H, W) and m.min() >= 0 and m.max() <= world - 1
    tx, ty = partition.tile_grid(W, H)
    counts = [partition.local_tile_count(W, H, r, world) for r in range(world)]
    assert sum(counts) == tx * ty and max(counts) - min(counts) <= 1
    # pixel-level: every tile is owned by exactly one rank, and counts agree with the kernel's formula
    tiles = (np.arange(H)[:, None] // 4) * tx + (np.arange(W)[None, :] // 8)
    for r in range(world):
        assert len(np.unique(tiles[m == r])) == counts[r]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, full_bitmap, full_ids, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, W = full_bitmap.shape[:2]
    mine = torch.from_numpy(partition.owner_map(W, H, world) == rank)
    bm = torch.from_numpy(full_bitmap).clone(); bm[~mine] = 0
    ids = torch.from_numpy(full_ids).clone(); ids[~mine] = 0
    partition.merge_frames(bm.view(-1), ids.view(-1), dst=0)
    if rank == 0:
        out["bitmap_ok"] = bool(np.array_equal(bm.numpy(), full_bitmap))
        out["ids_ok"] = bool(np.array_equal(ids.numpy(), full_ids))
    dist.destroy_process_group()


def test_world2_gloo_merge_reproduces_the_frame():
    sc, si, eye, target, angles, rnd, frames = gs.case_setup("spheres_full")
    h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
    o = oracle.Oracle(a, si.size.x, si.size.y, randoms=rnd)
    o.render(si, eye, target, angles, threads=2)
    mgr = mp.Manager(); out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, o.bitmap.copy(), o.ids.copy(), out), nprocs=2, join=True)
    assert out["bitmap_ok"] and out["ids_ok"]


def test_sample_split_shares_partition_the_iterations():
    """partition.SampleSplit deals the accumulation iterations out round-robin: every iteration to exactly one process."""
    class _S(partition.SampleSplit):
        def __init__(self, rank, world):
            self.rank, self.world = rank, world
    for world in (1, 2, 3, 8):
        for first, last in ((11, 15), (11, 11), (11, 110)):
            shares = [_S(r, world).iterations(first, last) for r in range(world)]
            flat = sorted(it for s in shares for it in s)
            assert flat == list(range(first, last + 1))
            assert max(len(s) for s in shares) - min(len(s) for s in shares) <= 1
