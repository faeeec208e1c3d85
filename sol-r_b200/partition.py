"""Frame split across GPUs: interleaved ownership of 8x4-pixel tiles, tile t -> rank t % world,
tiles numbered row-major.  Mirrors k_render's `tile = k * worldSize + rank` (csrc/engine.cu).  Frames merge
by summation because every rank leaves the pixels it does not own at zero.

The reference's dormant multi-GPU path splits contiguous row bands instead
(/root/reference/solr/engines/cuda/CudaRayTracer.cu:1694-1706, :1647-1672); cost per pixel varies by >10x between sky and
dense geometry, so bands balance badly."""
import numpy as np

TILE_W, TILE_H = 8, 4


def tile_grid(width, height):
    return (width + TILE_W - 1) // TILE_W, (height + TILE_H - 1) // TILE_H


def owner_map(width, height, world):
    """int32 [H, W]: owning rank of every pixel."""
    tx, ty = tile_grid(width, height)
    t = (np.arange(height)[:, None] // TILE_H) * tx + (np.arange(width)[None, :] // TILE_W)
    return (t % world).astype(np.int32)


def local_tile_count(width, height, rank, world):
    tx, ty = tile_grid(width, height)
    return (tx * ty - rank + world - 1) // world


class _DevicePtr:
    """Wraps a raw device pointer for torch.as_tensor via __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensors(engine, width, height):
    """Zero-copy torch views of the engine's device bitmap / id buffers (for NCCL collectives in place)."""
    import torch
    b, i, p = engine.device_buffers()
    bitmap = torch.as_tensor(_DevicePtr(b, (height * width * 3,), "|u1"), device="cuda")
    ids = torch.as_tensor(_DevicePtr(i, (height * width * 4,), "<i4"), device="cuda")
    return bitmap, ids


def merge_frames(bitmap, ids=None, dst=0):
    """The one exchange step of the path: sum-reduce the per-rank partial frames onto rank `dst`.
    Works on NCCL (device tensors, NVLink) and gloo (CPU tensors, tests)."""
    import torch.distributed as dist
    dist.reduce(bitmap, dst=dst, op=dist.ReduceOp.SUM)
    if ids is not None:
        dist.reduce(ids, dst=dst, op=dist.ReduceOp.SUM)
