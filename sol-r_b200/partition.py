"""Frame split across GPUs: interleaved ownership of 8x4-pixel tiles, tile t -> rank t % world,
tiles numbered row-major.  Mirrors k_render's `tile = k * worldSize + rank` (csrc/engine.cu).

The exchange step has two forms.  PeerFrame (the GPU path): the root GPU's device bitmap is mapped into every other
process and the ray kernels store finished pixels straight into it over NVLink, so the exchange is fused into the kernels and
only a stream-ordered barrier remains.  merge_frames (gloo tests, and NCCL as the plain form): every rank leaves the pixels it
does not own at zero and the partial frames are summed onto the root.

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


def broadcast_scene(engine, arrays, rank, world, src=0, randoms=None, textures=None):
    """Scene replication for the frame split (include/solr_b200.h b200_scene_layout ...): the root uploads the scene — host tree
    builds, PCIe copies — once; every other rank allocates device arrays of the same layout and receives them from the root's with
    one broadcast per array (NCCL over NVLink; the reference's dormant multi-GPU path uploads everything per device from the host,
    CudaRayTracer.cu:1540-1613).  `arrays` must be the same on every rank only as far as lights, materials and counts go: the boxes
    and primitives of non-root ranks are never read.  Returns the bytes received per rank."""
    import torch
    import torch.distributed as dist
    staged = dist.get_backend() != "nccl"  # gloo (tests: two ranks on one device) moves the arrays through host memory
    dev = torch.device("cpu") if staged else torch.device("cuda", torch.cuda.current_device())
    if rank == src:
        engine.upload(arrays, randoms=randoms, textures=textures)
        layout = torch.tensor(engine.scene_layout(), dtype=torch.int64, device=dev)
    else:
        engine.upload_small(arrays, randoms=randoms, textures=textures)
        layout = torch.zeros(32, dtype=torch.int64, device=dev)
    n = torch.tensor([layout.numel()], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    layout = layout[: int(n.item())].contiguous()
    dist.broadcast(layout, src)
    if rank != src:
        engine.adopt_layout([int(v) for v in layout.tolist()], int(arrays.get("nbLamps", 0)), int(arrays["lightInformationSize"]))
    total = 0
    for ptr, nbytes in engine.scene_device_arrays():
        if nbytes == 0:
            continue
        t = torch.as_tensor(_DevicePtr(ptr, (nbytes,), "|u1"), device="cuda")
        if staged:
            c = t.cpu()
            dist.broadcast(c, src)
            if rank != src:
                t.copy_(c)
        else:
            dist.broadcast(t, src)
        total += nbytes
    torch.cuda.synchronize()
    if rank != src:
        engine.adopt_finish()
    return total


def merge_frames(bitmap, ids=None, dst=0):
    """The one exchange step of the path: sum-reduce the per-rank partial frames onto rank `dst`.
    Works on NCCL (device tensors, NVLink) and gloo (CPU tensors, tests)."""
    import torch.distributed as dist
    dist.reduce(bitmap, dst=dst, op=dist.ReduceOp.SUM)
    if ids is not None:
        dist.reduce(ids, dst=dst, op=dist.ReduceOp.SUM)


class PeerFrame:
    """The fused exchange (include/solr_b200.h b200_peer_frame_*): rank `root` exports its device bitmap, the others open
    it, and from then on their kernels write finished pixels into the root's frame through NVLink peer memory.

    fence() is a one-element NCCL all-reduce ON THE STREAM THE ENGINE RENDERS ON: pass that torch stream as `stream` and the engine
    is switched to it (b200_set_stream), so the collective is ordered behind the kernels and their peer stores without a host
    round trip; without a stream the engine keeps its own and fence() drains it on the host before the collective (correct,
    slower).  Under gloo — tests with two processes on one GPU — it drains the engine's stream and meets the other ranks on the host.  A frame is: fence (the root has read the previous frame:
    nobody overwrites it early), render on every rank, fence (every rank's kernels — and with them their peer stores — are
    done), then the root's read-back."""

    HANDLE_BYTES = 64

    def __init__(self, lib, rank, world, root=0, stream=None):
        import ctypes
        import torch
        import torch.distributed as dist
        self.lib, self.rank, self.world, self.root, self.dist = lib, rank, world, root, dist
        self.torch, self.stream = torch, stream
        if stream is not None:
            lib.b200_set_stream(ctypes.c_void_p(stream.cuda_stream))
        box = [None]
        if rank == root:
            buf = ctypes.create_string_buffer(self.HANDLE_BYTES)
            if lib.b200_peer_frame_export(buf, self.HANDLE_BYTES) != 0:
                raise RuntimeError("b200_peer_frame_export failed")
            box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=root)
        if rank != root:
            self._handle = ctypes.create_string_buffer(box[0], self.HANDLE_BYTES)
            if lib.b200_peer_frame_open(self._handle, self.HANDLE_BYTES) != 0:
                raise RuntimeError("b200_peer_frame_open failed (no peer access to the root GPU?)")
        self.nccl = dist.get_backend() == "nccl"
        self.token = torch.zeros(1, dtype=torch.int32, device="cuda") if self.nccl else None

    def fence(self):
        if self.nccl and self.stream is not None:
            with self.torch.cuda.stream(self.stream):
                self.dist.all_reduce(self.token)
        elif self.nccl:
            self.lib.b200_synchronize()   # the engine renders on a stream of its own: its kernels first, then the collective
            self.dist.all_reduce(self.token)
            self.torch.cuda.current_stream().synchronize()
        else:
            self.lib.b200_synchronize()
            self.dist.barrier()

    def close(self):
        if self.rank != self.root:
            self.lib.b200_peer_frame_open(None, 0)


class SharedHostFrame:
    """One host frame for every process of a node (SceneHost::shareFrame, include/solr_b200.h b200_stream_target): when the frame's
    destination is HOST memory — the reference's render_end reads the frame and the id buffer back after every frame,
    CudaKernel.cpp:304-313 — no GPU needs another GPU's pixels.  Rank `root` creates a shared-memory segment, every rank maps and
    pins it, and from then on every GPU's ray kernels write the tiles they own straight into it over their own PCIe link, ids
    included.  A frame is: render_begin on every rank, fence() (every rank's kernels, and with them their writes, are done), then
    any rank reads `host_scene.bitmap()` / `primitive_ids()`: the whole frame.  fence() is PeerFrame's: a one-element all-reduce on
    the render stream under NCCL, a host barrier behind a drained stream under gloo."""

    def __init__(self, host_scene, lib, rank, world, root=0, stream=None, name=None):
        import os
        import ctypes
        import torch
        import torch.distributed as dist
        self.h, self.lib, self.rank, self.world, self.root, self.dist, self.torch, self.stream = host_scene, lib, rank, world, root, dist, torch, stream
        if stream is not None:
            lib.b200_set_stream(ctypes.c_void_p(stream.cuda_stream))
        box = [name or "/solr_b200_frame_%d" % os.getpid()]
        if world > 1:
            dist.broadcast_object_list(box, src=root)
        self.name = box[0]
        if rank == root:
            host_scene.share_frame(self.name, True)
        if world > 1:
            dist.barrier()
        if rank != root:
            host_scene.share_frame(self.name, False)
        if world > 1:
            dist.barrier()
        self.nccl = world > 1 and dist.get_backend() == "nccl"
        self.token = torch.zeros(1, dtype=torch.int32, device="cuda") if self.nccl else None

    def fence(self):
        if self.world == 1:
            self.lib.b200_synchronize()
        elif self.nccl and self.stream is not None:
            with self.torch.cuda.stream(self.stream):
                self.dist.all_reduce(self.token)
        elif self.nccl:
            self.lib.b200_synchronize()
            self.dist.all_reduce(self.token)
            self.torch.cuda.current_stream().synchronize()
        else:
            self.lib.b200_synchronize()
            self.dist.barrier()


class SampleSplit:
    """The second split north_star names: sample accumulation by GPU (include/solr_b200.h b200_accumulation_*).  Every process
    renders the WHOLE frame (engine partition 0 of 1).  Iterations 0..first-1 — the deepening passes and the first sample, whose
    per-pixel state (ids, first-hit depth) the later frames read — are rendered by everybody; the accumulation iterations
    first..last are dealt out round-robin; finish() sum-reduces the partial colour sums onto the root (NCCL over NVLink; gloo
    through host memory in the CPU-fenced tests) and the root stores them and packs the frame exactly as iteration `last` would.
    Equal to the sequential frames up to the order of the float additions (and, for pixels that see an emissive surface, the
    reference's running maximum over samples, CudaRayTracer.cu:553-557, which each process keeps over its own samples)."""

    def __init__(self, lib, rank, world, width, height, root=0):
        import torch
        import torch.distributed as dist
        self.lib, self.rank, self.world, self.root, self.dist, self.torch = lib, rank, world, root, dist, torch
        self.buf = torch.zeros(height * width * 4, dtype=torch.float32, device="cuda")
        self.nccl = dist.get_backend() == "nccl"

    def iterations(self, first, last):
        """This process's share of the accumulation iterations first..last."""
        return [it for it in range(first, last + 1) if (it - first) % self.world == self.rank]

    def begin(self):
        """Call after the shared iterations (0..first-1): the root keeps the first sample, the others start their sums at zero."""
        if self.rank != self.root:
            self.lib.b200_accumulation_clear()

    def finish(self, last):
        lib, torch = self.lib, self.torch
        if lib.b200_accumulation_export(self.buf.data_ptr()) != 0:
            raise RuntimeError("b200_accumulation_export failed")
        lib.b200_synchronize()   # the engine renders on its own stream
        if self.nccl:
            self.dist.reduce(self.buf, dst=self.root, op=self.dist.ReduceOp.SUM)
            torch.cuda.synchronize()
        else:
            host = self.buf.cpu()
            self.dist.reduce(host, dst=self.root, op=self.dist.ReduceOp.SUM)
            self.buf.copy_(host)
            torch.cuda.synchronize()
        if self.rank == self.root:
            if lib.b200_accumulation_import_and_pack(self.buf.data_ptr(), int(last)) != 0:
                raise RuntimeError("b200_accumulation_import_and_pack failed")
            lib.b200_synchronize()
