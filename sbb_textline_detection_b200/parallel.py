"""Multi-GPU plumbing: pages are independent (``textline_detector.run()`` is per image, main.py:2056),
so the THROUGHPUT path shards page-per-GPU with one process per GPU and NO data-path collective.  The only
exchange is at init: rank 0 broadcasts the packed frozen weights once (NCCL over NVLink on GPUs,
gloo on CPU for the tests).

LATENCY mode for one large page (BASELINE config 5): the tiles of a page are independent too
and every page pixel is owned by exactly one tile (main.py:294-364), so the ranks take contiguous tile ranges
of the reference's loop order and stitch into ONE label map: ``PageSharder(mode="p2p")`` maps the owner
rank's label buffer into every process (CUDA IPC) and each rank's fused head epilogue stores its pixels
straight into it over NVLink -- the stitch IS the exchange, there is no collective kernel;
``mode="allreduce"`` is the plain-NCCL alternative (MAX all-reduce of per-rank maps) it is measured against."""
from __future__ import annotations

import os

import numpy as np


def shard_pages(n_pages: int, rank: int, world: int):
    """Page p -> rank p mod world.  Returns this rank's page indices."""
    return list(range(rank, n_pages, world))


def init_distributed(backend: str | None = None):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* from the env (torchrun).  Returns
    (rank, world, local_rank); a single-process run needs no process group."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_blob(blob: bytes | None, src: int = 0, device=None) -> bytes:
    """Broadcast the packed weight blob from ``src`` to every rank (size first, then payload).
    Non-source ranks pass ``None``.  With world size 1 this is the identity."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        assert blob is not None
        return blob
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device(device if device is not None else ("cuda" if on_gpu else "cpu"))
    rank = dist.get_rank()
    n = torch.tensor([len(blob) if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src=src)
    if rank == src:
        payload = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    else:
        payload = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(payload, src=src)
    return blob if rank == src else payload.cpu().numpy().tobytes()


class AbiCommunicator:
    """An NCCL communicator created through the C ABI (sbb_nccl_comm_create): what a host WITHOUT torch would use.
    The 128-byte unique id travels from rank 0 to the other ranks over the process group that is already there
    (any launcher-provided channel would do)."""

    def __init__(self, device: int):
        import ctypes as C

        import torch.distributed as dist
        from . import _lib
        self._lib, self._C, self.device = _lib, C, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        box = [None]
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            _lib.check(_lib.lib().sbb_nccl_unique_id(buf))
            box[0] = bytes(buf)
        dist.broadcast_object_list(box, src=0)
        ident = (C.c_uint8 * 128).from_buffer_copy(box[0])
        comm = C.c_void_p()
        _lib.check(_lib.lib().sbb_nccl_comm_create(ident, self.world, self.rank, device, C.byref(comm)))
        self.comm = comm

    def broadcast_blob(self, blob: bytes | None, src: int = 0) -> bytes:
        """include/sbb_textline.h: sbb_model_broadcast -- every rank gets rank ``src``'s packed weight blob."""
        import torch.distributed as dist
        C = self._C
        n = [len(blob) if self.rank == src else 0]
        dist.broadcast_object_list(n, src=src)
        buf = bytearray(blob) if self.rank == src else bytearray(n[0])
        arr = (C.c_uint8 * n[0]).from_buffer(buf)
        self._lib.check(self._lib.lib().sbb_model_broadcast(arr, n[0], src, self.comm, self.device, None))
        del arr
        return blob if self.rank == src else bytes(buf)

    def close(self):
        if self.comm:
            self._lib.check(self._lib.lib().sbb_nccl_comm_destroy(self.comm))
            self.comm = None


def all_reduce_max(value: float) -> float:
    """MAX over ranks of a host scalar (bench.py: device-timed step, max over ranks)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_reduce_sum(value: float) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def tile_ranges(n_tiles: int, world: int):
    """Contiguous split of the reference's tile loop order: rank r -> (first, count)."""
    b = [r * n_tiles // world for r in range(world + 1)]
    return [(b[r], b[r + 1] - b[r]) for r in range(world)]


class PeerBuffer:
    """A device byte buffer owned by rank ``owner`` and mapped into every rank (CUDA IPC via the C ABI).
    ``ptr`` is valid on this rank's device; ``tensor`` (owner only) is a zero-copy torch view."""

    def __init__(self, nbytes: int, owner: int, device: int):
        import ctypes as C

        import torch
        import torch.distributed as dist
        from . import _lib
        self._lib, self._C = _lib, C
        self.owner, self.nbytes, self.device = owner, nbytes, device
        self.rank = dist.get_rank()
        handle = torch.zeros(64, dtype=torch.uint8, device=f"cuda:{device}")
        p = C.c_void_p()
        if self.rank == owner:
            hbuf = (C.c_uint8 * 64)()
            _lib.check(_lib.lib().sbb_peer_alloc(device, nbytes, C.byref(p), hbuf))
            handle.copy_(torch.frombuffer(bytearray(hbuf), dtype=torch.uint8))
        dist.broadcast(handle, src=owner)
        if self.rank != owner:
            hbuf = (C.c_uint8 * 64).from_buffer_copy(bytes(handle.cpu().numpy().tobytes()))
            _lib.check(_lib.lib().sbb_peer_open(device, hbuf, C.byref(p)))
        self.ptr = int(p.value)

    def tensor(self, shape):
        """Owner rank: torch uint8 view of the buffer."""
        import torch

        class _Iface:
            pass
        o = _Iface()
        o.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (self.ptr, False), "version": 2}
        return torch.as_tensor(o, device=f"cuda:{self.device}")

    def close(self):
        if self.ptr:
            f = self._lib.lib().sbb_peer_free if self.rank == self.owner else self._lib.lib().sbb_peer_close
            self._lib.check(f(self._C.c_void_p(self.ptr)))
            self.ptr = 0


class PageSharder:
    """One page across all ranks (latency mode).  Every rank calls ``run`` with the SAME page (rank ``owner``'s
    copy is broadcast when ``broadcast_page``); the stitched label map is returned on ``owner`` (None elsewhere)."""

    def __init__(self, model, H: int, W: int, owner: int = 0, mode: str = "p2p", margin: int = -1):
        import torch
        import torch.distributed as dist
        from .model import compute_tile_grid
        assert mode in ("p2p", "allreduce")
        self.model, self.H, self.W, self.owner, self.mode, self.margin = model, H, W, owner, mode, margin
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        nx, ny, _, _, _ = compute_tile_grid(H, W, model.tile_h, model.tile_w, margin)
        self.first, self.count = tile_ranges(nx * ny, self.world)[self.rank]
        self.dev = torch.device("cuda", model.device)
        if mode == "p2p":
            self.buf = PeerBuffer(H * W, owner, model.device)
            self.local = self.buf.tensor((H, W)) if self.rank == owner else None
        else:
            self.buf = None
            self.local = torch.empty((H, W), dtype=torch.uint8, device=self.dev)

    def run(self, page, broadcast_page: bool = True):
        import torch
        import torch.distributed as dist
        if broadcast_page:
            dist.broadcast(page, src=self.owner)
        if self.mode == "allreduce":
            self.model.predict_page_tile_range(page, self.local, self.first, self.count, keep_labels=False, margin=self.margin)
            dist.all_reduce(self.local, op=dist.ReduceOp.MAX)      # disjoint owners: MAX == union
            return self.local if self.rank == self.owner else None
        if self.rank == self.owner:
            self.local.zero_()
        torch.cuda.synchronize(self.dev)
        dist.barrier()                                             # the map is clear before anyone stores into it
        self.model.predict_page_tile_range(page, self.buf.ptr, self.first, self.count, keep_labels=True, margin=self.margin)
        torch.cuda.synchronize(self.dev)                           # this rank's peer stores are complete
        dist.barrier()
        return self.local if self.rank == self.owner else None

    def close(self):
        if self.buf is not None:
            self.buf.close()
