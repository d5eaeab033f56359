"""Multi-GPU plumbing: pages are independent (``textline_detector.run()`` is per image, main.py:2056),
so the path shards page-per-GPU with one process per GPU and NO data-path collective.  The only
exchange is at init: rank 0 broadcasts the packed frozen weights once (NCCL over NVLink on GPUs,
gloo on CPU for the tests)."""
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
