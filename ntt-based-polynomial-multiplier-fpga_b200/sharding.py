"""Batch sharding across GPUs: one process per GPU, contiguous row slices, NO data-path
collective (independent polynomial products share nothing but read-only tables --
SURVEY.md section 8e).  torch.distributed is used only for the start/stop barrier and the
max-over-ranks reduction of the timing."""
from __future__ import annotations

import os


def shard_bounds(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Rows [lo, hi) of rank `rank`: GPU g gets rows [g*B/G, (g+1)*B/G)."""
    if world < 1 or not (0 <= rank < world) or batch < 0:
        raise ValueError("bad shard request")
    return (batch * rank) // world, (batch * (rank + 1)) // world


def env_rank_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_distributed(backend: str):
    """Initialise torch.distributed from the torchrun environment (MASTER_ADDR defaults to
    127.0.0.1: the container hostname may not resolve)."""
    import torch.distributed as dist
    rank, world, _ = env_rank_world()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    # bench.py prints exactly one JSON line on stdout: keep NCCL's "NCCL version ..." banner
    # (printed at NCCL_DEBUG=VERSION and at WARN) out of it; whatever NCCL has to say goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1 and not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return dist


def barrier(device=None) -> None:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        if device is not None and dist.get_backend() == "nccl":
            dist.barrier(device_ids=[device])
        else:
            dist.barrier()


def max_over_ranks(value: float, device="cpu") -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_over_ranks(value: float, device="cpu") -> list:
    """Every rank's value, in rank order (diagnostics: which rank sets the max-over-ranks time)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [value]
    t = torch.tensor([value], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]
