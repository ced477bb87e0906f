"""One-process-per-GPU plumbing for the batch-sharded inference / explanation path.

The path has NO data-path collective: every rank owns `batch_per_gpu` images, weights are
replicated.  torch.distributed is used only to line ranks up (barrier) and to take the max of the device-side
timings; NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: str | None = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, local_rank, world


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value: float) -> float:
    """Slowest rank's value (device-side milliseconds) - the number a multi-GPU measurement must report."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of `total` images owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_to_rank0(t: torch.Tensor):
    """Host-side gather used only by parity checks (results are independent per image)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [t]
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(t.cpu(), out, dst=0)
    return out


def shutdown() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
