"""Multi-GPU plumbing for the batch-sharded replicas (no data-path collective).

One process per GPU (torchrun); each rank owns a full model replica and a contiguous shard of the image batch.
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used only for the timing barrier and the max-over-ranks
reduction of measured times.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `total` images for `rank` (first `total % world` ranks get one extra)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init(device: torch.device | None = None):
    rank, world, _ = env_rank()
    if world > 1 and not dist.is_initialized():
        if device is not None and device.type == "cuda":
            dist.init_process_group("nccl", device_id=device)
        else:
            dist.init_process_group("gloo")
    return rank, world


def barrier(device: torch.device | None = None):
    if dist.is_initialized():
        if device is not None and device.type == "cuda":
            dist.barrier(device_ids=[device.index])
        else:
            dist.barrier()
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(value: float, device: torch.device | None = None) -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None and device.type == "cuda" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shutdown():
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
