"""Multi-GPU inference = batch sharding, one process per GPU, no collective on the data path.

Images are independent through forward, decode and NMS (BatchNorm in eval mode,
GRN / GroupNorm / SPR / TaskDecomposition statistics are per sample, NMS is per image), so rank r
simply processes images [lo, hi) of the global batch.  torch.distributed is used for rendezvous,
barriers, the max-over-ranks timing and (optionally) gathering the per-image results on rank 0.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ("shard_range", "gather_detections", "max_over_ranks")


def shard_range(total: int, rank: int, world: int):
    """Contiguous, balanced slice of `total` items for `rank` (first `total % world` ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_detections(local, group=None, dst: int = 0):
    """local: list of (n_i, 6) tensors for this rank's slice -> on `dst`, the list for the whole batch in
    global image order (None elsewhere).  Control-plane only (variable-length results)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(local)
    world = dist.get_world_size(group)
    out = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object([t.cpu() for t in local], out, dst=dst, group=group)
    if out is None:
        return None
    return [t for part in out for t in part]


def max_over_ranks(value: float, device=None, group=None) -> float:
    """The slowest rank's time: multi-GPU numbers are the max over ranks, never rank 0's own clock."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
