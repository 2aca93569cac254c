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


def gather_packed(out: torch.Tensor, counts: torch.Tensor, group=None, dst: int = 0):
    """Fixed-size form of gather_detections for the engine's packed result: out (B, max_det, 6) fp32 and counts (B,)
    int32 of every rank -> on `dst` the (world*B, max_det, 6) / (world*B,) tensors in global image order (None
    elsewhere).  One `gather` per tensor on the group's backend (NCCL for device tensors, gloo for host tensors)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return out, counts
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if rank == dst:
        outs = [torch.empty_like(out) for _ in range(world)]
        cnts = [torch.empty_like(counts) for _ in range(world)]
    else:
        outs = cnts = None
    dist.gather(out, outs, dst=dst, group=group)
    dist.gather(counts, cnts, dst=dst, group=group)
    if rank != dst:
        return None, None
    return torch.cat(outs, 0), torch.cat(cnts, 0)


def bind_rank_to_cores(local_rank: int, local_world: int):
    """Give every rank of a node its own contiguous share of the host cores this process may use (the NUMA-local
    ones when the GPU's node is visible in sysfs), so that eight ranks' pinned-memory H2D submit threads do not
    migrate over one another.  Returns the core list it set, or None when affinity cannot be changed."""
    import os
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    try:  # prefer the cores local to this rank's GPU
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if torch.cuda.is_available() else None
        if bus is not None:
            dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
            dev = torch.cuda.get_device_properties(local_rank).pci_device_id
            path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
            if os.path.isfile(path):
                local = set()
                for part in open(path).read().strip().split(","):
                    if "-" in part:
                        a, b = part.split("-")
                        local.update(range(int(a), int(b) + 1))
                    elif part:
                        local.add(int(part))
                if len(local & set(cores)) >= local_world:
                    cores = sorted(local & set(cores))
    except Exception:
        pass
    per = max(1, len(cores) // max(local_world, 1))
    mine = cores[local_rank * per:(local_rank + 1) * per] or cores
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    return mine


class ShardedEngine:
    """Multi-GPU inference API: a GLOBAL batch in, per-image results in global image order out.

    One process per GPU (torch.distributed.run).  Every rank is given the same global batch (or just its own slice)
    as a pinned host tensor; rank r runs images [lo, hi) = shard_range(B_global, r, world) through its local Engine
    (H2D -> stem -> graph replay -> NMS -> packed detections) and the packed results are gathered on `dst` in image
    order.  There is no collective on the data path; the gather moves 7.2 KB per image.
    """

    def __init__(self, engine, global_batch: int, group=None, dst: int = 0):
        self.engine, self.group, self.dst = engine, group, dst
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.global_batch = global_batch
        self.lo, self.hi = shard_range(global_batch, self.rank, self.world)
        if self.hi - self.lo != engine.batch:
            raise ValueError(f"rank {self.rank}: slice of {self.hi - self.lo} images but the Engine was built for {engine.batch}")

    def local_slice(self, batch: torch.Tensor) -> torch.Tensor:
        """This rank's images of a global batch (a view; a local slice is passed through)."""
        if batch.shape[0] == self.global_batch:
            return batch[self.lo:self.hi]
        if batch.shape[0] == self.hi - self.lo:
            return batch
        raise ValueError(f"expected {self.global_batch} (global) or {self.hi - self.lo} (local) images, got {batch.shape[0]}")

    def submit(self, batch: torch.Tensor):
        return self.engine.submit(self.local_slice(batch))

    def collect(self, slot):
        """-> on `dst`: list of B_global (n_i, 6) CPU tensors in image order; None on the other ranks."""
        if self.world == 1:
            return self.engine.collect(slot)
        slot.done.synchronize()
        # the device-side packed results are gathered, so at most ONE submission per engine slot may be in flight here
        # (the slot's next batch would overwrite them); the ticket is released for the engine's bookkeeping
        blob = getattr(slot, "blob", None)
        if blob is not None:   # detections + counts in one allocation: ONE gather per batch
            parts = [torch.empty_like(blob) for _ in range(self.world)] if self.rank == self.dst else None
            dist.gather(blob, parts, dst=self.dst, group=self.group)
            out = cnt = None
            if parts is not None:
                n_out = slot.out.numel()
                host = torch.stack(parts).cpu()
                out = host[:, :n_out].reshape(self.world * slot.out.shape[0], *slot.out.shape[1:])
                cnt = host[:, n_out:].contiguous().view(torch.int32).reshape(-1)
        else:
            out, cnt = gather_packed(slot.out, slot.counts, self.group, self.dst)
        if hasattr(slot, "k"):
            slot.slot.outstanding[slot.k] = False
        if out is None:
            return None
        out, cnt = out.cpu(), cnt.tolist()
        return [out[i, :cnt[i]] for i in range(len(cnt))]

    def __call__(self, batch: torch.Tensor):
        return self.collect(self.submit(batch))
