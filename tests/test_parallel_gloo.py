"""CPU, world_size 2 over gloo: the batch-sharding host logic of the N>1 path."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mgdt_yolo_b200.parallel import ShardedEngine, gather_detections, gather_packed, max_over_ranks, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 32, 255, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeSlot:
    pass


class _FakeEngine:
    """Engine stand-in (no GPU here): image with value v -> (local index + 1) rows filled with v."""

    def __init__(self, batch):
        self.batch = batch

    def submit(self, x):
        s = _FakeSlot()
        s.done = type("E", (), {"synchronize": lambda self: None})()
        s.out = x.view(-1, 1, 1).expand(-1, 4, 6).contiguous()
        s.counts = torch.arange(1, x.shape[0] + 1, dtype=torch.int32)
        return s


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 7
        lo, hi = shard_range(total, rank, world)
        # stand-in for Engine output: image i yields i rows whose first column is the global image index
        local = [torch.full((i, 6), float(i)) for i in range(lo, hi)]
        allr = gather_detections(local, dst=0)
        slow = max_over_ranks(10.0 + rank)
        # fixed-size packed results (the Engine's output format) gathered in global image order
        po, pc = gather_packed(torch.full((3, 4, 6), float(rank)), torch.full((3,), rank, dtype=torch.int32))
        # the sharding API over a stand-in engine: global batch of 6 -> slices of 3 -> ordered results on rank 0
        se = ShardedEngine(_FakeEngine(3), 6)
        res = se(torch.arange(6, dtype=torch.float32).view(6, 1))
        dist.barrier()
        if rank == 0:
            ok = len(allr) == total and all(t.shape == (i, 6) and (t == i).all() for i, t in enumerate(allr))
            ok = ok and po.shape == (6, 4, 6) and pc.tolist() == [0, 0, 0, 1, 1, 1] and bool((po[3:] == 1).all())
            ok = ok and len(res) == 6 and all(t.shape == (i % 3 + 1, 6) and bool((t == float(i)).all()) for i, t in enumerate(res))
            q.put((ok, slow))
        else:
            assert allr is None and slow == 10.0 + world - 1 and po is None and res is None
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, slow = q.get(timeout=10)
    assert ok and slow == 11.0
