"""CPU, world_size 2 over gloo: the batch-sharding host logic of the N>1 path."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mgdt_yolo_b200.parallel import gather_detections, max_over_ranks, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 32, 255, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


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
        dist.barrier()
        if rank == 0:
            ok = len(allr) == total and all(t.shape == (i, 6) and (t == i).all() for i, t in enumerate(allr))
            q.put((ok, slow))
        else:
            assert allr is None and slow == 10.0 + world - 1
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
