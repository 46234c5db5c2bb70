"""World-size-2 gloo test of the multi-GPU host logic (sharding, boundary hand-off, histogram combine) on CPU tensors."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcaller_b200 import dist as mdist


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = mdist.shard_range(101, rank, world)
        depth = torch.zeros(16, dtype=torch.int32)
        meth = torch.zeros(16, dtype=torch.int32)
        first = torch.full((16,), 2 ** 63 - 1, dtype=torch.int64)
        for i in range(lo, hi):
            s = i % 16
            depth[s] += 1
            meth[s] += i % 2
            first[s] = min(int(first[s]), mdist.rank_row_base(rank) + (i - lo))
        mdist.allreduce_histogram(depth, meth, first)
        # rank 0 leaves one window pending; rank 1's first kept line lies on contig 3 and it has nothing pending
        res = mdist.exchange_boundaries(3 if rank == 1 else 0, 1 if rank == 0 else 0, torch.device("cpu"))
        q.put((rank, lo, hi, depth.tolist(), meth.tolist(), first.tolist(), res))
    finally:
        dist.destroy_process_group()


def test_two_rank_histogram_and_boundary():
    world, port = 2, 29000 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, d0, m0, f0, b0), (r1, lo1, hi1, d1, m1, f1, b1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 50, 50, 101)
    assert d0 == d1 and m0 == m1 and f0 == f1
    want_d = [sum(1 for i in range(101) if i % 16 == s) for s in range(16)]
    want_m = [sum(i % 2 for i in range(101) if i % 16 == s) for s in range(16)]
    assert d0 == want_d and m0 == want_m
    assert f0 == list(range(16))                 # every slot is first seen by rank 0's rows 0..15
    assert b0 == (1, 3)                          # rank 0's pending window is closed by rank 1's first kept line
    assert b1 == (0, -1)                         # last rank: dropped, like the last window of a file


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 100, 1001):
        for w in (1, 2, 4, 8):
            parts = [mdist.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))


def test_plan_chunks_cuts_on_read_boundaries():
    import numpy as np
    from mcaller_b200.stream import plan_chunks
    offs = np.array([0, 100, 250, 900, 1000, 1800])
    cuts = plan_chunks(offs, 2000, 500)
    assert cuts[-1] == 2000 and all(c in set(offs.tolist()) | {2000} for c in cuts)
    assert cuts == [250, 900, 1000, 1800, 2000]   # read [250,900) is larger than the chunk and goes alone
    assert plan_chunks(offs, 2000, 5000) == [2000]
