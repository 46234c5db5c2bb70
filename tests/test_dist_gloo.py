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
        counts = torch.zeros(32, dtype=torch.int32)               # depth | meth packed like Engine.d_counts
        depth, meth = counts[:16], counts[16:]
        first = torch.full((16,), 2 ** 63 - 1, dtype=torch.int64)
        for i in range(lo, hi):
            s = i % 16
            depth[s] += 1
            meth[s] += i % 2
            first[s] = min(int(first[s]), mdist.rank_row_base(rank) + (i - lo))
        mdist.allreduce_histogram(counts, first)
        # rank 0's slice starts on contig 0, rank 1's first kept line lies on contig 3: rank 0's open window is closed by
        # contig 3, nobody closes the last rank's (dropped like the last window of a file)
        allk = mdist.gather_first_kept(torch.tensor([3 if rank == 1 else 0], dtype=torch.int64))
        q.put((rank, lo, hi, depth.tolist(), meth.tolist(), first.tolist(), (allk.tolist(), mdist.closing_contig(allk, rank))))
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
    assert b0 == ([0, 3], 3)                     # rank 0's pending window is closed by rank 1's first kept line
    assert b1 == ([0, 3], -1)                    # last rank: dropped, like the last window of a file


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


def test_reference_byte_ranges_and_merge(tmp_path):
    """multigpu.byte_ranges is the reference's split (mCaller.py:63-68); merge_outputs concatenates the ranks' files in
    offset order and removes them."""
    import math
    from mcaller_b200 import multigpu
    for size in (0, 1, 10, 1001, 12345):
        for w in (1, 2, 3, 8):
            r = multigpu.byte_ranges(size, w)
            chunk = int(math.ceil(size / float(w))) if size else 0
            assert [a for a, _ in r] == [chunk * i for i in range(w)] and r[-1][1] == size
    tsv = tmp_path / "x.eventalign.tsv"
    tsv.write_bytes(b"y" * 1001)
    for i, (a, _) in enumerate(multigpu.byte_ranges(1001, 3)):
        open(multigpu.ec_tmp_name(str(tsv), 6, a), "w").write("rank%d\n" % i)
    out = multigpu.merge_outputs(str(tsv), 6, 3)
    assert open(out).read() == "rank0\nrank1\nrank2\n"
    assert sorted(os.listdir(tmp_path)) == ["x.eventalign.diffs.6", "x.eventalign.tsv"]
