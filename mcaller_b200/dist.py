"""Multi-GPU plumbing (one process per GPU, torch.distributed): reads shard across ranks with no data-path
collective; the only exchanges are (1) one all-gather of a 2-int boundary token per rank -- the window a rank leaves
open at the end of its slice is closed by the first kept line of the next non-empty rank (reference semantics of
extract_contexts.py:179 across the byte-range split of mCaller.py:63-68) -- and (2) the all-reduce of the per-site
histograms (sum for depth / methylated counts, min for the first-seen row index)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous slice [lo, hi) of n_items owned by `rank`."""
    lo = (n_items * rank) // world
    hi = (n_items * (rank + 1)) // world
    return lo, hi


def exchange_boundaries(first_kept_contig, n_pending, device, group=None):
    """All ranks learn every rank's (contig of first kept line or -1, pending windows at slice end).
    Returns (resolved_here, closing_contig): how many of this rank's pending windows are closed by a later rank and the
    contig index that closes them (-1 = none: dropped like the reference drops the last window of the file)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.tensor([int(first_kept_contig if first_kept_contig is not None else -1), int(n_pending)], dtype=torch.int64, device=device)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    closing = -1
    for r in range(rank + 1, world):
        c = int(allv[r][0])
        if c >= 0:
            closing = c
            break
    return (int(n_pending) if closing >= 0 else 0), closing


def allreduce_histogram(depth, meth, first, group=None):
    """In-place combine of the per-site count tables across ranks (NCCL over NVLink on the GPU box, gloo in CPU tests)."""
    dist.all_reduce(depth, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(meth, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(first, op=dist.ReduceOp.MIN, group=group)


def rank_row_base(rank):
    """Row-index offset that keeps first-seen order global: rows of rank r sort after those of rank r-1."""
    return int(rank) << 40
