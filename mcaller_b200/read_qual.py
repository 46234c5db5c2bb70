"""Read-quality ingest (reference read_qual.py:6-19) and the device lookup table used by stage 4.

`extract_read_quality` keeps the reference's contract: {id.split(':')[0].split('_')[0]: mean phred} for FASTQ or
FASTQ.gz.  The FASTQ scan itself is host code in this round (SURVEY.md section 8f ranks its GPU version next)."""
import gzip

import numpy as np

from ._lib import QUAL_DTYPE

_FNV_BASIS = np.uint64(14695981039346656037)
_FNV_BASIS2 = np.uint64(0x84222325CBF29CE4)
_FNV_PRIME = np.uint64(1099511628211)


def extract_read_quality(fastqfi):
    read2qual = {}
    opener = gzip.open if fastqfi.find(".gz") != -1 else open
    with opener(fastqfi, "rb") as fh:
        while True:
            head = fh.readline()
            if not head:
                break
            if not head.strip():
                continue
            fh.readline()
            fh.readline()
            q = fh.readline().rstrip(b"\r\n")
            toks = head[1:].split()
            rid = (toks[0] if toks else b"").decode().split(":")[0].split("_")[0]
            read2qual[rid] = np.mean(np.frombuffer(q, dtype=np.uint8).astype(np.int64) - 33) if len(q) else np.float64("nan")
    return read2qual


def fnv_pair(keys):
    """Vectorised FNV-1a (two bases) over a list of byte strings -> (h, h2, lengths)."""
    n = len(keys)
    lens = np.fromiter((len(k) for k in keys), dtype=np.int64, count=n)
    L = int(lens.max()) if n else 0
    mat = np.zeros((n, max(L, 1)), dtype=np.uint8)
    for i, k in enumerate(keys):
        mat[i, :len(k)] = np.frombuffer(k, dtype=np.uint8)
    h = np.full(n, _FNV_BASIS, dtype=np.uint64)
    h2 = np.full(n, _FNV_BASIS2, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for j in range(L):
            act = lens > j
            c = mat[act, j].astype(np.uint64)
            h[act] = (h[act] ^ c) * _FNV_PRIME
            h2[act] = (h2[act] ^ c) * _FNV_PRIME
    h[h == 0] = np.uint64(1)
    return h, h2, lens


def build_quality_table(read2qual):
    """Open-addressing table (linear probing, load <= 0.5) of mc_qual_entry records."""
    keys = [k.encode() for k in read2qual]
    vals = np.fromiter((float(read2qual[k]) for k in read2qual), dtype=np.float64, count=len(keys))
    n = len(keys)
    size = 16
    while size < 2 * n + 2:
        size *= 2
    table = np.zeros(size, dtype=QUAL_DTYPE)
    if n == 0:
        return table
    h, h2, lens = fnv_pair(keys)
    mask = np.uint64(size - 1)
    slot = (h & mask).astype(np.int64)
    pending = np.arange(n)
    while len(pending):
        s = slot[pending]
        free = table["hash"][s] == 0
        # among pending items whose slot is free, the first per slot wins this round
        cand = pending[free]
        _, first = np.unique(slot[cand], return_index=True)
        win = cand[first]
        table["hash"][slot[win]] = h[win]
        table["check"][slot[win]] = (h2[win] >> np.uint64(32)).astype(np.uint32)
        table["len"][slot[win]] = lens[win].astype(np.uint32)
        table["qual"][slot[win]] = vals[win]
        placed = np.zeros(n, dtype=bool)
        placed[win] = True
        pending = pending[~placed[pending]]
        slot[pending] = (slot[pending] + 1) & (size - 1)
    return table
