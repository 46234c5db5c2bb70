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


class DeviceQualityTable(object):
    """The read2qual mapping built and kept on the GPU (mc_fastq_index + mc_fastq_quality).  Accepted wherever the
    reference passes its `read2qual` dict into extract_features."""

    def __init__(self, table, size, n_records, bad_headers):
        self.table, self.size, self.n_records, self.bad_headers = table, size, n_records, bad_headers

    def to_host(self):
        """Entries as a numpy QUAL_DTYPE array (tests / debugging)."""
        return self.table.cpu().numpy().view(QUAL_DTYPE)


def extract_read_quality_device(fastqfi, device="cuda"):
    """GPU version of extract_read_quality: FASTQ(.gz) bytes -> DeviceQualityTable."""
    import ctypes as C
    import torch
    from . import _lib, engine
    engine.require_cuda()
    L = _lib.lib()
    opener = gzip.open if fastqfi.find(".gz") != -1 else open
    with opener(fastqfi, "rb") as fh:
        data = fh.read()
    n = len(data)
    dev = torch.device(device)
    cap = ((n + 15) // 16) * 16 + 16
    d_text = torch.full((cap,), 10, dtype=torch.uint8, device=dev)
    if n:
        d_text[:n] = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).to(dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    nt = max(int(L.mc_fastq_tiles(n)), 1)
    d_cnt = torch.zeros(nt, dtype=torch.int32, device=dev)
    d_off = torch.zeros(nt, dtype=torch.int32, device=dev)
    d_ws = torch.empty(int(L.mc_workspace_bytes(nt)), dtype=torch.uint8, device=dev)
    d_small = torch.zeros(8, dtype=torch.int64, device=dev)
    # first call sizes the line index (capacity 1 is enough to count), second call fills it
    d_lines = torch.zeros(1, dtype=torch.int64, device=dev)
    _lib.check(L.mc_fastq_index(C.c_void_p(d_text.data_ptr()), n, C.c_void_p(d_cnt.data_ptr()), C.c_void_p(d_off.data_ptr()),
                                C.c_void_p(d_lines.data_ptr()), 1, C.c_void_p(d_small.data_ptr()), C.c_void_p(d_ws.data_ptr()), st))
    n_nl = int(d_small[0].item())
    n_lines = n_nl + (1 if n and data[-1:] != b"\n" else 0)          # a last line without newline still counts
    d_lines = torch.zeros(n_nl + 2, dtype=torch.int64, device=dev)
    _lib.check(L.mc_fastq_index(C.c_void_p(d_text.data_ptr()), n, C.c_void_p(d_cnt.data_ptr()), C.c_void_p(d_off.data_ptr()),
                                C.c_void_p(d_lines.data_ptr()), n_nl + 2, C.c_void_p(d_small.data_ptr()), C.c_void_p(d_ws.data_ptr()), st))
    if n_lines > n_nl:                                              # no final newline: sentinel one past the virtual one
        d_lines[n_lines] = n + 1
    n_rec = n_lines // 4
    size = 16
    while size < 2 * n_rec + 2:
        size *= 2
    d_table = torch.zeros(size * QUAL_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_owner = torch.zeros(size, dtype=torch.int32, device=dev)
    d_mean = torch.zeros(max(n_rec, 1), dtype=torch.float64, device=dev)
    _lib.check(L.mc_fastq_quality(C.c_void_p(d_text.data_ptr()), n, C.c_void_p(d_lines.data_ptr()), n_lines, C.c_void_p(d_table.data_ptr()),
                                  size, C.c_void_p(d_owner.data_ptr()), C.c_void_p(d_mean.data_ptr()), C.c_void_p(d_small.data_ptr() + 8), st))
    stats = d_small.cpu().numpy()
    if stats[3]:
        raise RuntimeError("quality table overflow")
    return DeviceQualityTable(d_table, size, int(stats[1]), int(stats[2]))
