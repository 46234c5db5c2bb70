"""Host driver of the GPU pipeline: one `run_chunk` = stages 1-7 of include/mcaller_b200.h over one chunk of
eventalign TSV text resident in device memory.  PyTorch is used only to own device buffers and streams.

The engine never computes anything on the CPU: without a CUDA device or the built library it raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CALL_DTYPE, MC_C_COUNT, MC_TEXT_PAD, RECORD_DTYPE, check

C_LINES, C_KEPT, C_RECORDS, C_SHORT, C_UNKNOWN, C_NNN, C_BADPOS, C_LONGLINE, C_OVERFLOW = range(9)


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.McallerCudaError("mcaller_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


class ChunkResult(object):
    """Device-side results of one chunk plus the scalar counters copied to the host."""
    __slots__ = ("n_calls", "n_records", "n_segments", "counters", "calls_dev", "missing_quality", "hist_skipped", "nbytes")

    def calls(self):
        """All rows (kinds 0/1/2) as a host numpy structured array (D2H copy)."""
        if self.n_calls == 0:
            return np.zeros(0, dtype=CALL_DTYPE)
        raw = self.calls_dev[: self.n_calls * CALL_DTYPE.itemsize].cpu().numpy()
        return raw.view(CALL_DTYPE)


class Engine(object):
    def __init__(self, refindex, models=None, qual_table=None, skip_thresh=0, qual_thresh=0.0, two_models=False,
                 device="cuda", dense=None, histogram=True):
        require_cuda()
        self.L = _lib.lib()
        self.ref = refindex
        self.models = models
        self.device = torch.device(device)
        self.skip_thresh = int(skip_thresh)
        self.qual_thresh = float(qual_thresh)
        self.two_models = 1 if two_models else 0
        # the -q filter drops whole reads, so window closers can be any kept line: record every kept line then
        self.dense = bool(dense) if dense is not None else (self.qual_thresh > 0.0)
        if qual_table is None:
            qual_table = np.zeros(16, dtype=_lib.QUAL_DTYPE)
        if hasattr(qual_table, "table") and hasattr(qual_table, "size"):          # read_qual.DeviceQualityTable
            self.qual_table_size = int(qual_table.size)
            self.d_qual = qual_table.table
        else:
            self.qual_table_size = len(qual_table)
            self.d_qual = torch.from_numpy(qual_table.view(np.uint8).reshape(-1).copy()).to(self.device)
        self.histogram = histogram
        ns = max(refindex.n_sites, 1)
        self.d_depth = torch.zeros(ns, dtype=torch.int32, device=self.device)
        self.d_meth = torch.zeros(ns, dtype=torch.int32, device=self.device)
        self.d_first = torch.full((ns,), 2 ** 63 - 1, dtype=torch.int64, device=self.device)   # 'never seen'
        self.row_base = 0
        self._bufs = {}
        self.d_small = torch.zeros(64, dtype=torch.int64, device=self.device)       # counters / scalars
        self.h_small = np.zeros(64, dtype=np.uint64)
        self.launches = 0
        self.scan_events = None          # set to [] to collect (start, end) CUDA events around the scan kernel

    # ---- buffers -----------------------------------------------------------------------------------------------
    def _buf(self, name, nbytes):
        nbytes = int(max(nbytes, 256))
        t = self._bufs.get(name)
        if t is None or t.numel() < nbytes:
            t = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
            self._bufs[name] = t
        return t

    def _sptr(self):
        # launches go to torch's current stream: wrap calls in `with torch.cuda.stream(s)` to use another one
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _read_small(self, off, n):
        check(self.L.mc_read_u64(C.c_void_p(self.d_small.data_ptr() + 8 * off), n, self.h_small.ctypes.data_as(C.c_void_p), self._sptr()))
        return self.h_small[:n].copy()

    @staticmethod
    def padded_capacity(nbytes):
        return ((nbytes + 15) // 16) * 16 + MC_TEXT_PAD + 64

    def upload(self, data):
        """Host bytes -> device text tensor with the '\\n' padding the scan kernel requires."""
        n = len(data)
        t = torch.full((self.padded_capacity(n),), 10, dtype=torch.uint8, device=self.device)
        if n:
            t[:n] = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).to(self.device)
        return t

    # ---- the pipeline --------------------------------------------------------------------------------------------
    def run_chunk(self, d_text, nbytes, rec_cap=None, call_cap=None):
        """d_text: uint8 device tensor, >= padded_capacity(nbytes) long, bytes past nbytes all '\\n'."""
        L, st = self.L, self._sptr()
        if d_text.numel() < nbytes + MC_TEXT_PAD:
            raise ValueError("text tensor lacks the MC_TEXT_PAD newline padding")
        res = ChunkResult()
        res.nbytes = nbytes
        n_tiles = L.mc_num_tiles(nbytes)
        tile_tab = self._buf("tile_tab", 8 * max(n_tiles, 1))
        if rec_cap is None:
            # sparse mode: ~3 records per 29-line chunk plus the 64-slot reservation blocks of the resident warps
            slack = 256 * min(n_tiles, 8192) + 4096
            rec_cap = (nbytes // 24 + slack) if self.dense else (nbytes // 1024 + 2 * n_tiles + slack)
        rec_cap = min(rec_cap, 2 ** 32 - 2)
        while True:
            rec_a = self._buf("rec_a", 32 * rec_cap)
            self.d_small.zero_()
            if self.scan_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            check(L.mc_scan(C.c_void_p(d_text.data_ptr()), nbytes, self.ref.ref(), 1 if self.dense else 0, C.c_void_p(rec_a.data_ptr()),
                            rec_cap, C.c_void_p(tile_tab.data_ptr()), C.c_void_p(self.d_small.data_ptr()), st))
            if self.scan_events is not None:
                e1.record()
                self.scan_events.append((e0, e1))
            self.launches += 1
            cnt = self._read_small(0, MC_C_COUNT)
            if cnt[C_OVERFLOW] == 0 and cnt[C_RECORDS] <= rec_cap:
                break
            if rec_cap >= 2 ** 32 - 2:
                raise _lib.McallerCudaError("chunk produces more than 2^32 records; use smaller chunks")
            rec_cap = min(int(cnt[C_RECORDS]) + 256 * min(n_tiles, 8192) + 4096, 2 ** 32 - 2)      # demand known now: redo the scan once
        res.counters = {nm: int(cnt[i]) for i, nm in enumerate(_lib.COUNTER_NAMES)}
        reserved = int(cnt[C_RECORDS])
        res.n_records = 0
        res.n_calls = 0
        res.n_segments = 0
        res.missing_quality = 0
        res.hist_skipped = 0
        res.calls_dev = None
        if reserved == 0:
            return res
        ws = self._buf("ws", L.mc_workspace_bytes(max(reserved, n_tiles)))
        rec_b = self._buf("rec_b", 32 * reserved)
        check(L.mc_order_records(C.c_void_p(d_text.data_ptr()), nbytes, C.c_void_p(tile_tab.data_ptr()), n_tiles, C.c_void_p(rec_a.data_ptr()), rec_cap,
                                 C.c_void_p(rec_b.data_ptr()), reserved, C.c_void_p(self.d_small.data_ptr() + 8 * 21),
                                 C.c_void_p(ws.data_ptr()), st))
        self.launches += 6                           # filler resolution + 3 scan kernels + gather + record finishing
        n_rec = int(self._read_small(21, 1)[0])
        res.n_records = n_rec
        if n_rec == 0:
            return res
        seg_start = self._buf("seg_start", 4 * (n_rec + 2))
        check(L.mc_segment_reads(C.c_void_p(d_text.data_ptr()), C.c_void_p(rec_b.data_ptr()), n_rec, C.c_void_p(seg_start.data_ptr()),
                                 C.c_void_p(self.d_small.data_ptr() + 8 * 16), C.c_void_p(ws.data_ptr()), st))
        self.launches += 5                           # segmentation: flags + 3 scan kernels + starts
        n_seg = int(self._read_small(16, 1)[0])
        res.n_segments = n_seg
        seg_qual = self._buf("seg_qual", 8 * n_seg)
        seg_count = self._buf("seg_count", 4 * n_seg)
        check(L.mc_segment_quality(C.c_void_p(d_text.data_ptr()), C.c_void_p(rec_b.data_ptr()), C.c_void_p(seg_start.data_ptr()), n_seg,
                                   C.c_void_p(self.d_qual.data_ptr()), self.qual_table_size, C.c_void_p(seg_qual.data_ptr()),
                                   C.c_void_p(self.d_small.data_ptr() + 8 * 17), st))
        self.launches += 1
        if call_cap is None:
            call_cap = n_rec // 4 + 1024
        while True:
            calls = self._buf("calls", 128 * call_cap)
            check(L.mc_build_windows(C.c_void_p(rec_b.data_ptr()), n_rec, C.c_void_p(seg_start.data_ptr()), n_seg,
                                     C.c_void_p(seg_qual.data_ptr()), self.ref.ref(), self.skip_thresh, self.qual_thresh, self.two_models,
                                     C.c_void_p(calls.data_ptr()), call_cap, C.c_void_p(seg_count.data_ptr()),
                                     C.c_void_p(self.d_small.data_ptr() + 8 * 18), C.c_void_p(ws.data_ptr()), st))
            self.launches += 7                       # first-'M' pre-pass + 2 window passes + 3 scan kernels + capacity check
            v = self._read_small(17, 3)
            res.missing_quality = int(v[0])
            n_calls = int(v[1])
            if n_calls <= call_cap:
                break
            call_cap = n_calls + 16
        res.n_calls = n_calls
        res.calls_dev = calls
        if n_calls and self.models is not None:
            check(L.mc_classify(C.c_void_p(calls.data_ptr()), n_calls, self.models.array, st))
            self.launches += 1
        if n_calls and self.histogram and self.models is not None:
            check(L.mc_hist_accumulate(C.c_void_p(calls.data_ptr()), n_calls, C.c_void_p(self.d_depth.data_ptr()),
                                       C.c_void_p(self.d_meth.data_ptr()), C.c_void_p(self.d_first.data_ptr()), self.ref.n_sites,
                                       self.row_base, C.c_void_p(self.d_small.data_ptr() + 8 * 20), st))
            self.launches += 1
            self.row_base += n_calls
        return res

    def count_rows(self, res):
        """Device-side row statistics of a chunk -> dict (one tiny kernel + 48-byte read-back)."""
        self.d_small[24:31].zero_()
        if res.n_calls:
            check(self.L.mc_count_calls(C.c_void_p(res.calls_dev.data_ptr()), res.n_calls, C.c_void_p(self.d_small.data_ptr() + 8 * 24), self._sptr()))
            self.launches += 1
        v = self._read_small(24, 7)
        return dict(calls=int(v[0]), pending=int(v[1]), too_many_skips=int(v[2]), multi=int(v[3]), errors=int(v[4]), methylated=int(v[5]),
                    pending_too_many_skips=int(v[6]))

    def records(self, n_rec):
        """Ordered stage-1 records of the last chunk (host copy; test helper)."""
        return self._bufs["rec_b"][: n_rec * 32].cpu().numpy().view(RECORD_DTYPE)

    def histogram_host(self):
        return (self.d_depth.cpu().numpy().view(np.uint32), self.d_meth.cpu().numpy().view(np.uint32),
                self.d_first.cpu().numpy().view(np.uint64))

    def reset_histogram(self):
        self.d_depth.zero_()
        self.d_meth.zero_()
        self.d_first.fill_(2 ** 63 - 1)
        self.row_base = 0
