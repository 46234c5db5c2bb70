"""Streaming a host-resident eventalign buffer through the Engine: read-aligned chunks, double-buffered H2D copies on
a side stream overlapped with the kernels, results copied back per chunk.  This is the path a file-based run takes
(host memory -> PCIe -> HBM -> kernels -> rows back), and what bench.py times as `e2e`."""
import numpy as np
import torch

from ._lib import CALL_DTYPE, MC_TEXT_PAD


def plan_chunks(read_offsets, total_bytes, chunk_bytes):
    """Chunk end offsets that fall on read boundaries; read_offsets = byte offset of the first line of each read."""
    offs = np.asarray(read_offsets, dtype=np.int64)
    cuts, start = [], 0
    while start < total_bytes:
        lim = start + chunk_bytes
        if lim >= total_bytes:
            cuts.append(int(total_bytes))
            break
        j = int(np.searchsorted(offs, lim, side="right")) - 1
        end = int(offs[j]) if j >= 0 and offs[j] > start else None
        if end is None:                       # a single read larger than the chunk: take it whole
            j2 = int(np.searchsorted(offs, start, side="right"))
            end = int(offs[j2]) if j2 < len(offs) else int(total_bytes)
        cuts.append(end)
        start = end
    return cuts


class HostStreamer(object):
    def __init__(self, engine, chunk_bytes=1 << 30):
        self.eng = engine
        self.dev = engine.device
        self.chunk_bytes = int(chunk_bytes)
        cap = engine.padded_capacity(self.chunk_bytes)
        self.dbuf = [torch.empty(cap, dtype=torch.uint8, device=self.dev) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.h_calls = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _enqueue_copy(self, slot, host_buf, a, b):
        n = b - a
        if n > self.chunk_bytes:
            raise ValueError("chunk of %d bytes exceeds the staging buffers (%d)" % (n, self.chunk_bytes))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])
            self.dbuf[slot][:n].copy_(host_buf[a:b], non_blocking=True)
            self.dbuf[slot][n:n + MC_TEXT_PAD + 16].fill_(10)
            self.ready[slot].record(self.copy_stream)
        self.h2d_bytes += n

    def run(self, host_buf, cuts, fetch_calls=True):
        """host_buf: pinned uint8 CPU tensor; cuts: chunk end offsets from plan_chunks.  Returns dict of totals; rows of
        every chunk are copied to pinned host memory when fetch_calls (the D2H leg of the end-to-end path)."""
        eng = self.eng
        cur = torch.cuda.current_stream(self.dev)
        for e in self.free:
            e.record(cur)
        tot = dict(calls=0, pending_resolved=0, too_many_skips=0, multi=0, errors=0, methylated=0, records=0, lines=0, rows=0)
        bounds = [0] + list(cuts)
        if len(bounds) > 1:
            self._enqueue_copy(0, host_buf, bounds[0], bounds[1])
        pending_prev = pending_tms_prev = 0
        for i in range(len(bounds) - 1):
            slot = i & 1
            if i + 2 < len(bounds):
                self._enqueue_copy((i + 1) & 1, host_buf, bounds[i + 1], bounds[i + 2])
            cur.wait_event(self.ready[slot])
            n = bounds[i + 1] - bounds[i]
            res = eng.run_chunk(self.dbuf[slot], n)
            st = eng.count_rows(res)
            if fetch_calls and res.n_calls:
                nb = res.n_calls * CALL_DTYPE.itemsize
                if self.h_calls is None or self.h_calls.numel() < nb:
                    self.h_calls = torch.empty(int(nb * 1.5) + 4096, dtype=torch.uint8, pin_memory=True)
                self.h_calls[:nb].copy_(res.calls_dev[:nb], non_blocking=True)
                self.d2h_bytes += nb
            self.free[slot].record(cur)
            # a window left open by the previous chunk closes on this chunk's first kept line (any kept line does)
            if (pending_prev or pending_tms_prev) and res.counters["kept"] > 0:
                tot["pending_resolved"] += pending_prev
                tot["too_many_skips"] += pending_tms_prev
                pending_prev = pending_tms_prev = 0
            pending_prev += st["pending"]
            pending_tms_prev += st["pending_too_many_skips"]
            for kx in ("calls", "too_many_skips", "multi", "errors", "methylated"):
                tot[kx] += st[kx]
            tot["records"] += res.n_records
            tot["lines"] += res.counters["lines"]
            tot["rows"] += res.n_calls
        cur.synchronize()
        tot["calls"] += tot["pending_resolved"]
        tot["dropped_at_eof"] = pending_prev
        return tot
