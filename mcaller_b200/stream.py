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


class TextSink(object):
    """Renders the rows of a chunk as `.diffs.<k>` text with the native writer (mc_format_rows, host threads) straight from
    the pinned row buffer and the host copy of the TSV (read names are cut from it).  Used by HostStreamer one chunk behind
    the GPU, so the formatting overlaps the next chunk's copy and kernels.  A window still open at the end of a chunk comes
    back completed as slot 0 of a later chunk (mc_carry_rows); its read name lies in the earlier chunk's text, so the sink
    keeps those few bytes when it sees the open row."""

    def __init__(self, refindex, k, base="A", with_prob=True, keep=False, max_threads=0):
        import ctypes as C
        names = refindex.names
        n = len(names)
        self._keep = [nm.encode() for nm in names] + list(refindex.marked_bytes(0)) + list(refindex.marked_bytes(1))
        self.names, self.fwd, self.rev = (C.c_char_p * n)(*self._keep[:n]), (C.c_char_p * n)(*self._keep[n:2 * n]), (C.c_char_p * n)(*self._keep[2 * n:])
        self.lens = (C.c_int64 * n)(*[len(x) for x in self._keep[n:2 * n]])
        self.n, self.k, self.with_prob = n, int(k), 1 if with_prob else 0
        self.base, self.mod = base.encode(), (b"m6A" if base == "A" else b"m" + base.encode())
        self.max_threads = int(max_threads)
        self.out = None
        self.text_bytes = 0
        self.render_seconds = 0.0             # host time spent in the native writer
        self.carry_name = None                # read name of the window currently open across a chunk edge
        self.kept = [] if keep else None      # rendered text per chunk (tests)

    def reset(self):
        self.carry_name = None
        self.text_bytes = 0
        if self.kept is not None:
            self.kept = []

    def render(self, h_calls, n_calls, host_text_ptr, max_read_len=256):
        """h_calls: pinned uint8 tensor (or numpy uint8 array) holding n_calls rows; host_text_ptr: address of the chunk's
        first byte in host memory (0 when the rows carry no read_off into a text, e.g. the row of Engine.close_carry)."""
        import ctypes as C
        from . import _lib
        if n_calls <= 0:
            return 0
        raw = h_calls.numpy() if hasattr(h_calls, "numpy") and not isinstance(h_calls, np.ndarray) else h_calls
        rows = raw[:n_calls * CALL_DTYPE.itemsize].view(CALL_DTYPE)
        name = self.carry_name
        nlen = len(name) if name is not None else 0
        cap = n_calls * (96 + 26 * (self.k + 1) + max(max_read_len, nlen)) + 4096
        if self.out is None or len(self.out) < cap:
            self.out = C.create_string_buffer(int(cap * 1.25))
        import time as _time
        t0 = _time.perf_counter()
        r = _lib.lib().mc_format_rows(C.c_void_p(rows.ctypes.data), n_calls, C.c_void_p(host_text_ptr), name, nlen, self.names, self.fwd,
                                      self.rev, self.lens, self.n, self.k, self.base, self.mod, self.with_prob, self.max_threads,
                                      self.out, len(self.out))
        self.render_seconds += _time.perf_counter() - t0
        if r < 0:
            _lib.check(int(r) if r > -100 else -1)
        if rows[0]["kind"] != _lib.MC_NONE and rows[0]["read_off"] < 0:
            self.carry_name = None            # the carried window has been written (or counted) with this chunk
        last = rows[n_calls - 1]
        if last["close_rec"] == _lib.PENDING and last["kind"] in (_lib.MC_CALL, _lib.MC_TOO_MANY_SKIPS) and last["read_off"] >= 0:
            self.carry_name = C.string_at(host_text_ptr + int(last["read_off"]), int(last["read_len"]))
        self.text_bytes += int(r)
        if self.kept is not None:
            self.kept.append(self.out.raw[:r])
        return int(r)


def last_boundary(arr, total, boundary_before, span=1 << 22):
    """Offset of the last read boundary inside arr[:total] (uint8 numpy array; 0: none) -- what
    boundary_before(bytes(arr[:total])) returns, but only a tail is searched (and copied).  The tail starts at a line start,
    so the first line seen is complete, and grows until it holds a boundary."""
    while True:
        a = max(0, total - span)
        if a > 0:
            nl = np.flatnonzero(arr[a:min(total, a + (1 << 20))] == 10)
            if len(nl) == 0:
                span *= 4
                continue
            a += int(nl[0]) + 1
        tail = arr[a:total].tobytes()
        cut = boundary_before(tail, len(tail))
        if cut > 0:
            return a + cut
        if a == 0:
            return 0
        span *= 4


class FileStreamer(object):
    """A byte range of an eventalign file -> read-aligned chunks through the Engine, pipelined: a reader thread fills a ring
    of pinned host buffers with parallel preads (cutting at the last read boundary and carrying the incomplete read into
    the next buffer), the H2D copy of chunk i+1 runs on a side stream while chunk i is in the kernels, and the caller
    renders chunk i's rows while the next copy is in flight.  `boundary_before(bytes) -> offset` finds the last read
    boundary of a buffer (extract_contexts.read_boundary_before)."""

    def __init__(self, engine, chunk_bytes, boundary_before, slots=3, readers=4):
        self.eng, self.dev = engine, engine.device
        self.chunk_bytes = int(chunk_bytes)
        self.boundary_before = boundary_before
        self.n_slots, self.readers = int(slots), int(readers)
        self.pin = [None] * self.n_slots          # pinned uint8 tensors, grown on demand
        self.dbuf = [None, None]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.h2d_bytes = 0

    def _ensure_pin(self, slot, nbytes, keep=0):
        t = self.pin[slot]
        if t is None or t.numel() < nbytes:
            nt = torch.empty(int(nbytes * 1.25) + (1 << 16), dtype=torch.uint8, pin_memory=True)
            if t is not None and keep:
                nt[:keep].copy_(t[:keep])
            self.pin[slot] = nt
        return self.pin[slot]

    def _last_boundary(self, arr, total):
        return last_boundary(arr, total, self.boundary_before)

    def _producer(self, path, lo, hi, q_free, q_ready):
        import os
        from concurrent.futures import ThreadPoolExecutor
        try:
            fd = os.open(path, os.O_RDONLY)
            try:
                with ThreadPoolExecutor(max_workers=self.readers) as pool:
                    def read_into(mv, off):
                        done = 0
                        while done < len(mv):
                            got = os.preadv(fd, [mv[done:]], off + done)
                            if got <= 0:
                                raise IOError("short read at offset %d" % (off + done))
                            done += got

                    pos, carry, slot = lo, 0, q_free.get()
                    while pos < hi or carry:
                        want = min(self.chunk_bytes, hi - pos)
                        arr = self._ensure_pin(slot, carry + want, keep=carry).numpy()
                        if want:
                            step = max(1 << 24, (want + self.readers - 1) // self.readers)
                            futs = [pool.submit(read_into, memoryview(arr[carry + a:carry + min(want, a + step)]), pos + a)
                                    for a in range(0, want, step)]
                            for f in futs:
                                f.result()
                        pos += want
                        total = carry + want
                        if pos < hi:
                            cut = self._last_boundary(arr, total)
                            if cut == 0:                     # a single read larger than the chunk: keep reading into this slot
                                carry = total
                                continue
                            nslot = q_free.get()
                            narr = self._ensure_pin(nslot, total - cut + self.chunk_bytes).numpy()
                            narr[:total - cut] = arr[cut:total]
                            q_ready.put((slot, cut))
                            slot, carry = nslot, total - cut
                        else:
                            if total:
                                q_ready.put((slot, total))
                            carry = 0
            finally:
                os.close(fd)
            q_ready.put(None)
        except BaseException as e:                          # surfaces in the consumer
            q_ready.put(e)

    def _enqueue_copy(self, dslot, slot, n):
        cap = self.eng.padded_capacity(n)
        if self.dbuf[dslot] is None or self.dbuf[dslot].numel() < cap:
            self.free[dslot].synchronize()
            self.dbuf[dslot] = torch.empty(int(cap * 1.1) + 4096, dtype=torch.uint8, device=self.dev)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[dslot])
            self.dbuf[dslot][:n].copy_(self.pin[slot][:n], non_blocking=True)
            self.dbuf[dslot][n:n + MC_TEXT_PAD + 16].fill_(10)
            self.ready[dslot].record(self.copy_stream)
        self.h2d_bytes += n

    def chunks(self, path, lo, hi):
        """Yields (ChunkResult, host text of the chunk as a uint8 numpy view, nbytes) in file order.  The view and the
        result's device buffers are valid until the next item is requested."""
        import queue
        import threading
        q_free, q_ready = queue.Queue(), queue.Queue()
        for sl in range(self.n_slots):
            q_free.put(sl)
        th = threading.Thread(target=self._producer, args=(path, lo, hi, q_free, q_ready), daemon=True)
        th.start()

        def take():
            item = q_ready.get()
            if isinstance(item, BaseException):
                raise item
            return item
        cur = torch.cuda.current_stream(self.dev)
        for e in self.free:
            e.record(cur)
        item = take()
        if item is not None:
            self._enqueue_copy(0, item[0], item[1])
        i = 0
        while item is not None:
            nxt = take()
            if nxt is not None:
                self._enqueue_copy((i + 1) & 1, nxt[0], nxt[1])
            slot, n = item
            cur.wait_event(self.ready[i & 1])
            res = self.eng.run_chunk(self.dbuf[i & 1], n)
            yield res, self.pin[slot].numpy()[:n], n
            self.free[i & 1].record(cur)
            q_free.put(slot)
            item = nxt
            i += 1
        th.join()


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
        self.h_calls = [None, None]           # pinned row buffers, alternating so one can be rendered while the other fills
        self.calls_done = [torch.cuda.Event() for _ in range(2)]
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

    def run(self, host_buf, cuts, fetch_calls=True, sink=None):
        """host_buf: pinned uint8 CPU tensor; cuts: chunk end offsets from plan_chunks.  Returns dict of totals; rows of
        every chunk are copied to pinned host memory when fetch_calls (the D2H leg of the end-to-end path) and, with a
        TextSink, rendered as `.diffs` text one chunk behind the GPU.  Windows open at a chunk end are carried on the device
        (Engine.d_carry) and appear, completed, as slot 0 of the next chunk with a kept line; the one still open after the
        last chunk is left in the carry for the caller (Engine.close_carry: next rank / end of file)."""
        eng = self.eng
        fetch_calls = fetch_calls or sink is not None
        late = None                                   # (slot, n_calls, chunk start) of the chunk whose rows are not rendered yet

        def render(item):
            slot_c, n_c, start = item
            self.calls_done[slot_c].synchronize()
            sink.render(self.h_calls[slot_c], n_c, host_buf.data_ptr() + start)
        cur = torch.cuda.current_stream(self.dev)
        for e in self.free:
            e.record(cur)
        tot = dict(calls=0, too_many_skips=0, multi=0, errors=0, methylated=0, records=0, lines=0, rows=0)
        bounds = [0] + list(cuts)
        if len(bounds) > 1:
            self._enqueue_copy(0, host_buf, bounds[0], bounds[1])
        st = None
        for i in range(len(bounds) - 1):
            slot = i & 1
            if i + 2 < len(bounds):
                self._enqueue_copy((i + 1) & 1, host_buf, bounds[i + 1], bounds[i + 2])
            cur.wait_event(self.ready[slot])
            n = bounds[i + 1] - bounds[i]
            res = eng.run_chunk(self.dbuf[slot], n)
            st = res.stats
            if fetch_calls and res.n_calls:
                nb = res.n_calls * CALL_DTYPE.itemsize
                sc = i & 1
                if self.h_calls[sc] is None or self.h_calls[sc].numel() < nb:
                    self.h_calls[sc] = torch.empty(int(nb * 1.5) + 4096, dtype=torch.uint8, pin_memory=True)
                self.h_calls[sc][:nb].copy_(res.calls_dev[:nb], non_blocking=True)
                self.calls_done[sc].record(cur)
                self.d2h_bytes += nb
                if sink is not None:
                    if late is not None:
                        render(late)                  # the previous chunk's rows, while this chunk's copy-back is in flight
                    late = (sc, res.n_calls, bounds[i])
            self.free[slot].record(cur)
            for kx in ("calls", "too_many_skips", "multi", "errors", "methylated"):
                tot[kx] += st[kx]
            tot["records"] += res.n_records
            tot["lines"] += res.counters["lines"]
            tot["rows"] += res.n_calls
        cur.synchronize()
        if sink is not None and late is not None:
            render(late)
        # the window (call or too-many-skips event) still open after the last chunk: Engine.close_carry decides its fate
        tot["open_at_end"] = (st["pending"] + st["pending_too_many_skips"]) if st is not None else 0
        return tot
