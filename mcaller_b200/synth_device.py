"""Device-side synthetic eventalign generator (bench / test tooling): wraps mc_synth_* of the C ABI.  Emits the same
bytes as mcaller_b200.synth.generate for the same SynthSpec, directly into HBM."""
import ctypes as C

import numpy as np
import torch

from . import _lib, synth
from ._lib import MC_TEXT_PAD, check


class DeviceSynth(object):
    def __init__(self, spec, refindex, meth_maps=None):
        """spec: synth.SynthSpec; refindex: ReferenceIndex built from synth.genome(spec, ci) sequences (its global
        coordinate layout and letters are reused); meth_maps: {ci: (fwd_bitmap, rev_bitmap)} methylated sites or None."""
        self.spec, self.ref = spec, refindex
        self.L = _lib.lib()
        dev = refindex.device
        self.device = dev
        names = [n for n, _ in spec.contigs]
        name_bytes = b"".join(n.encode() for n in names)
        name_off = np.cumsum([0] + [len(n.encode()) for n in names]).astype(np.int32)
        gbase = np.array([refindex.contig_base[refindex.names.index(n)] for n in names], dtype=np.int64)
        mean, sd = synth.model_table(spec)

        def devt(a):
            return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

        self._t = dict(names=devt(np.frombuffer(name_bytes + b"\0" * 8, dtype=np.uint8).copy()), name_off=devt(name_off),
                       clen=devt(np.array([l for _, l in spec.contigs], dtype=np.int32)),
                       bounds=devt(np.array(spec.read_bounds, dtype=np.int64)), gbase=devt(gbase),
                       mean=devt(mean.astype(np.int32)), sd=devt(sd.astype(np.int32)))
        s = _lib.SynthSpec()
        s.seed, s.n_contigs = spec.seed, len(names)
        s.len_min, s.len_max, s.p_skip, s.p_nnn, s.margin = spec.len_min, spec.len_max, spec.p_skip, spec.p_nnn, spec.margin
        s.d_names, s.d_name_off = self._t["names"].data_ptr(), self._t["name_off"].data_ptr()
        s.d_contig_len, s.d_read_bounds, s.d_gbase = self._t["clen"].data_ptr(), self._t["bounds"].data_ptr(), self._t["gbase"].data_ptr()
        s.d_genome = refindex._t["bases"].data_ptr()
        s.d_model_mean, s.d_model_sd = self._t["mean"].data_ptr(), self._t["sd"].data_ptr()
        s.meth = 0
        if meth_maps is not None:
            n = refindex.total_bits
            f = np.zeros(n + 64, dtype=np.uint8)
            r = np.zeros(n + 64, dtype=np.uint8)
            for ci, (mf, mr) in meth_maps.items():
                b0 = int(gbase[ci])
                f[b0:b0 + len(mf)] = mf
                r[b0:b0 + len(mr)] = mr
            self._t["mf"] = devt(np.packbits(f, bitorder="little").view(np.int32))
            self._t["mr"] = devt(np.packbits(r, bitorder="little").view(np.int32))
            s.meth, s.d_meth_fwd, s.d_meth_rev = 1, self._t["mf"].data_ptr(), self._t["mr"].data_ptr()
        self.struct = s

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def sizes(self, read0, n):
        d = torch.empty(n, dtype=torch.int64, device=self.device)
        check(self.L.mc_synth_sizes(C.byref(self.struct), read0, n, self.spec.n_reads, C.c_void_p(d.data_ptr()), self._st()))
        return d

    def generate(self, read0, n, out=None):
        """TSV text of reads [read0, read0+n) -> (uint8 device tensor with '\\n' padding, nbytes, per-read offsets)."""
        sz = self.sizes(read0, n)
        incl = torch.cumsum(sz, 0)
        total = int(incl[-1].item()) if n else 0
        offs = incl - sz
        cap = ((total + 15) // 16) * 16 + MC_TEXT_PAD + 64
        if out is None or out.numel() < cap:
            out = torch.empty(cap, dtype=torch.uint8, device=self.device)
        out[total:cap].fill_(10)
        check(self.L.mc_synth_write(C.byref(self.struct), read0, n, self.spec.n_reads, C.c_void_p(offs.data_ptr()),
                                    C.c_void_p(out.data_ptr()), self._st()))
        return out, total, offs

    def genome_check(self):
        """Device genome letters (mc_synth_genome) for comparison with the host generator."""
        n = self.ref.total_bits
        d = torch.empty(n, dtype=torch.uint8, device=self.device)
        check(self.L.mc_synth_genome(C.byref(self.struct), C.c_void_p(d.data_ptr()), n, self._st()))
        return d


def quality_table_for(spec, read_lo=0, read_hi=None):
    """{read-name prefix: mean phred} for the synthetic reads without writing a FASTQ (vectorised)."""
    read_hi = spec.n_reads if read_hi is None else read_hi
    s = synth.stream_seed(spec.seed, synth.S_QUAL)
    idx = np.arange(read_lo, read_hi, dtype=np.uint64)
    q = np.zeros(len(idx), dtype=np.float64)
    tot = np.zeros(len(idx), dtype=np.int64)
    for j in range(24):
        tot += 3 + (synth.H(s, idx, np.uint64(j)) % np.uint64(22)).astype(np.int64)
    q = tot / 24.0
    sn = synth.stream_seed(spec.seed, synth.S_NAME)
    a = synth.H(sn, idx, np.uint64(0))
    b = synth.H(sn, idx, np.uint64(1))
    keys = []
    for x, y in zip(a.tolist(), b.tolist()):
        hx = "%016x%016x" % (x, y)
        keys.append("%s-%s-%s-%s-%s" % (hx[0:8], hx[8:12], hx[12:16], hx[16:20], hx[20:32]))
    return keys, q
