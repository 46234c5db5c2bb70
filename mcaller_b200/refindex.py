"""Device-resident reference index: the per-strand target bit maps that replace the reference's
'M'-marked contig strings (extract_contexts.py:154-160, :176) on the GPU, plus the host-side site table.

Global coordinate space: contigs are laid out in name-sorted order (stage 1 binary-searches the names), each
starting at a multiple of 64 and followed by >= 64 zero bits, so a k-mer window or a context never reads into the
next contig.
"""
import ctypes as C

import numpy as np
import torch

from . import refmark
from ._lib import MC_MAXK, RefIndex


def _pack_bits(bits_u8):
    """uint8 0/1 array (length multiple of 32) -> uint32 words, bit i of word w = position 32*w+i."""
    return np.packbits(bits_u8, bitorder="little").view("<u4")


class ReferenceIndex(object):
    def __init__(self, seqs, base, motif=None, positions_file=None, k=6, device="cuda"):
        if not (1 <= k <= MC_MAXK):
            raise ValueError("num_variables must be in 1..%d" % MC_MAXK)
        self.k = k
        self.base = base
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.names = sorted(seqs, key=lambda s: s.encode())
        if len(self.names) >= 65535:
            raise ValueError("too many contigs")
        # one bytes copy of each 'M'-marked strand per contig, shared by every writer (native and Python)
        self._marked_b = {}
        # The reference marks a contig when it first meets it in the TSV (extract_contexts.py:154-160), so a bad positions
        # row only stops a run that actually touches that contig.  Here all contigs are marked up front; a contig whose
        # marking fails is kept without targets and remembered: extract_features raises if the run records a line on it.
        self.mark_errors = {}
        bases_g, lens = [], []
        off = 0
        for nm in self.names:
            seq = seqs[nm]
            try:
                fwd, rev = refmark.mark_reference(seq, base, motif=motif, positions_file=positions_file, contig=nm)
            except refmark.MarkError as e:
                if len(self.names) == 1:
                    raise
                self.mark_errors[nm] = e
                fwd = rev = seq
            self._marked_b[nm] = (fwd.encode("ascii"), rev.encode("ascii"))
            bases_g.append(off)
            lens.append(len(seq))
            off = ((off + len(seq) + 64 + 63) // 64) * 64
        self.total_bits = max(off, 64)
        n = self.total_bits
        fwd_b = np.zeros(n + 64, dtype=np.uint8)
        rev_b = np.zeros(n + 64, dtype=np.uint8)
        letters = np.full(n + 64, ord("N"), dtype=np.uint8)
        for nm, b0, ln in zip(self.names, bases_g, lens):
            f, r = self._marked_b[nm]
            fwd_b[b0:b0 + ln] = np.frombuffer(f, dtype=np.uint8) == ord("M")
            rev_b[b0:b0 + ln] = np.frombuffer(r, dtype=np.uint8) == ord("M")
            letters[b0:b0 + ln] = np.frombuffer(seqs[nm].encode("ascii"), dtype=np.uint8)
        both = fwd_b | rev_b
        cand = both.copy()
        for c in range(1, k):
            cand[:-c] |= both[c:]
        self.site_fwd_bits = fwd_b
        self.site_rev_bits = rev_b
        w_f, w_r, w_c = _pack_bits(fwd_b), _pack_bits(rev_b), _pack_bits(cand)
        pc_f = np.array([bin(x).count("1") for x in range(256)], dtype=np.uint32)
        cnt_f = pc_f[w_f.view(np.uint8)].reshape(-1, 4).sum(axis=1)
        cnt_r = pc_f[w_r.view(np.uint8)].reshape(-1, 4).sum(axis=1)
        self.n_fwd = int(cnt_f.sum())
        self.n_rev = int(cnt_r.sum())
        self.n_sites = self.n_fwd + self.n_rev
        rank_f = (np.cumsum(cnt_f) - cnt_f).astype(np.uint32)
        rank_r = (np.cumsum(cnt_r) - cnt_r + self.n_fwd).astype(np.uint32)
        # host-side site table: slot -> (contig index, position, reverse?)
        gf = np.flatnonzero(fwd_b[:n])
        gr = np.flatnonzero(rev_b[:n])
        gall = np.concatenate([gf, gr])
        starts = np.array(bases_g, dtype=np.int64)
        ci = np.searchsorted(starts, gall, side="right") - 1
        self.site_contig = ci.astype(np.int32)
        self.site_pos = (gall - starts[ci]).astype(np.int64)
        self.site_rev = np.concatenate([np.zeros(len(gf), np.uint8), np.ones(len(gr), np.uint8)])
        self.contig_base = starts
        self.contig_len = np.array(lens, dtype=np.int32)

        def dev(a):
            return torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

        name_bytes = b"".join(nm.encode() for nm in self.names)
        name_off = np.cumsum([0] + [len(nm.encode()) for nm in self.names]).astype(np.int32)
        self._t = dict(
            names=dev(np.frombuffer(name_bytes + b"\0" * 16, dtype=np.uint8).copy()), name_off=dev(name_off), base=dev(starts),
            len=dev(self.contig_len), site_fwd=dev(w_f.view(np.int32)), site_rev=dev(w_r.view(np.int32)),
            cand=dev(w_c.view(np.int32)), rank_fwd=dev(rank_f.view(np.int32)), rank_rev=dev(rank_r.view(np.int32)),
            bases=dev(letters))
        s = RefIndex()
        s.n_contigs, s.k, s.total_bits = len(self.names), k, self.total_bits
        s.d_names, s.d_name_off = self._t["names"].data_ptr(), self._t["name_off"].data_ptr()
        s.d_base, s.d_len = self._t["base"].data_ptr(), self._t["len"].data_ptr()
        s.d_site_fwd, s.d_site_rev, s.d_cand = self._t["site_fwd"].data_ptr(), self._t["site_rev"].data_ptr(), self._t["cand"].data_ptr()
        s.d_rank_fwd, s.d_rank_rev = self._t["rank_fwd"].data_ptr(), self._t["rank_rev"].data_ptr()
        s.d_bases = self._t["bases"].data_ptr()
        self.struct = s

    def ref(self):
        return C.byref(self.struct)

    def marked_bytes(self, strand):
        """The 'M'-marked copy of every contig (strand 0 = forward, 1 = reverse) as bytes, in index order; encoded once and
        shared by every writer (native and Python)."""
        return [self._marked_b[nm][strand] for nm in self.names]

    def context(self, contig_index, mpos, rev):
        """revcomp(last_ref[mpos-k+1:mpos+k], last_rev) (extract_contexts.py:194)."""
        src = self._marked_b[self.names[contig_index]][1 if rev else 0]
        return refmark.revcomp(src[mpos - self.k + 1:mpos + self.k].decode("ascii"), bool(rev))
