"""Reference marking: which reference positions are call targets, per strand.

Host-side, one-off per contig.  Mirrors the semantics of the
reference helpers so that the same positions end up 'M' on the same strand:

* motif mode  (`-m`)  -- reference extract_contexts.py:33-41 + :60-64: every occurrence of
  `base` inside the motif becomes 'M'; matching is Python `str.replace` (leftmost,
  non-overlapping, literal).  The reverse strand uses revcomp(motif) with the complement base,
  in forward coordinates.
* positions mode (`-p`) -- extract_contexts.py:45-56 + :65-69: '+' rows mark the forward copy,
  '-' rows the reverse copy; the base found must be `base` / its complement (or already 'M').

The marked strings are kept on the host (they are what the 2k-1 context column is cut from,
extract_contexts.py:194); the device only gets bit maps derived from them.
"""
import numpy as np

_COMP = {"A": "T", "C": "G", "T": "A", "G": "C", "N": "N", "M": "M"}
_COMP_TABLE = str.maketrans("ACGTNM", "TGCANM")

IUPAC = {
    "A": "A", "C": "C", "G": "G", "T": "T", "R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC",
    "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT",
}


class MarkError(ValueError):
    """A `-p` position does not carry the expected base (reference prints and exits, :52-54)."""


def comp(seq):
    """Complement; raises KeyError on characters outside ACGTNM like the reference (:11-15)."""
    for ch in seq:
        if ch not in _COMP:
            raise KeyError(ch)
    return seq.translate(_COMP_TABLE)


def revcomp(seq, rev=True):
    """Reverse complement when `rev`, identity otherwise (reference :18-22)."""
    if not rev:
        return seq
    return comp(seq)[::-1]


def strand(rev):
    return "-" if rev else "+"


def read_fasta(path):
    """Ordered {contig id: upper-case sequence}; id = first whitespace token of the header."""
    seqs = {}
    name, parts = None, []
    with open(path, "r") as fh:
        for ln in fh:
            if ln.startswith(">"):
                if name is not None:
                    seqs[name] = "".join(parts).upper()
                toks = ln[1:].split()
                name, parts = (toks[0] if toks else ""), []
            elif name is not None:
                parts.append(ln.strip())
    if name is not None:
        seqs[name] = "".join(parts).upper()
    return seqs


def mark_motif(seq, motif, base):
    """Literal motif replace with every `base` in the motif turned into 'M' (reference :33-41)."""
    marked_motif = motif.replace(base, "M")
    return seq.replace(motif, marked_motif)


def mark_positions(seq, positions, base):
    """0-based positions -> 'M'; raises MarkError where the reference would quit (:45-56)."""
    buf = bytearray(seq, "ascii")
    want = ord(base)
    for p in positions:
        if buf[p] == want or buf[p] == 77:
            buf[p] = 77
        else:
            raise MarkError("Base %d does not correspond to methylated base - check reference positions are 0-based" % p)
    return buf.decode("ascii")


def read_positions(path, contig):
    """(forward positions, reverse positions) listed for `contig` (reference :66-67)."""
    fwd, rev = [], []
    with open(path, "r") as fh:
        for row in fh.read().split("\n"):
            f = row.split()
            if len(f) > 1 and f[0] == contig:
                if f[2] == "+":
                    fwd.append(int(f[1]))
                elif f[2] == "-":
                    rev.append(int(f[1]))
    return fwd, rev


def expand_iupac_sites(seq, motif, base):
    """Documented extension: all (possibly overlapping) IUPAC matches of `motif`;
    returns sorted positions of every `base` letter of the motif inside each match."""
    arr = np.frombuffer(seq.encode("ascii"), dtype=np.uint8)
    n, m = len(arr), len(motif)
    if n < m:
        return []
    ok = np.ones(n - m + 1, dtype=bool)
    for j, ch in enumerate(motif):
        allowed = IUPAC[ch]
        col = arr[j:n - m + 1 + j]
        hit = np.zeros(n - m + 1, dtype=bool)
        for a in allowed:
            hit |= col == ord(a)
        ok &= hit
    starts = np.flatnonzero(ok)
    offs = [j for j, ch in enumerate(motif) if ch == base]
    out = set()
    for o in offs:
        out.update((starts + o).tolist())
    return sorted(out)


def _is_plain(motif):
    return all(ch in "ACGT" for ch in motif)


def mark_reference(seq, base, motif=None, positions_file=None, contig=None):
    """(fwd_marked, rev_marked) for one contig (reference methylate_references :60-73)."""
    if not positions_file and motif:
        if _is_plain(motif):
            fwd = mark_motif(seq, motif, base)
            rev = mark_motif(seq, revcomp(motif), _COMP[base])
        else:
            # extension: IUPAC motif (the reference raises KeyError in revcomp for these)
            rc = "".join({"R": "Y", "Y": "R", "K": "M", "M": "K", "B": "V", "V": "B", "D": "H", "H": "D"}.get(c, _COMP.get(c, c))
                         for c in motif[::-1])
            fwd = mark_positions(seq, expand_iupac_sites(seq, motif, base), base)
            rev = mark_positions(seq, expand_iupac_sites(seq, rc, _COMP[base]), _COMP[base])
        return fwd, rev
    if positions_file:
        fpos, rpos = read_positions(positions_file, contig)
        return mark_positions(seq, fpos, base), mark_positions(seq, rpos, _COMP[base])
    raise ValueError("no motifs or positions specified")


def site_bitmap(marked):
    """uint8 0/1 array, 1 where the marked string holds 'M'."""
    return (np.frombuffer(marked.encode("ascii"), dtype=np.uint8) == 77).astype(np.uint8)
