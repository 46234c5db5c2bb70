"""Deterministic synthetic `nanopolish eventalign` data (SURVEY.md section 8d recipe).

Everything is derived from a counter-based 64-bit hash (splitmix64 finaliser) with
integer arithmetic only, so the numpy implementation here and the CUDA generator in
csrc/synth.cu (used by bench.py to fill HBM without touching the host) emit the
same bytes for the same (seed, read index).  This module is host-side tooling for
tests, golden-vector generation and the bounded CPU-baseline sample; it is not on
the product path.

Row layout mirrors the reference fixture testdata/masonread1.eventalign.tsv
(13 tab-separated columns, `%.2f` currents; reference extract_contexts.py:139).
"""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_G = np.uint64(0x9E3779B97F4A7C15)
_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)

# stream ids (must match csrc/synth_common.cuh)
S_REF, S_MODEL, S_LEN, S_START, S_STRAND, S_NAME, S_QUAL, S_POS, S_EV, S_EV2, S_E0, S_METH, S_NEXTRA = range(13)

K_MODEL = 6  # nanopolish k-mer length
READ_SUFFIX = "_Basecall_1D_template"
MAX_EVENTS = 20

# current offsets (centi-pA) added to events whose k-mer covers a methylated site,
# indexed by column c = site - position (0..5); chosen so the shipped MLP separates
# the two classes reasonably (see tools/make_golden.py --tune-offsets)
METH_OFFSETS = (150, -250, 400, -350, 300, 120)


def _u64(x):
    return np.asarray(x).astype(np.uint64)


def mix(x):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = _u64(x) + _G
        z = (z ^ (z >> np.uint64(30))) * _C1
        z = (z ^ (z >> np.uint64(27))) * _C2
        return z ^ (z >> np.uint64(31))


def stream_seed(seed, stream):
    return mix(np.uint64(seed) * np.uint64(64) + np.uint64(stream))


def H(sseed, a, b=0):
    with np.errstate(over="ignore"):
        return mix(mix(_u64(sseed) + _u64(a)) + _u64(b))


def read_stream_seed(seed, stream, i):
    """Per-read stream seed: stream_seed + i * golden (wrapping)."""
    with np.errstate(over="ignore"):
        return stream_seed(seed, stream) + np.uint64(i) * _G


_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNM", b"TGCANM"):
    _COMP[_a] = _b


class SynthSpec(object):
    """Parameters of one synthetic data set."""

    def __init__(self, seed=0, contigs=(("ecoli", 4600000),), n_reads=100, len_min=1000, len_max=3000,
                 p_skip=1966, p_nnn=2621, header=False, meth=False):
        self.seed = int(seed)
        self.contigs = [(str(n), int(l)) for n, l in contigs]
        self.n_reads = int(n_reads)
        self.len_min = int(len_min)
        self.len_max = int(len_max)
        self.p_skip = int(p_skip)      # out of 65536 (3 %)
        self.p_nnn = int(p_nnn)        # out of 65536 (4 %)
        self.header = bool(header)
        self.meth = bool(meth)
        self.margin = 16
        tot = sum(l for _, l in self.contigs)
        cum = np.cumsum([0] + [l for _, l in self.contigs])
        # reads per contig proportional to length, contiguous read-index ranges
        self.read_bounds = [(self.n_reads * int(c)) // tot for c in cum]
        for _, l in self.contigs:
            assert l >= self.len_max + 2 * self.margin + K_MODEL + 8, "contig too short for len_max"


def genome(spec, ci):
    """Contig `ci` as a uint8 array of ACGT (i.i.d. uniform)."""
    n = spec.contigs[ci][1]
    s = stream_seed(spec.seed, S_REF)
    h = H(s, np.uint64(ci), np.arange(n, dtype=np.uint64))
    return _BASES[((h >> np.uint64(13)) & np.uint64(3)).astype(np.int64)]


def model_table(spec):
    """(mean_centi[4096], sd_centi[4096]); k-mer index = base-4 number, A=0..T=3, first base most significant."""
    s = stream_seed(spec.seed, S_MODEL)
    h = H(s, np.arange(4096, dtype=np.uint64), 0)
    mean = 5500 + (h % np.uint64(7001)).astype(np.int64)
    sd = 120 + ((h >> np.uint64(32)) % np.uint64(181)).astype(np.int64)
    return mean, sd


def read_meta(spec, i):
    """(contig index, start, length, reverse?, e0) of read i."""
    ci = 0
    while not (spec.read_bounds[ci] <= i < spec.read_bounds[ci + 1]):
        ci += 1
    clen = spec.contigs[ci][1]
    n_c = spec.read_bounds[ci + 1] - spec.read_bounds[ci]
    i_c = i - spec.read_bounds[ci]
    length = spec.len_min + int(H(stream_seed(spec.seed, S_LEN), i) % np.uint64(spec.len_max - spec.len_min + 1))
    span = clen - 2 * spec.margin - spec.len_max - K_MODEL
    stride = max(span // n_c, 1)
    start = spec.margin + min((i_c * span) // n_c + int(H(stream_seed(spec.seed, S_START), i) % np.uint64(stride)), span)
    rev = int(H(stream_seed(spec.seed, S_STRAND), i) & np.uint64(1))
    e0 = int(H(stream_seed(spec.seed, S_E0), i) % np.uint64(50))
    return ci, start, length, rev, e0


def read_name(spec, i):
    s = stream_seed(spec.seed, S_NAME)
    a = int(H(s, i, 0))
    b = int(H(s, i, 1))
    hx = "%016x%016x" % (a, b)
    return "%s-%s-%s-%s-%s%s" % (hx[0:8], hx[8:12], hx[12:16], hx[16:20], hx[20:32], READ_SUFFIX)


def read_quality_string(spec, i, n=24):
    """FASTQ quality characters (phred 3..24) for read i; the sequence line is n 'A's."""
    s = stream_seed(spec.seed, S_QUAL)
    q = 3 + (H(s, i, np.arange(n, dtype=np.uint64)) % np.uint64(22)).astype(np.int64)
    return "".join(chr(33 + int(x)) for x in q), q


def _fmt2(v):
    """`%.2f` of a centi-unit integer (sign aware)."""
    sgn = "-" if v < 0 else ""
    v = abs(int(v))
    return "%s%d.%02d" % (sgn, v // 100, v % 100)


def meth_sites(spec, ci, bitmap):
    """Subset of marked sites that carry the synthetic methylation signal (hash bit)."""
    pos = np.flatnonzero(bitmap)
    s = stream_seed(spec.seed, S_METH)
    keep = (H(s, np.uint64(ci), pos.astype(np.uint64)) & np.uint64(1)).astype(bool)
    out = np.zeros_like(bitmap)
    out[pos[keep]] = 1
    return out


def read_lines(spec, i, genomes, model, site_maps=None):
    """All TSV lines (list of str, no newline) of read i.

    genomes: list of uint8 arrays per contig; model: (mean, sd);
    site_maps: optional {ci: (fwd_bitmap, rev_bitmap)} uint8 arrays of *methylated* sites.
    """
    ci, start, length, rev, e0 = read_meta(spec, i)
    g = genomes[ci]
    cname = spec.contigs[ci][0]
    name = read_name(spec, i)
    mean, sd = model
    pos = np.arange(start, start + length, dtype=np.uint64)
    hp = H(read_stream_seed(spec.seed, S_POS, i), pos, 0)
    skipped = (hp & np.uint64(0xFFFF)) < np.uint64(spec.p_skip)
    nev = np.ones(length, dtype=np.int64)
    alive = np.ones(length, dtype=bool)
    for j in range(6):
        fld = (hp >> np.uint64(16 + 8 * j)) & np.uint64(0xFF)
        alive &= fld < np.uint64(123)
        nev += alive
    extra = (H(read_stream_seed(spec.seed, S_NEXTRA, i), pos, 0) % np.uint64(14)).astype(np.int64)
    nev = np.where(nev == 7, 7 + extra, nev)
    nev = np.where(skipped, 0, nev)
    total = int(nev.sum())
    first = np.cumsum(nev) - nev
    s_ev = read_stream_seed(spec.seed, S_EV, i)
    s_ev2 = read_stream_seed(spec.seed, S_EV2, i)
    out = []
    for pi in range(length):
        n = int(nev[pi])
        if n == 0:
            continue
        p = start + pi
        kb = g[p:p + K_MODEL]
        ref_kmer = kb.tobytes().decode()
        if rev:
            mk = _COMP[kb][::-1]
        else:
            mk = kb
        model_kmer = mk.tobytes().decode()
        kidx = 0
        for c in mk:
            kidx = kidx * 4 + b"ACGT".index(int(c))
        mm, ms = int(mean[kidx]), int(sd[kidx])
        off = 0
        if site_maps is not None:
            bm = site_maps[ci][1 if rev else 0]
            for c in range(K_MODEL):
                if bm[p + c]:
                    off += METH_OFFSETS[c]
        hv = H(s_ev, np.uint64(p), np.arange(n, dtype=np.uint64))
        hv2 = H(s_ev2, np.uint64(p), np.arange(n, dtype=np.uint64))
        for j in range(n):
            v = int(hv[j])
            v2 = int(hv2[j])
            idx = int(first[pi]) + j
            e = e0 + (total - 1 - idx if rev else idx)
            ssum = ((v >> 16) & 0xFFF) + ((v >> 28) & 0xFFF) + ((v >> 40) & 0xFFF) + ((v >> 52) & 0xFFF) - 8190
            noise = (abs(ssum) * 240) // 2365
            if ssum < 0:
                noise = -noise
            evc = mm + noise + off
            stdv = 500 + v2 % 2500
            dur = 100 + (v2 >> 16) % 900
            if (v & 0xFFFF) < spec.p_nnn:
                tail = "NNNNNN\t0.00\t0.00\tinf"
            else:
                z = (abs(evc - mm) * 100) // ms
                if evc < mm:
                    z = -z
                tail = "%s\t%s\t%s\t%s" % (model_kmer, _fmt2(mm), _fmt2(ms), _fmt2(z))
            out.append("%s\t%d\t%s\t%s\tt\t%d\t%s\t%d.%03d\t0.%05d\t%s" % (
                cname, p, ref_kmer, name, e, _fmt2(evc), stdv // 1000, stdv % 1000, dur, tail))
    return out


HEADER = "contig\tposition\treference_kmer\tread_name\tstrand\tevent_index\tevent_level_mean\tevent_stdv\tevent_length\tmodel_kmer\tmodel_mean\tmodel_stdv\tstandardized_level"


def generate(spec, site_maps=None, reads=None):
    """Return (tsv_bytes, fasta_text, fastq_text, qual_means dict) for the whole data set."""
    genomes = [genome(spec, ci) for ci in range(len(spec.contigs))]
    model = model_table(spec)
    lines = [HEADER] if spec.header else []
    fq = []
    quals = {}
    for i in (range(spec.n_reads) if reads is None else reads):
        lines.extend(read_lines(spec, i, genomes, model, site_maps))
        qs, q = read_quality_string(spec, i)
        nm = read_name(spec, i)
        fq.append("@%s\n%s\n+\n%s\n" % (nm, "A" * len(qs), qs))
        quals[nm] = float(np.mean(q))
    fasta = []
    for ci, (nm, ln) in enumerate(spec.contigs):
        s = genomes[ci].tobytes().decode()
        fasta.append(">%s\n" % nm)
        fasta.extend(s[j:j + 60] + "\n" for j in range(0, ln, 60))
    tsv = ("\n".join(lines) + "\n").encode() if lines else b""
    return tsv, "".join(fasta), "".join(fq), quals
