"""Host-side logic that runs without a GPU: reference marking, model-key table, chunk / worker-range cutting,
read-quality ingest and lookup table, synthetic generator determinism, C-ABI surface."""
import io
import os
import random
import re

import numpy as np
import pytest

import golden_cases as gc
from mcaller_b200 import extract_contexts as ec, read_qual, refmark, synth


def test_revcomp_and_strand():
    assert ec.revcomp("GATCM") == "MGATC"
    assert ec.revcomp("GATC", False) == "GATC"
    assert ec.comp("ACGTNM") == "TGCANM"
    assert ec.strand(True) == "-" and ec.strand(False) == "+"
    with pytest.raises(KeyError):
        ec.revcomp("CAAY")          # reference has no IUPAC codes in base_comps (:11)


def test_mark_motif_semantics():
    # every `base` of the motif is marked, matching is leftmost non-overlapping (reference :39-40)
    assert refmark.mark_motif("TTGATCGATCAA", "GATC", "A") == "TTGMTCGMTCAA"
    assert refmark.mark_motif("GAGAGAG", "GAGAG", "A") == "GMGMGAG"
    assert refmark.mark_motif("CAAC", "AA", "A") == "CMMC"
    f, r = refmark.mark_reference("TTGATCAAT", "A", motif="GATC")
    assert f == "TTGMTCAAT" and r == "TTGAMCAAT"
    f, r = refmark.mark_reference("ACGT", "A", motif="A")
    assert f == "MCGT" and r == "ACGM"


def test_mark_positions(tmp_path):
    p = tmp_path / "pos.txt"
    p.write_text("c1\t1\t+\tm6A\nc1\t3\t-\tm6A\nc2\t0\t+\tm6A\n\n")
    f, r = refmark.mark_reference("CATT", "A", positions_file=str(p), contig="c1")
    assert f == "CMTT" and r == "CATM"
    with pytest.raises(refmark.MarkError):
        refmark.mark_reference("CCTT", "A", positions_file=str(p), contig="c1")


def test_base_models_table():
    t = ec.base_models("A", True)
    assert t["MG"] == "MG" and t["AG"] == "MG"
    for kx in ("MC", "MA", "MT", "MM", "AT", "AC", "AA", "AM"):
        assert t[kx] == "MH"
    g = ec.base_models("C", True)
    assert set(g.values()) == {"general"} and "MG" in g and "TM" in g
    assert ec.base_models("A", False)["MG"] == "general"


def _mk_tsv(n_reads, rnd, junk=False):
    lines, starts = [], []
    off = 0
    for r in range(n_reads):
        starts.append(off)
        for j in range(rnd.randint(1, 30)):
            ln = "ctg\t%d\tACGTAC\tread%04d_x\tt\t%d\t80.00\t1.0\t0.001\tACGTAC\t81.00\t1.5\t0.1\n" % (100 + j, r, j)
            lines.append(ln)
            off += len(ln)
            if junk and rnd.random() < 0.3:          # malformed lines inside a read must not look like read boundaries
                jl = rnd.choice(["\n", "ctg\t5\tAAA\tother\tt\n", "x y z\n"])
                lines.append(jl)
                off += len(jl)
    return "".join(lines).encode(), starts


@pytest.mark.parametrize("junk", [False, True])
def test_read_boundary_helpers(tmp_path, junk):
    rnd = random.Random(5)
    data, starts = _mk_tsv(40, rnd, junk)
    path = tmp_path / "x.tsv"
    path.write_bytes(data)
    sset = sorted(starts)
    with open(path, "rb") as fh:
        for off in [0, 1, 5, starts[3], starts[3] + 1, starts[10] - 1, len(data) - 1, len(data)]:
            want = min([s for s in sset if s >= off] + [len(data)])
            assert ec.read_boundary_after(fh, off, len(data), probe=257) == want, off
    for lim in [len(data), len(data) - 3, starts[7] + 5, starts[1] + 1, 10]:
        buf = data[:lim]
        comp_end = buf.rfind(b"\n") + 1
        good = [i for i in range(comp_end) if (i == 0 or buf[i - 1:i] == b"\n") and len(buf[i:buf.find(b"\n", i)].split()) >= 12]
        want = max([s for s in sset if good and s <= good[-1]] + [0])
        assert ec.read_boundary_before(buf) == want, lim


def test_read_quality_and_table(tmp_path):
    fq = tmp_path / "r.fastq"
    fq.write_text("@abc_Basecall_1D_template extra\nACGT\n+\n!!II\n@zz:1D_000:template\nAC\n+\n5I\n")
    q = read_qual.extract_read_quality(str(fq))
    assert set(q) == {"abc", "zz"}
    assert q["abc"] == np.mean([0, 0, 40, 40]) and q["zz"] == np.mean([20, 40])
    import gzip
    gz = tmp_path / "r.fastq.gz"
    with gzip.open(gz, "wt") as fh:
        fh.write(fq.read_text())
    assert read_qual.extract_read_quality(str(gz)) == q
    d = {"r%d" % i: float(i) / 7 for i in range(5000)}
    t = read_qual.build_quality_table(d)
    assert len(t) & (len(t) - 1) == 0 and (t["hash"] != 0).sum() == len(d)
    h, h2, ln = read_qual.fnv_pair([k.encode() for k in d])
    mask = len(t) - 1
    for i, k in enumerate(list(d)[:500]):
        s = int(h[i]) & mask
        while True:
            assert t["hash"][s] != 0
            if t["hash"][s] == h[i] and t["len"][s] == ln[i] and t["check"][s] == int(h2[i]) >> 32:
                assert t["qual"][s] == d[k]
                break
            s = (s + 1) & mask


def test_fixture_quality_matches_reference_value():
    q = read_qual.extract_read_quality(os.path.join(gc.GOLD, "masonread1", "masonread1.fastq"))
    assert q == {"26dd376e-9d82-41fc-921e-71e559c8e8d1": 7.055265349382997}


def test_synth_is_deterministic_and_wellformed():
    spec = synth.SynthSpec(seed=21, contigs=[("a", 3000), ("b", 2500)], n_reads=6, len_min=100, len_max=300, header=True)
    t1, fa1, fq1, q1 = synth.generate(spec)
    t2, fa2, fq2, q2 = synth.generate(synth.SynthSpec(seed=21, contigs=[("a", 3000), ("b", 2500)], n_reads=6, len_min=100, len_max=300, header=True))
    assert t1 == t2 and fa1 == fa2 and fq1 == fq2
    lines = t1.decode().split("\n")[1:-1]
    assert all(len(l.split("\t")) == 13 for l in lines)
    g = {"a": synth.genome(spec, 0).tobytes().decode(), "b": synth.genome(spec, 1).tobytes().decode()}
    for l in lines[:500]:
        f = l.split("\t")
        assert g[f[0]][int(f[1]):int(f[1]) + 6] == f[2]
        assert f[9] in (f[2], refmark.revcomp(f[2]), "NNNNNN")
        assert re.fullmatch(r"\d+\.\d\d", f[6])


def test_golden_inputs_are_stable(tmp_path):
    """The committed golden outputs are only meaningful if the inputs can be rebuilt bit for bit."""
    import hashlib, json
    for name in ("gatc_s0", "A_s0", "adversarial", "masonread1_p"):
        gold = json.load(open(os.path.join(gc.GOLD, name + ".json")))
        d = tmp_path / name
        d.mkdir()
        inp = gc.build_inputs(gc.CASES[name], str(d))
        assert hashlib.sha256(open(inp["tsv"], "rb").read()).hexdigest() == gold["tsv_sha256"]


def test_c_abi_exports_match_header():
    """Every function declared in include/mcaller_b200.h is exported by the built library and bound in _lib."""
    from mcaller_b200 import _lib, build
    build.build()
    hdr = open(os.path.join(os.path.dirname(gc.GOLD), "..", "include", "mcaller_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char \*)\s*\*?\s*(mc_[a-z0-9_]+)\s*\(", hdr, re.M))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for fn in declared:
        assert hasattr(lib, fn), fn
    assert declared == set(_lib.EXPORTS)
    assert lib.mc_version() == 1
    assert lib.mc_num_tiles(1) == 1 and lib.mc_num_tiles(3840) == 1 and lib.mc_num_tiles(3841) == 2
    assert lib.mc_workspace_bytes(1000) > 8000


def test_struct_sizes_match_abi():
    import ctypes as C
    from mcaller_b200 import _lib
    assert _lib.RECORD_DTYPE.itemsize == 32 and _lib.CALL_DTYPE.itemsize == 128 and _lib.QUAL_DTYPE.itemsize == 24
    from mcaller_b200 import build
    build.build()
    lib = _lib.load()
    assert lib.mc_sizeof(0) == _lib.RECORD_DTYPE.itemsize and lib.mc_sizeof(1) == _lib.CALL_DTYPE.itemsize
    assert lib.mc_sizeof(2) == C.sizeof(_lib.RefIndex) and lib.mc_sizeof(3) == C.sizeof(_lib.Model)
    assert lib.mc_sizeof(4) == _lib.QUAL_DTYPE.itemsize and lib.mc_sizeof(5) == C.sizeof(_lib.SynthSpec)
    assert lib.mc_sizeof(6) == C.sizeof(_lib.LocusEntry) == 24


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from mcaller_b200 import _lib, engine
    with pytest.raises(_lib.McallerCudaError):
        engine.require_cuda()
    with pytest.raises(_lib.McallerCudaError):
        ec.extract_features("x.tsv", "x.fa", {}, 6, 0, 0, "m.pkl", "NN", 0, endline=1, base="A", motif="GATC")


def test_iupac_expansion_matches_regex():
    """Documented extension: IUPAC motifs are expanded on both strands (all, also overlapping, matches)."""
    import re as _re
    rnd = random.Random(9)
    seq = "".join(rnd.choice("ACGT") for _ in range(20000))
    for motif, base in (("CAAYNNNNNRTAC", "A"), ("GANTC", "A"), ("CCWGG", "C"), ("RAACY", "A")):
        pat = "".join("[%s]" % refmark.IUPAC[c] for c in motif)
        want = set()
        for m in _re.finditer("(?=(%s))" % pat, seq):
            for j, ch in enumerate(motif):
                if ch == base:
                    want.add(m.start() + j)
        assert refmark.expand_iupac_sites(seq, motif, base) == sorted(want)
    f, r = refmark.mark_reference("TTCAACGGGGGATACTTGTATCCCCCGTTGAA", "A", motif="CAAYNNNNNRTAC")
    assert f.count("M") == 3 and r.count("M") == 3          # one instance per strand, three A's / T's each
