"""Pins the CPU oracle (oracle/mcaller_oracle.c) to the reference: its own shipped fixtures and the outputs the
unmodified reference produced on deterministic inputs (tests/golden/*.json, tools/make_golden.py)."""
import json
import os

import pytest

import golden_cases as gc

FAST = ["masonread1_p", "masonread1_gatc", "masonread1_gatc_s2", "gatc_s0", "gatc_s2", "A_s0", "A_s2", "gaa_s1", "pos_p", "gatc_q", "gat_q_handoff", "A_q_handoff",
        "adversarial", "bare_r94", "bare_caay_p", "cg_c", "gatc_s1", "gaa_s2", "rf_gatc", "lr_gatc", "nbc_gatc"]


def _oracle_run(orc, name, tmp_path):
    case = gc.CASES[name]
    inp = gc.build_inputs(case, str(tmp_path))
    tsv = open(inp["tsv"], "rb").read()
    base = case.get("base", "A")
    motif = case.get("motif")
    if motif and len(motif) == 1:
        base = motif
    return orc.extract(tsv, inp["fasta"], orc.read_fastq_quals(inp["fastq"]), k=6, skip_thresh=case.get("s", 0),
                       qual_thresh=case.get("q", 0.0), model=orc.load_pickle(inp["model"]), base=base, motif=motif,
                       positions=inp.get("positions"))


@pytest.mark.parametrize("name", FAST)
def test_oracle_reproduces_reference_output(name, tmp_path, oracle):
    gold = json.load(open(os.path.join(gc.GOLD, name + ".json")))
    r = _oracle_run(oracle, name, tmp_path)
    assert "".join(x + "\n" for x in r["rows"]) == gold["diffs"]
    assert r["counters"] == gold["counters"]


@pytest.mark.parametrize("name", ["gatc_s0", "gatc_s2", "masonread1_p", "cg_c"])
def test_oracle_bed(name, oracle):
    gold = json.load(open(os.path.join(gc.GOLD, name + ".json")))
    for b in gold["beds"]:
        a = b["args"]
        if "--ref" in a:
            continue            # reporting variant of the same counts: covered by the GPU drop-in test
        d, t = int(a[a.index("-d") + 1]), float(a[a.index("-t") + 1])
        if "--gff" in a or "--vo" in a or "-p" in a:
            pos_lines = None
            if "-p" in a:
                import tempfile
                with tempfile.TemporaryDirectory() as td:
                    gc.write_bed_positions(gold["diffs"], os.path.join(td, "p.txt"))
                    pos_lines = open(os.path.join(td, "p.txt")).read()
            rows = oracle.aggregate_variants(gold["diffs"], d, t, "--control" in a, pos_lines, "--vo" in a, "--gff" in a)
        else:
            rows = oracle.aggregate(gold["diffs"], d, t, "--control" in a)
        assert "".join(x + "\n" for x in rows) == b["bed"], a


def test_oracle_bed_deep(oracle):
    """make_bed -p / --vo / --gff on the deep-coverage synthetic `.diffs` (tools/make_golden.py bed_deep)."""
    import hashlib
    import tempfile
    gold = json.load(open(os.path.join(gc.GOLD, "bed_deep.json")))
    text = gc.deep_diffs_text()
    assert hashlib.sha256(text.encode()).hexdigest() == gold["diffs_sha256"]
    with tempfile.TemporaryDirectory() as td:
        gc.write_bed_positions(text, os.path.join(td, "p.txt"))
        pos_text = open(os.path.join(td, "p.txt")).read()
    for b in gold["beds"]:
        a = b["args"]
        rows = oracle.aggregate_variants(text, int(a[a.index("-d") + 1]), float(a[a.index("-t") + 1]), False,
                                         pos_text if "-p" in a else None, "--vo" in a, "--gff" in a)
        assert "".join(x + "\n" for x in rows) == b["bed"], a


def test_reference_fixture_diffs(tmp_path, oracle):
    """testdata/masonread1.eventalign.diffs.6: keys, contexts, strands and the 7 feature strings (probabilities in that
    file predate the shipped model)."""
    r = _oracle_run(oracle, "masonread1_p", tmp_path)
    fx = open(os.path.join(gc.GOLD, "masonread1", "masonread1.eventalign.diffs.6")).read().strip().split("\n")
    assert len(r["rows"]) == 9
    for a, b in zip(r["rows"], fx):
        assert a.split("\t")[:6] == b.split("\t")[:6]


def test_reference_fixture_train_windows(tmp_path, oracle):
    """testdata/masonread1.eventalign.diffs.6.train (44 windows, older 6-column format, 12 significant digits):
    the same windows come out of the oracle run on the labelled positions file (set equality, features to 1e-9)."""
    case = dict(gc.CASES["masonread1_p"], positions="test_positions.txt")
    inp = gc.build_inputs(case, str(tmp_path))
    # the positions file lists sites outside the reconstructed FASTA span; the reference quits on those (:52-54)
    keep = [ln for ln in open(inp["positions"]) if ln.split() and 13200 <= int(ln.split()[1]) <= 26370]
    with open(inp["positions"], "w") as fh:
        fh.writelines(keep)
    r = oracle.extract(open(inp["tsv"], "rb").read(), inp["fasta"], oracle.read_fastq_quals(inp["fastq"]), k=6, model=None,
                       base="A", positions=inp["positions"])
    mine = {}
    for c in r["calls"]:
        mine[(c["mpos"], c["context"], "-" if c["rev"] else "+")] = c["feat"]
    fx = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(gc.GOLD, "masonread1", "masonread1.eventalign.diffs.6.train"))]
    assert len(fx) == 44
    hit = 0
    for f in fx:
        key = (int(f[1]), f[2], f[4])
        assert key in mine, key
        vals = [float(x) for x in f[3].split(",")]
        assert all(abs(a - b) < 1e-9 for a, b in zip(vals, mine[key][:7]))
        hit += 1
    assert hit == 44


def test_reference_fixture_bed(oracle):
    fx = os.path.join(gc.GOLD, "masonread1")
    rows = oracle.aggregate(open(os.path.join(fx, "masonread1.eventalign.diffs.6")).read(), 1, 0.5, False)
    assert "".join(x + "\n" for x in rows) == open(os.path.join(fx, "masonread1.methylation.summary.bed")).read()
