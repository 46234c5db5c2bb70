"""GPU parity against the golden vectors produced by the unmodified reference (tools/make_golden.py) and against the
CPU oracle, through the public drop-in API (extract_contexts.extract_features / make_bed.aggregate_by_pos, which call
the C ABI of libmcaller_b200.so).  Bar: rows, keys, contexts, strands, feature text, labels, probability text, the five
stdout counters and BED rows are all byte-identical to the reference's output."""
import json
import os

import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu

CASES = list(gc.CASES)


def _run_case(name, tmp_path, capsys, chunk_bytes=None):
    from mcaller_b200 import extract_contexts as ec, read_qual
    case = gc.CASES[name]
    gold = json.load(open(os.path.join(gc.GOLD, name + ".json")))
    inp = gc.build_inputs(case, str(tmp_path))
    read2qual = read_qual.extract_read_quality(inp["fastq"])
    base = case.get("base", "A")
    motif = case.get("motif")
    if motif and len(motif) == 1:
        base = motif
    out = ".".join(inp["tsv"].split(".")[:-1]) + ".diffs.6.tmp0"
    if os.path.exists(out):
        os.remove(out)
    old = ec.CHUNK_BYTES
    if chunk_bytes:
        ec.CHUNK_BYTES = chunk_bytes
    try:
        ec.extract_features(inp["tsv"], inp["fasta"], read2qual, 6, case.get("s", 0), case.get("q", 0.0), inp["model"], "NN", 0,
                            endline=os.path.getsize(inp["tsv"]), base=base, motif=motif, positions_list=inp.get("positions"))
    finally:
        ec.CHUNK_BYTES = old
    stdout = capsys.readouterr().out
    return gold, open(out).read(), stdout, out


@pytest.mark.parametrize("name", CASES)
def test_diffs_match_reference(name, tmp_path, capsys, cuda_lib):
    gold, mine, stdout, _ = _run_case(name, tmp_path, capsys)
    assert gold["diffs"] is not None
    g_rows, m_rows = gold["diffs"].split("\n"), mine.split("\n")
    assert len(g_rows) == len(m_rows)
    for g, m in zip(g_rows, m_rows):
        assert g == m
    c = gold["counters"]
    assert "%d observations\n" % c["observations"] in stdout
    assert "%d positions\n" % c["positions"] in stdout
    assert "%d regions with multiple methylated bases\n" % c["multi"] in stdout
    assert "%d observations with skips included\n" % c["with_skips"] in stdout
    assert "%d observations with too many skips\n" % c["too_many_skips"] in stdout


@pytest.mark.parametrize("name", ["gatc_s1", "A_s2", "adversarial", "gatc_q", "gat_q_handoff", "A_q_handoff"])
@pytest.mark.parametrize("chunk", [40000, 300000])
def test_chunked_equals_whole(name, chunk, tmp_path, capsys, cuda_lib):
    """Streaming in small read-aligned chunks (window hand-off across chunk edges) must not change a byte."""
    gold, mine, _, _ = _run_case(name, tmp_path, capsys, chunk_bytes=chunk)
    assert mine == gold["diffs"]


@pytest.mark.parametrize("name", ["gatc_s0", "gatc_s2", "masonread1_p"])
def test_bed_matches_reference(name, tmp_path, capsys, cuda_lib):
    from mcaller_b200 import make_bed as mb
    gold = json.load(open(os.path.join(gc.GOLD, name + ".json")))
    diffs = os.path.join(str(tmp_path), "syn.eventalign.diffs.6")
    with open(diffs, "w") as fh:
        fh.write(gold["diffs"])
    ref_fa = None
    if any("--ref" in b["args"] for b in gold["beds"]):
        ref_fa = gc.build_inputs(gc.CASES[name], str(tmp_path))["fasta"]
    for b in gold["beds"]:
        a = b["args"]
        out = os.path.join(str(tmp_path), "o.bed")
        pos_list = None
        if "-p" in a:
            pos_list = os.path.join(str(tmp_path), a[a.index("-p") + 1])
            gc.write_bed_positions(gold["diffs"], pos_list)
        mb.aggregate_by_pos(diffs, out, int(a[a.index("-d") + 1]), float(a[a.index("-t") + 1]), pos_list, "--control" in a, "--vo" in a,
                            "--gff" in a, ref_fa if "--ref" in a else None, False, "x", False)
        assert b["rc"] == 0 and b["bed"] is not None, a
        assert open(out).read() == b["bed"], a


def test_reference_fixture_bed(tmp_path, cuda_lib):
    """The reference's own golden BED from its own golden diffs (make_bed.py -d 1 -t 0.5)."""
    from mcaller_b200 import make_bed as mb
    fx = os.path.join(gc.GOLD, "masonread1")
    out = os.path.join(str(tmp_path), "o.bed")
    mb.aggregate_by_pos(os.path.join(fx, "masonread1.eventalign.diffs.6"), out, 1, 0.5, None, False, False, False, None, False, "x", False)
    assert open(out).read() == open(os.path.join(fx, "masonread1.methylation.summary.bed")).read()


def test_reference_fixture_features(tmp_path, capsys, cuda_lib):
    """masonread1.eventalign.diffs.6: windows, contexts, strands and the 7 feature strings are the reference's own golden
    values (its probability column predates the shipped model)."""
    _, mine, _, _ = _run_case("masonread1_p", tmp_path, capsys)
    fx = open(os.path.join(gc.GOLD, "masonread1", "masonread1.eventalign.diffs.6")).read().strip().split("\n")
    rows = mine.strip().split("\n")
    assert len(rows) == len(fx) == 9
    for a, b in zip(rows, fx):
        assert a.split("\t")[:6] == b.split("\t")[:6]


def test_train_mode_windows_match_reference_fixture(tmp_path, capsys, cuda_lib):
    """extract_features(train=True): the 44 labelled windows of the reference's own
    masonread1.eventalign.diffs.6.train come out of the GPU path (set equality, features to 1e-9: the fixture was printed
    with 12 significant digits by the python2-era reference)."""
    from mcaller_b200 import extract_contexts as ec, read_qual
    case = dict(gc.CASES["masonread1_p"], positions="test_positions.txt")
    inp = gc.build_inputs(case, str(tmp_path))
    keep = [ln for ln in open(inp["positions"]) if ln.split() and 13200 <= int(ln.split()[1]) <= 26370]
    with open(inp["positions"], "w") as fh:
        fh.writelines(keep)
    pos_label = {(f[0], int(f[1]), f[2]): f[3] for f in (ln.split() for ln in keep)}       # train_model.pos2label
    sig, ctx = ec.extract_features(inp["tsv"], inp["fasta"], read_qual.extract_read_quality(inp["fastq"]), 6, 0, 0.0, None, "NN", 0,
                                   endline=os.path.getsize(inp["tsv"]), train=True, pos_label=pos_label, base="A",
                                   positions_list=inp["positions"])
    mine = {}
    for row in open(os.path.join(str(tmp_path), "syn.eventalign.diffs.6.train.tmp0")):
        f = row.rstrip("\n").split("\t")
        mine[(int(f[2]), f[3], f[5], f[6])] = [float(x) for x in f[4].split(",")]
    fx = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(gc.GOLD, "masonread1", "masonread1.eventalign.diffs.6.train"))]
    assert len(fx) == 44 == len(mine)
    for f in fx:
        key = (int(f[1]), f[2], f[4], f[5])
        assert key in mine, key
        assert all(abs(a - b) < 1e-9 for a, b in zip([float(x) for x in f[3].split(",")], mine[key]))
    assert sum(len(v) for v in sig["general"].values()) == 44 and set(sig["general"]) == {"A", "m6A"}


def test_bed_deep_matches_reference(tmp_path, cuda_lib):
    """make_bed -p / --vo / --gff --vo on a deep-coverage `.diffs` (depth up to 520: every branch of numpy's pairwise sum;
    17-digit and scientific-notation feature text): t-test columns, probability lists and fracLow/fracUp are
    byte-identical to the reference's output (scipy 1.18.1 / numpy 2.3.5, tools/make_golden.py bed_deep)."""
    import hashlib
    from mcaller_b200 import make_bed as mb
    gold = json.load(open(os.path.join(gc.GOLD, "bed_deep.json")))
    text = gc.deep_diffs_text()
    assert hashlib.sha256(text.encode()).hexdigest() == gold["diffs_sha256"]
    diffs = os.path.join(str(tmp_path), "syn.eventalign.diffs.6")
    with open(diffs, "w") as fh:
        fh.write(text)
    for b in gold["beds"]:
        a = b["args"]
        out = os.path.join(str(tmp_path), "o.bed")
        pos_list = None
        if "-p" in a:
            pos_list = os.path.join(str(tmp_path), a[a.index("-p") + 1])
            gc.write_bed_positions(text, pos_list)
        mb.aggregate_by_pos(diffs, out, int(a[a.index("-d") + 1]), float(a[a.index("-t") + 1]), pos_list, False, "--vo" in a, "--gff" in a,
                            None, False, "x", False)
        assert b["rc"] == 0 and open(out).read() == b["bed"], a


def _cli_args(case, inp):
    return gc.cli_args(case, inp)


@pytest.mark.parametrize("name", ["gatc_s1", "A_s2", "adversarial", "gatc_q", "gat_q_handoff", "A_q_handoff", "pos_p"])
@pytest.mark.parametrize("workers", [3, 7])
def test_worker_ranges_equal_single_worker(name, workers, tmp_path, capsys, cuda_lib):
    """`-t N` (mCaller.py:63-68 byte ranges, one extract_features call per range, worker i on GPU i mod n): each worker closes
    its last open window by probing the text after its range, so the concatenated files equal the reference's -t 1 output."""
    from mcaller_b200 import cli
    case = gc.CASES[name]
    gold = json.load(open(os.path.join(gc.GOLD, name + ".json")))
    inp = gc.build_inputs(case, str(tmp_path))
    assert cli.mcaller_main(_cli_args(case, inp) + ["-t", str(workers)]) == 0
    capsys.readouterr()
    assert open(os.path.join(str(tmp_path), "syn.eventalign.diffs.6")).read() == gold["diffs"]


def _default_mode_beds(gold):
    return [b for b in gold["beds"] if "-p" not in b["args"] and "--vo" not in b["args"]]


@pytest.mark.parametrize("name", ["gatc_s0", "gatc_s2"])
@pytest.mark.parametrize("chunk", [None, 60000])
def test_bed_from_device_histogram_matches_reference(name, chunk, tmp_path, capsys, cuda_lib):
    """The fused path: per-site histogram on the device (mc_hist_accumulate, windows carried over chunk edges on the device)
    -> make_bed.aggregate_from_histogram == the reference's make_bed.py output on its own `.diffs` file, byte for byte
    (thresholds -d / -t, --control, --gff, --ref, first-seen row order), without parsing any `.diffs` text."""
    from mcaller_b200 import extract_contexts as ec, make_bed as mb, read_qual
    case = gc.CASES[name]
    gold = json.load(open(os.path.join(gc.GOLD, name + ".json")))
    inp = gc.build_inputs(case, str(tmp_path))
    old = ec.CHUNK_BYTES
    if chunk:
        ec.CHUNK_BYTES = chunk
    try:
        run = ec.RangeRun(inp["tsv"], inp["fasta"], read_qual.extract_read_quality(inp["fastq"]), 6, case.get("s", 0), 0.0, inp["model"], 0,
                          endline=os.path.getsize(inp["tsv"]), base="A", motif=case["motif"], histogram=True)
        run.stream()
    finally:
        ec.CHUNK_BYTES = old
    row = run.eng.close_carry(-1)                               # end of the file: the last open window is dropped
    assert int(row[0]["kind"]) == 3
    assert open(run.tsv_output).read() == gold["diffs"]
    depth, meth, first = run.eng.histogram_host()
    beds = _default_mode_beds(gold)
    assert len(beds) >= 2
    for b in beds:
        a = b["args"]
        out = os.path.join(str(tmp_path), "h.bed")
        mb.aggregate_from_histogram(run.ref, depth, meth, first, out, int(a[a.index("-d") + 1]), float(a[a.index("-t") + 1]),
                                    control="--control" in a, gff="--gff" in a, ref=inp["fasta"] if "--ref" in a else None,
                                    odd_rows=run.eng.odd_rows())
        assert open(out).read() == b["bed"], a
    capsys.readouterr()


def test_histogram_keeps_rows_closed_by_another_contig(tmp_path, capsys, cuda_lib):
    """Reference quirk (extract_contexts.py:216): column 1 of a row is the CLOSING line's contig.  Such rows cannot be keyed
    by site slot; the device hands them over (Engine.odd_rows) and the BED built from the histogram still equals make_bed's
    on the `.diffs` text -- the A_s0 golden has truncated reads whose last window is closed by a read on the next contig."""
    from mcaller_b200 import extract_contexts as ec, make_bed as mb, read_qual
    case = gc.CASES["A_s0"]
    gold = json.load(open(os.path.join(gc.GOLD, "A_s0.json")))
    inp = gc.build_inputs(case, str(tmp_path))
    run = ec.RangeRun(inp["tsv"], inp["fasta"], read_qual.extract_read_quality(inp["fastq"]), 6, 0, 0.0, inp["model"], 0,
                      endline=os.path.getsize(inp["tsv"]), base="A", motif="A", histogram=True)
    run.stream()
    run.eng.close_carry(-1)
    assert open(run.tsv_output).read() == gold["diffs"]
    odd = run.eng.odd_rows()
    depth, meth, first = run.eng.histogram_host()
    out = os.path.join(str(tmp_path), "h.bed")
    mb.aggregate_from_histogram(run.ref, depth, meth, first, out, 1, 0.5, odd_rows=odd)
    want = [b for b in gold["beds"] if b["args"] == ["-d", "1", "-t", "0.5"]][0]["bed"]
    assert open(out).read() == want
    assert int(depth.sum()) + len(odd) == gold["counters"]["observations"]
    capsys.readouterr()


def _multi_rank_case(name, world, tmp_path, bed_args, backend=None):
    """python -m mcaller_b200.cli mCaller ... --gpus <world> --bed in a subprocess (the ranks are spawned from it)."""
    import subprocess
    import sys
    case = gc.CASES[name]
    gold = json.load(open(os.path.join(gc.GOLD, name + ".json")))
    inp = gc.build_inputs(case, str(tmp_path))
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + os.pathsep + env.get("PYTHONPATH", "")
    if backend:
        env["MCALLER_B200_DIST_BACKEND"] = backend
    env["MCALLER_B200_CHUNK_BYTES"] = "70000"                    # several chunks per rank: carries inside and across ranks
    cmd = [sys.executable, "-m", "mcaller_b200.cli", "mCaller"] + gc.cli_args(case, inp) + ["--gpus", str(world), "--bed"] + bed_args
    p = subprocess.run(cmd, cwd=str(tmp_path), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:]
    return gold, p.stdout


@pytest.mark.parametrize("name,world", [("gatc_s0", 2), ("gatc_s2", 2), ("gatc_s2", 3)])
def test_multi_rank_product_path_on_one_gpu(name, world, tmp_path, cuda_lib):
    """The multi-GPU product path (mcaller_b200.multigpu) with several ranks sharing this GPU (gloo carries the two
    exchanges): concatenated `.diffs` == the reference's -t 1 output and the BED written from the all-reduced histogram ==
    the reference's make_bed.py output, byte for byte."""
    gold, _ = _multi_rank_case(name, world, tmp_path, ["--bed_min_read_depth", "2", "--bed_mod_threshold", "0.5"] if name == "gatc_s2"
                               else ["--bed_min_read_depth", "3", "--bed_mod_threshold", "0.3"], backend="gloo")
    assert open(os.path.join(str(tmp_path), "syn.eventalign.diffs.6")).read() == gold["diffs"]
    want = [b for b in gold["beds"] if b["args"] == (["-d", "2", "-t", "0.5"] if name == "gatc_s2" else ["-d", "3", "-t", "0.3"])][0]["bed"]
    assert open(os.path.join(str(tmp_path), "syn.methylation.summary.bed")).read() == want
    assert not [f for f in os.listdir(str(tmp_path)) if ".tmp" in f]


def test_multi_gpu_product_path_nccl(tmp_path, cuda_lib):
    """Same over NCCL with one rank per GPU (needs >= 2 GPUs: `gpurun --gpus 2`)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    gold, _ = _multi_rank_case("gatc_s2", 2, tmp_path, ["--bed_min_read_depth", "2", "--bed_mod_threshold", "0.5"])
    assert open(os.path.join(str(tmp_path), "syn.eventalign.diffs.6")).read() == gold["diffs"]
    want = [b for b in gold["beds"] if b["args"] == ["-d", "2", "-t", "0.5"]][0]["bed"]
    assert open(os.path.join(str(tmp_path), "syn.methylation.summary.bed")).read() == want


def _build_big_inputs_on_device(case, outdir):
    """Inputs of a BIG_CASES entry, the TSV written by the device generator (bit-identical to the host generator the golden
    was made with -- the sha256 of the TSV is checked by the caller)."""
    import numpy as np
    from mcaller_b200 import synth, synth_device
    from mcaller_b200.refindex import ReferenceIndex
    spec = synth.SynthSpec(**case["spec"])
    genomes = [synth.genome(spec, ci) for ci in range(len(spec.contigs))]
    seqs = {nm: genomes[ci].tobytes().decode() for ci, (nm, _) in enumerate(spec.contigs)}
    ref = ReferenceIndex(seqs, case.get("base", "A"), motif=case.get("motif"), k=6)
    meth = {}
    for ci, (nm, ln) in enumerate(spec.contigs):
        b0 = int(ref.contig_base[ref.names.index(nm)])
        meth[ci] = (synth.meth_sites(spec, ci, ref.site_fwd_bits[b0:b0 + ln]), synth.meth_sites(spec, ci, ref.site_rev_bits[b0:b0 + ln]))
    gen = synth_device.DeviceSynth(spec, ref, meth if case.get("meth") else None)
    d_text, n, _ = gen.generate(0, spec.n_reads)
    paths = {"tsv": os.path.join(outdir, "syn.eventalign.tsv"), "fasta": os.path.join(outdir, "ref.fasta"),
             "fastq": os.path.join(outdir, "syn.fastq"), "model": os.path.join(gc.GOLD, "models", case["model"])}
    d_text[:n].cpu().numpy().tofile(paths["tsv"])
    with open(paths["fasta"], "w") as fh:
        for nm, s in seqs.items():
            fh.write(">%s\n" % nm)
            fh.write("\n".join(s[j:j + 60] for j in range(0, len(s), 60)) + "\n")
    with open(paths["fastq"], "w") as fh:
        for i in range(spec.n_reads):
            qs, _ = synth.read_quality_string(spec, i)
            fh.write("@%s\n%s\n+\n%s\n" % (synth.read_name(spec, i), "A" * len(qs), qs))
    return paths


def test_thousand_read_case_matches_reference_hashes(tmp_path, capsys, cuda_lib):
    """1 200 reads / ~0.6 GB of eventalign TSV streamed in 64 MB chunks: the `.diffs.6` file, the five stdout counters and the
    BED files (make_bed drop-in on the text AND straight from the device histogram) hash to what the UNMODIFIED reference
    produced on the same bytes (tests/golden/gatc_1k_s1.json, tools/make_golden.py) -- the reference, not only the oracle, at
    a size where chunk edges, carried windows and multi-block kernels all occur."""
    import hashlib
    from mcaller_b200 import cli, extract_contexts as ec, make_bed as mb, read_qual
    case = gc.BIG_CASES["gatc_1k_s1"]
    gold = json.load(open(os.path.join(gc.GOLD, "gatc_1k_s1.json")))
    inp = _build_big_inputs_on_device(case, str(tmp_path))
    assert hashlib.sha256(open(inp["tsv"], "rb").read()).hexdigest() == gold["tsv_sha256"]
    old = ec.CHUNK_BYTES
    ec.CHUNK_BYTES = 64 << 20
    try:
        assert cli.mcaller_main(gc.cli_args(case, inp)) == 0
        stdout = capsys.readouterr().out
        diffs_path = os.path.join(str(tmp_path), "syn.eventalign.diffs.6")
        diffs = open(diffs_path, "rb").read()
        assert diffs.count(b"\n") == gold["diffs_rows"]
        assert diffs.split(b"\n")[0].decode() == gold["diffs_first_row"] and diffs.split(b"\n")[-2].decode() == gold["diffs_last_row"]
        assert hashlib.sha256(diffs).hexdigest() == gold["diffs_sha256"]
        c = gold["counters"]
        for line in ("%d observations\n" % c["observations"], "%d positions\n" % c["positions"],
                     "%d regions with multiple methylated bases\n" % c["multi"], "%d observations with skips included\n" % c["with_skips"],
                     "%d observations with too many skips\n" % c["too_many_skips"]):
            assert line in stdout
        # the fused route: histogram on the device while the rows are written
        run = ec.RangeRun(inp["tsv"], inp["fasta"], read_qual.extract_read_quality(inp["fastq"]), 6, case["s"], 0.0, inp["model"], 0,
                          endline=os.path.getsize(inp["tsv"]), base="A", motif=case["motif"], histogram=True)
        os.remove(run.tsv_output) if os.path.exists(run.tsv_output) else None
        run.stream()
        run.eng.close_carry(-1)
        assert hashlib.sha256(open(run.tsv_output, "rb").read()).hexdigest() == gold["diffs_sha256"]
        depth, meth, first = run.eng.histogram_host()
        assert int(depth.sum()) + len(run.eng.odd_rows()) == gold["diffs_rows"]
    finally:
        ec.CHUNK_BYTES = old
    for b in gold["beds"]:
        a = b["args"]
        d_, t_ = int(a[a.index("-d") + 1]), float(a[a.index("-t") + 1])
        out1, out2 = os.path.join(str(tmp_path), "t.bed"), os.path.join(str(tmp_path), "h.bed")
        mb.aggregate_by_pos(diffs_path, out1, d_, t_, None, "--control" in a, False, False, None, False, "x", False)
        mb.aggregate_from_histogram(run.ref, depth, meth, first, out2, d_, t_, control="--control" in a, odd_rows=run.eng.odd_rows())
        for out in (out1, out2):
            bed = open(out, "rb").read()
            assert bed.count(b"\n") == b["rows"] and hashlib.sha256(bed).hexdigest() == b["sha256"], (a, out)
    capsys.readouterr()


def test_bad_positions_row_only_matters_for_contigs_in_the_tsv(tmp_path, capsys, cuda_lib):
    """The reference marks a contig when the TSV first names it (extract_contexts.py:154-160): a positions row with the wrong
    base on a contig that never occurs in the TSV is harmless (output == golden), on a contig that does occur it stops the
    run (reference: print + sys.exit, here ReferenceAbort)."""
    from mcaller_b200 import extract_contexts as ec, read_qual
    case = gc.CASES["pos_p"]
    gold = json.load(open(os.path.join(gc.GOLD, "pos_p.json")))
    inp = gc.build_inputs(case, str(tmp_path))
    with open(inp["fasta"], "a") as fh:
        fh.write(">zz_unused\n" + "ACGT" * 50 + "\n")
    with open(inp["positions"], "a") as fh:
        fh.write("zz_unused\t1\t+\tm6A\n")                      # position 1 of zz_unused is 'C', not the target base
    q = read_qual.extract_read_quality(inp["fastq"])
    out = ".".join(inp["tsv"].split(".")[:-1]) + ".diffs.6.tmp0"
    ec.extract_features(inp["tsv"], inp["fasta"], q, 6, case.get("s", 0), 0.0, inp["model"], "NN", 0, endline=os.path.getsize(inp["tsv"]),
                        base="A", positions_list=inp["positions"])
    assert open(out).read() == gold["diffs"]
    os.remove(out)
    seq = [ln for ln in open(inp["fasta"]).read().split(">") if ln.startswith("p1\n")][0].split("\n", 1)[1].replace("\n", "")
    bad = next(i for i in range(20, 200) if seq[i] != "A")
    with open(inp["positions"], "a") as fh:
        fh.write("p1\t%d\t+\tm6A\n" % bad)
    with pytest.raises(ec.ReferenceAbort):
        ec.extract_features(inp["tsv"], inp["fasta"], q, 6, case.get("s", 0), 0.0, inp["model"], "NN", 0, endline=os.path.getsize(inp["tsv"]),
                            base="A", positions_list=inp["positions"])
    capsys.readouterr()
