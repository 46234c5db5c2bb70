#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) under the
import shim in tools/ref_shim on deterministic inputs.  Run in the build container only
(the reference is not available on the GPU box); outputs are committed under tests/golden/.

    python tools/make_golden.py            # all cases
    python tools/make_golden.py gatc_s0    # one case

Each case becomes tests/golden/<case>.json holding: how to rebuild the inputs
(SynthSpec + post-processing name, or the fixture name), the reference command line, a sha256
of the TSV, and the reference's outputs: the `.diffs.<k>` rows, the five stdout counters and
the make_bed BED rows.
"""
import gzip
import hashlib
import json
import os
import pickle
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mcaller_b200 import synth  # noqa: E402
import golden_cases  # noqa: E402  (tests/golden_cases.py: shared case table + input builders)

REF = "/root/reference"
SHIM = os.path.join(HERE, "ref_shim")
GOLD = os.path.join(ROOT, "tests", "golden")


def run(cmd, cwd):
    env = dict(os.environ)
    env["PYTHONPATH"] = SHIM
    env["PYTHONWARNINGS"] = "ignore"
    p = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    return p.returncode, p.stdout, p.stderr


def counters(stdout):
    out = {}
    for key, pat in (("observations", r"^(\d+) observations$"), ("positions", r"^(\d+) positions$"),
                     ("multi", r"^(\d+) regions with multiple methylated bases$"),
                     ("with_skips", r"^(\d+) observations with skips included$"),
                     ("too_many_skips", r"^(\d+) observations with too many skips$")):
        m = re.search(pat, stdout, re.M)
        out[key] = int(m.group(1)) if m else None
    return out


def make_case(name):
    case = golden_cases.CASES[name]
    tmp = tempfile.mkdtemp(prefix="gold_")
    try:
        inputs = golden_cases.build_inputs(case, tmp, models_dir=os.path.join(GOLD, "models"))
        cmd = [sys.executable, os.path.join(REF, "mCaller.py")] + golden_cases.cli_args(case, inputs)
        rc, so, se = run(cmd, tmp)
        diffs_path = os.path.join(tmp, "syn.eventalign.diffs.%d" % case.get("k", 6))
        rec = {"case": name, "cmd": golden_cases.cli_args(case, {k: os.path.basename(v) for k, v in inputs.items()}),
               "rc": rc, "tsv_sha256": hashlib.sha256(open(inputs["tsv"], "rb").read()).hexdigest(),
               "counters": counters(so)}
        if not os.path.exists(diffs_path):
            rec["error"] = (so[-2000:] + "\n" + se[-2000:])
            rec["diffs"] = None
        else:
            rec["diffs"] = open(diffs_path).read()
            for bed_args in case.get("beds", [["-d", "1", "-t", "0.5"]]):
                for f in os.listdir(tmp):
                    if f.endswith(".bed") or f.endswith(".gff"):
                        os.remove(os.path.join(tmp, f))
                if "-p" in bed_args:
                    golden_cases.write_bed_positions(rec["diffs"], os.path.join(tmp, bed_args[bed_args.index("-p") + 1]))
                rc2, so2, se2 = run([sys.executable, os.path.join(REF, "make_bed.py"), "-f", "syn.eventalign.diffs.%d" % case.get("k", 6)] + bed_args, tmp)
                beds = [f for f in os.listdir(tmp) if f.endswith(".bed") or f.endswith(".gff")]
                rec.setdefault("beds", []).append({"args": bed_args, "rc": rc2,
                                                   "bed": open(os.path.join(tmp, beds[0])).read() if beds else None})
        with open(os.path.join(GOLD, name + ".json"), "w") as fh:
            json.dump(rec, fh, indent=1, sort_keys=True)
        nrows = rec["diffs"].count("\n") if rec["diffs"] else -1
        print("%-18s rc=%d rows=%d counters=%s" % (name, rc, nrows, rec["counters"]))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def make_big_case(name):
    """Reference outputs of a BIG_CASES entry as counts + sha256 (the rows themselves would be tens of MB)."""
    case = golden_cases.BIG_CASES[name]
    tmp = tempfile.mkdtemp(prefix="gold_")
    try:
        inputs = golden_cases.build_inputs(case, tmp, models_dir=os.path.join(GOLD, "models"))
        cmd = [sys.executable, os.path.join(REF, "mCaller.py")] + golden_cases.cli_args(case, inputs)
        rc, so, se = run(cmd, tmp)
        diffs_path = os.path.join(tmp, "syn.eventalign.diffs.6")
        diffs = open(diffs_path, "rb").read()
        rows = diffs.split(b"\n")
        rec = {"case": name, "rc": rc, "tsv_sha256": hashlib.sha256(open(inputs["tsv"], "rb").read()).hexdigest(),
               "tsv_bytes": os.path.getsize(inputs["tsv"]), "counters": counters(so), "diffs_rows": diffs.count(b"\n"),
               "diffs_sha256": hashlib.sha256(diffs).hexdigest(), "diffs_first_row": rows[0].decode(), "diffs_last_row": rows[-2].decode(),
               "beds": []}
        for bed_args in case["beds"]:
            for f in os.listdir(tmp):
                if f.endswith(".bed") or f.endswith(".gff"):
                    os.remove(os.path.join(tmp, f))
            rc2, so2, se2 = run([sys.executable, os.path.join(REF, "make_bed.py"), "-f", "syn.eventalign.diffs.6"] + bed_args, tmp)
            beds = [f for f in os.listdir(tmp) if f.endswith(".bed") or f.endswith(".gff")]
            bed = open(os.path.join(tmp, beds[0]), "rb").read()
            rec["beds"].append({"args": bed_args, "rc": rc2, "rows": bed.count(b"\n"), "sha256": hashlib.sha256(bed).hexdigest(),
                                "first_row": bed.split(b"\n")[0].decode()})
        with open(os.path.join(GOLD, name + ".json"), "w") as fh:
            json.dump(rec, fh, indent=1, sort_keys=True)
        print("%-18s rc=%d rows=%d counters=%s beds=%s" % (name, rc, rec["diffs_rows"], rec["counters"], [b["rows"] for b in rec["beds"]]))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def make_bed_deep():
    """make_bed-only golden on a synthetic deep-coverage `.diffs.6` (golden_cases.deep_diffs_text)."""
    tmp = tempfile.mkdtemp(prefix="gold_")
    try:
        text = golden_cases.deep_diffs_text()
        with open(os.path.join(tmp, "syn.eventalign.diffs.6"), "w") as fh:
            fh.write(text)
        rec = {"case": "bed_deep", "diffs_sha256": hashlib.sha256(text.encode()).hexdigest(), "beds": []}
        for bed_args in golden_cases.BED_DEEP_VARIANTS:
            for f in os.listdir(tmp):
                if f.endswith(".bed") or f.endswith(".gff"):
                    os.remove(os.path.join(tmp, f))
            if "-p" in bed_args:
                golden_cases.write_bed_positions(text, os.path.join(tmp, bed_args[bed_args.index("-p") + 1]))
            rc2, so2, se2 = run([sys.executable, os.path.join(REF, "make_bed.py"), "-f", "syn.eventalign.diffs.6"] + bed_args, tmp)
            beds = [f for f in os.listdir(tmp) if f.endswith(".bed") or f.endswith(".gff")]
            rec["beds"].append({"args": bed_args, "rc": rc2, "bed": open(os.path.join(tmp, beds[0])).read() if beds else None})
            print("bed_deep %s rc=%d rows=%d" % (bed_args, rc2, rec["beds"][-1]["bed"].count("\n") if beds else -1))
        with open(os.path.join(GOLD, "bed_deep.json"), "w") as fh:
            json.dump(rec, fh, indent=1, sort_keys=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def copy_fixtures():
    """Reference-owned data fixtures needed on the GPU box (not source code)."""
    os.makedirs(os.path.join(GOLD, "models"), exist_ok=True)
    os.makedirs(os.path.join(GOLD, "masonread1"), exist_ok=True)
    for f in ("r95_twobase_model_NN_6_m6A.pkl", "r94_model_NN_6_m6A.pkl", "CAAYNNNNNRTAC_model_6_m6A.pkl",
              "CRAANNNNNNNTGC_model_6_m6A.pkl"):
        shutil.copy(os.path.join(REF, f), os.path.join(GOLD, "models", f))
    # bare re-pickle of the motif model: lets the reference run it unmodified
    env_py = ("import pickle,sys; m=pickle.load(open(sys.argv[1],'rb'),encoding='latin');"
              "pickle.dump(m['general'],open(sys.argv[2],'wb'))")
    subprocess.check_call([sys.executable, "-W", "ignore", "-c", env_py,
                           os.path.join(REF, "CAAYNNNNNRTAC_model_6_m6A.pkl"),
                           os.path.join(GOLD, "models", "CAAY_bare_model_6_m6A.pkl")],
                          env=dict(os.environ, PYTHONPATH=SHIM))
    td = os.path.join(REF, "testdata")
    with open(os.path.join(td, "masonread1.eventalign.tsv"), "rb") as src, \
            gzip.GzipFile(os.path.join(GOLD, "masonread1", "masonread1.eventalign.tsv.gz"), "wb", mtime=0) as dst:
        shutil.copyfileobj(src, dst)
    for f in ("masonread1.fastq", "test_positions_m6A.txt", "test_positions_A.txt", "test_positions.txt",
              "masonread1.eventalign.diffs.6", "masonread1.eventalign.diffs.6.train",
              "masonread1.methylation.summary.bed", "pb_ecoli_polished_assembly.fasta.fai"):
        shutil.copy(os.path.join(td, f), os.path.join(GOLD, "masonread1", f))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    if not os.path.exists(os.path.join(GOLD, "models", "CAAY_bare_model_6_m6A.pkl")):
        copy_fixtures()
    names = sys.argv[1:] or (list(golden_cases.CASES) + ["bed_deep"] + list(golden_cases.BIG_CASES))
    for nm in names:
        if nm == "bed_deep":
            make_bed_deep()
        elif nm in golden_cases.BIG_CASES:
            make_big_case(nm)
        else:
            make_case(nm)
