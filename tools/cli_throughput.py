#!/usr/bin/env python
"""Throughput of the drop-in extract_features() on a FILE (page cache -> pinned buffers -> H2D -> kernels -> .diffs text on
disk), i.e. what `mCaller.py -t 1` spends in its hot path -- and, with --gpus N, of the multi-GPU product path
(`python -m mcaller_b200.cli mCaller ... --gpus N --bed`: N ranks, rows per rank, NCCL histogram, BED from the histogram).
usage: python tools/cli_throughput.py [--reads N] [--gpus N] [--skip S]"""
import argparse
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=8000)
    ap.add_argument("--dir", default=None)
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--skip", type=int, default=0)
    a = ap.parse_args()
    import io
    import contextlib
    import torch
    from mcaller_b200 import extract_contexts as ec, synth, synth_device
    from mcaller_b200.refindex import ReferenceIndex
    spec = synth.SynthSpec(seed=0, contigs=[("ecoli", 4600000)], n_reads=a.reads, len_min=1000, len_max=3000)
    seqs = {"ecoli": synth.genome(spec, 0).tobytes().decode()}
    ref = ReferenceIndex(seqs, "A", motif="GATC", k=6)
    meth = {0: (synth.meth_sites(spec, 0, ref.site_fwd_bits[:4600000]), synth.meth_sites(spec, 0, ref.site_rev_bits[:4600000]))}
    gen = synth_device.DeviceSynth(spec, ref, meth)
    d_text, n, _ = gen.generate(0, a.reads)
    keys, q = synth_device.quality_table_for(spec, 0, a.reads)
    quals = dict(zip(keys, q.tolist()))
    d = a.dir or tempfile.mkdtemp(prefix="mc_cli_")
    os.makedirs(d, exist_ok=True)
    tsv, fasta = os.path.join(d, "syn.eventalign.tsv"), os.path.join(d, "ref.fasta")
    with open(tsv, "wb") as fh:
        step = 1 << 28
        for o in range(0, n, step):
            fh.write(d_text[o:min(n, o + step)].cpu().numpy().tobytes())
    with open(fasta, "w") as fh:
        fh.write(">ecoli\n%s\n" % seqs["ecoli"])
    del d_text, gen
    torch.cuda.empty_cache()
    model = os.path.join(ROOT, "tests", "golden", "models", "r95_twobase_model_NN_6_m6A.pkl")
    if a.gpus:
        import subprocess
        fastq = os.path.join(d, "syn.fastq")
        with open(fastq, "w") as fh:
            for i in range(a.reads):
                qs, _ = synth.read_quality_string(spec, i)
                fh.write("@%s\n%s\n+\n%s\n" % (synth.read_name(spec, i), "A" * len(qs), qs))
        cmd = [sys.executable, "-m", "mcaller_b200.cli", "mCaller", "-m", "GATC", "-r", fasta, "-e", tsv, "-f", fastq, "-d", model, "-b", "A",
               "-s", str(a.skip), "--gpus", str(a.gpus), "--bed", "--bed_min_read_depth", "15", "--bed_mod_threshold", "0.5"]
        env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
        for rep in range(2):                   # first pass warms the page cache
            t0 = time.perf_counter()
            p = subprocess.run(cmd, cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            dt = time.perf_counter() - t0
            if p.returncode:
                sys.exit(p.stdout[-3000:])
        rows = sum(1 for _ in open(os.path.join(d, "syn.eventalign.diffs.6"), "rb"))
        bed = sum(1 for _ in open(os.path.join(d, "syn.methylation.summary.bed"), "rb"))
        stream = [ln for ln in p.stdout.split("\n") if ln.startswith("rank ")]
        print("mCaller --gpus %d --bed on a %.2f GB file (%d reads, -s %d): %.2f s wall for the whole command (process start, CUDA / NCCL "
              "init, FASTQ, reference marking included), %d rows -> %.0f calls/s, %d BED rows" % (a.gpus, n / 1e9, a.reads, a.skip, dt, rows, rows / dt, bed))
        print("\n".join(stream))
        return
    out = os.path.join(d, "syn.eventalign.diffs.6.tmp0")
    res = []
    for rep in range(2):                       # first pass warms the page cache, buffers and the CUDA context
        if os.path.exists(out):
            os.remove(out)
        buf = io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(buf):
            ec.extract_features(tsv, fasta, quals, 6, 0, 0.0, model, "NN", 0, endline=n, base="A", motif="GATC")
        dt = time.perf_counter() - t0
        calls = int(buf.getvalue().split("\n")[1].split()[0])
        res.append((dt, calls))
    dt, calls = res[-1]
    print("extract_features on a %.2f GB file (%d reads): %.2f s, %d calls -> %.0f calls/s, %.2f GB/s (first pass %.2f s); output %.1f MB"
          % (n / 1e9, a.reads, dt, calls, calls / dt, n / 1e9 / dt, res[0][0], os.path.getsize(out) / 1e6))


if __name__ == "__main__":
    main()
