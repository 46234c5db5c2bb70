"""Runs the UNMODIFIED reference (oracle/_ref, laid out by oracle/make_ref.sh from /root/reference) as a subprocess:
`python oracle/_ref/mCaller.py ...` and `python oracle/_ref/make_bed.py ...` under the import shim.

Test infrastructure only (like everything under oracle/): bench.py's `--impl reference` arm and its `cpu_baseline` leg
time it, tests compare against it.  Nothing under mcaller_b200/ imports this module.
"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SHIM = os.path.join(REF_DIR, "shim")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "mCaller.py")) and os.path.isdir(SHIM)


def build(force=False):
    """Lay out oracle/_ref from the reference checkout (only possible where /root/reference exists)."""
    if available() and not force:
        return REF_DIR
    subprocess.check_call(["sh", os.path.join(HERE, "make_ref.sh")], stdout=subprocess.DEVNULL)
    return REF_DIR


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = SHIM
    env["PYTHONWARNINGS"] = "ignore"
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    # the reference's workers are plain forked Python processes: keep BLAS / OpenMP pools from oversubscribing the cores
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        env[v] = "1"
    return env


def scratch_dir(prefix="mcaller_ref_"):
    """A clean, dot-free working directory in memory-backed storage when there is one (reference quirks Q7/Q8: stale
    *.tmp* files pollute the output and make_bed.py cuts the output name at the first '.')."""
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.mkdtemp(prefix=prefix, dir=base)


def run_mcaller(workdir, tsv, fasta, fastq, model, threads=1, motif=None, positions=None, base="A", k=6, skip=0, qual=0,
                classifier="NN", timeout=3600):
    """One run of the reference CLI (mCaller.py:118-184).  Returns dict(wall_s, rc, stdout, diffs=<path of .diffs.<k>>).
    Wall time covers the whole process: interpreter start-up, imports, FASTQ/FASTA parsing, workers, merge."""
    out = ".".join(tsv.split(".")[:-1]) + ".diffs." + str(k)
    for f in os.listdir(workdir):                      # Q7: the reference appends to stale tmp files
        if ".tmp" in f or f == os.path.basename(out):
            os.remove(os.path.join(workdir, f))
    cmd = [sys.executable, os.path.join(REF_DIR, "mCaller.py")]
    cmd += ["-m", motif] if motif else ["-p", positions]
    cmd += ["-r", fasta, "-e", tsv, "-f", fastq, "-d", model, "-b", base, "-n", str(k), "-s", str(skip), "-q", str(qual),
            "-c", classifier, "-t", str(int(threads))]
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=workdir, env=_env(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    wall = time.perf_counter() - t0
    return dict(wall_s=wall, rc=p.returncode, stdout=p.stdout, stderr=p.stderr, diffs=out if os.path.exists(out) else None, cmd=cmd)


def run_make_bed(workdir, diffs, args=("-d", "15", "-t", "0.5"), timeout=3600):
    """One run of the reference make_bed.py (make_bed.py:166-203) on a `.diffs.<k>` file -> dict(wall_s, rc, bed=<path>)."""
    for f in os.listdir(workdir):
        if f.endswith(".bed") or f.endswith(".gff"):
            os.remove(os.path.join(workdir, f))
    cmd = [sys.executable, os.path.join(REF_DIR, "make_bed.py"), "-f", os.path.basename(diffs)] + list(args)
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=workdir, env=_env(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    wall = time.perf_counter() - t0
    beds = [f for f in os.listdir(workdir) if f.endswith(".bed") or f.endswith(".gff")]
    return dict(wall_s=wall, rc=p.returncode, stdout=p.stdout, stderr=p.stderr, bed=os.path.join(workdir, beds[0]) if beds else None)


def startup_seconds(workdir):
    """Wall time of importing what mCaller.py imports (interpreter + numpy/sklearn/scipy/pandas) -- reported next to the
    run time so a reader can see how much of a short run is fixed start-up cost."""
    code = "import sys; sys.path.insert(0, %r); import mCaller" % REF_DIR
    t0 = time.perf_counter()
    subprocess.run([sys.executable, "-c", code], cwd=workdir, env=_env(), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0


def cleanup(workdir):
    shutil.rmtree(workdir, ignore_errors=True)
