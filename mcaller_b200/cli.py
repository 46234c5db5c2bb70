"""Command lines with the reference's flags (mCaller.py:122-141, make_bed.py:169-182) on top of the GPU path.

    python -m mcaller_b200.cli mCaller  -m GATC -r ref.fa -e x.eventalign.tsv -f x.fastq -d model.pkl -b A
    python -m mcaller_b200.cli make_bed -f x.eventalign.diffs.6 -d 15 -t 0.5

The reference's own mCaller.py / make_bed.py can be used unchanged instead; these entry points exist
so the path can be run where the reference checkout is absent.  `-t N` splits the TSV into N read-aligned byte ranges
like the reference (mCaller.py:63-68); range i is handled on GPU i mod n_gpus, one after another in this process.
`--gpus N` (extension) runs N ranks at once, one process per GPU (mcaller_b200.multigpu): rows are written per rank and
concatenated, the per-site histograms are all-reduced over NCCL, and with `--bed` rank 0 writes make_bed.py's BED / GFF
straight from the combined histogram (`--bed_min_read_depth`, `--bed_mod_threshold`, `--bed_control`, `--bed_gff`,
`--bed_ref` mirror make_bed.py's -d / -t / --control / --gff / --ref).
"""
import math
import os
import sys
from argparse import ArgumentParser


def mcaller_main(argv=None):
    from . import extract_contexts as ec, read_qual, refmark
    p = ArgumentParser(description="Classify bases as methylated or unmethylated", prog="mCaller")
    g = p.add_mutually_exclusive_group(required=True)
    g.add_argument("-p", "--positions", type=str)
    g.add_argument("-m", "--motif", type=str)
    p.add_argument("-r", "--reference", type=str, required=True)
    p.add_argument("-e", "--tsv", type=str, required=True)
    p.add_argument("-f", "--fastq", type=str, required=True)
    p.add_argument("-t", "--threads", type=int, default=1)
    p.add_argument("-b", "--base", type=str, default="A")
    p.add_argument("-n", "--num_variables", type=int, default=6)
    p.add_argument("--train", action="store_true", default=False)
    p.add_argument("--training_tsv", type=str)
    p.add_argument("-d", "--modelfile", type=str)
    p.add_argument("-s", "--skip_thresh", type=int, default=0)
    p.add_argument("-q", "--qual_thresh", type=float, default=0)
    p.add_argument("-c", "--classifier", type=str, default="NN")
    p.add_argument("--plot_training", action="store_true", default=False)
    p.add_argument("-v", "--version", action="version", version="%(prog)s v1.0 (mcaller_b200)")
    p.add_argument("--gpus", type=int, default=0, help="run N ranks, one process per GPU (0: single process, -t ranges in turn)")
    p.add_argument("--bed", action="store_true", help="with --gpus: write the make_bed.py summary from the all-reduced histogram")
    p.add_argument("--bed_min_read_depth", type=int, default=15)
    p.add_argument("--bed_mod_threshold", type=float, default=0.5)
    p.add_argument("--bed_control", action="store_true")
    p.add_argument("--bed_gff", action="store_true")
    p.add_argument("--bed_ref", type=str)
    a = p.parse_args(argv)
    if a.train or a.training_tsv:
        raise NotImplementedError("training stays with the reference (out of scope of the accelerated path)")
    if a.base == "A":
        mod = "m6A"
    elif a.base == "C":
        mod = "m5C"
    else:
        print("classification only available for A or C bases so far")
        return 0
    modelfile = a.modelfile or (os.path.dirname(os.path.realpath(sys.argv[0])) + "/model_" + a.classifier + "_" + str(a.num_variables) + "_" + mod + ".pkl")
    assert os.path.isfile(modelfile), "model file not found at " + modelfile
    base = a.motif if (a.motif and len(a.motif) == 1) else a.base
    assert a.skip_thresh < a.num_variables / 2, "too many skips with only " + str(a.num_variables) + " variables - try < half"
    assert os.path.isfile(a.fastq), "fastq file not found at " + a.fastq
    k = a.num_variables
    if a.gpus and a.gpus > 0:
        from . import multigpu
        print("%d contigs" % len(refmark.read_fasta(a.reference)))
        print("%d GPUs" % a.gpus)
        multigpu.run(dict(tsv=a.tsv, reference=a.reference, fastq=a.fastq, modelfile=modelfile, k=k, skip=a.skip_thresh, qual=a.qual_thresh,
                          base=base, motif=a.motif, positions=a.positions, bed=a.bed, bed_depth=a.bed_min_read_depth,
                          bed_thresh=a.bed_mod_threshold, bed_control=a.bed_control, bed_gff=a.bed_gff, bed_ref=a.bed_ref), a.gpus)
        print("Finished extracting signals")
        return 0
    read2qual = read_qual.extract_read_quality_device(a.fastq)        # FASTQ scanned on the GPU (same mapping as read_qual.py)
    print("%d contigs" % len(refmark.read_fasta(a.reference)))
    print("%d threads" % a.threads)
    out = ".".join(a.tsv.split(".")[:-1]) + ".diffs." + str(k)
    size = os.path.getsize(a.tsv)
    n = max(1, a.threads)
    chunk = int(math.ceil(size / float(n)))
    tmps = []
    for i in range(n):
        start = chunk * i
        tmp = out + ".tmp" + str(start)
        if os.path.exists(tmp):
            os.remove(tmp)                       # the reference appends to stale files (quirk Q7); we start clean
        ec.extract_features(a.tsv, a.reference, read2qual, k, a.skip_thresh, a.qual_thresh, modelfile, a.classifier, start,
                            endline=min(size, chunk * (i + 1)), train=False, base=base, motif=a.motif, positions_list=a.positions)
        tmps.append(tmp)
    print("Finished extracting signals")
    with open(out, "wb") as dst:                 # ranges do not overlap: concatenation in offset order == -t 1 output
        for tmp in tmps:
            with open(tmp, "rb") as src:
                dst.write(src.read())
            os.remove(tmp)
    return 0


def make_bed_main(argv=None):
    from . import make_bed as mb
    p = ArgumentParser(description="Produce bed file of methylated positions based on mCaller output")
    p.add_argument("-d", "--min_read_depth", type=int, default=15)
    p.add_argument("-t", "--mod_threshold", type=float, default=0.5)
    p.add_argument("-f", "--mCaller_file", type=str, required=True)
    p.add_argument("-p", "--positions", type=str)
    p.add_argument("--control", action="store_true")
    p.add_argument("--gff", action="store_true")
    p.add_argument("--ref", type=str)
    p.add_argument("--plot", action="store_true")
    p.add_argument("--plotsummary", action="store_true")
    p.add_argument("--plotdir", type=str, default="mCaller_position_plots")
    p.add_argument("--vo", action="store_true")
    a = p.parse_args(argv)
    assert os.path.isfile(a.mCaller_file), "file not found at " + a.mCaller_file
    out = mb.output_name(a.mCaller_file, a.positions, a.control, a.gff)
    print(a.mCaller_file)
    mb.aggregate_by_pos(a.mCaller_file, out, a.min_read_depth, a.mod_threshold, a.positions, a.control, a.vo, a.gff, a.ref, a.plot,
                        a.plotdir, a.plotsummary)
    return 0


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] not in ("mCaller", "make_bed"):
        sys.exit("usage: python -m mcaller_b200.cli {mCaller|make_bed} [flags of the reference CLI]")
    sys.exit(mcaller_main(sys.argv[2:]) if sys.argv[1] == "mCaller" else make_bed_main(sys.argv[2:]))
