"""Drop-in replacement for the hot part of the reference's make_bed.py: `aggregate_by_pos` (make_bed.py:67-164) and
`check_thresh` (:21-28) with the same signatures.  The per-position depth / methylated counts are built on the GPU
(mc_diffs_aggregate: tab-split of the `.diffs.<k>` rows + integer atomics in a hash table); the host only formats
the BED/GFF rows of the surviving loci in first-seen order.  Reporting variants that need per-read value lists
(-p positions with t-tests, --vo, --plot*) are out of scope and raise NotImplementedError.
"""
import ctypes as C
import os

import numpy as np

LOCUS_DTYPE = np.dtype([("hash", "<u8"), ("first_off", "<u8"), ("depth", "<u4"), ("meth", "<u4")])


def check_thresh(locus_list, mod_thresh, depth_thresh, control):
    """reference make_bed.py:21-28 on a 0/1 list (kept for API compatibility)."""
    if len(locus_list) >= depth_thresh:
        frac = np.mean(locus_list)
        if not control and frac >= mod_thresh:
            return True
        if control and frac < mod_thresh:
            return True
        return False


def _check_counts(depth, meth, mod_thresh, depth_thresh, control):
    if depth >= depth_thresh:
        frac = np.float64(meth) / np.float64(depth)        # np.mean of a 0/1 list
        return (not control and frac >= mod_thresh) or (control and frac < mod_thresh)
    return False


def count_loci(meth_fi):
    """GPU pass over the file -> list of (chrom, pos, context, strand, depth, meth) in first-seen order."""
    import torch
    from . import _lib, engine
    engine.require_cuda()
    L = _lib.lib()
    data = open(meth_fi, "rb").read()
    n = len(data)
    if n == 0:
        return []
    dev = torch.device("cuda")
    d_text = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).to(dev)
    size = 1024
    while size < 2 * (n // 48 + 16):
        size *= 2
    init = np.zeros(size, dtype=LOCUS_DTYPE)
    init["first_off"] = np.uint64(0xFFFFFFFFFFFFFFFF)
    d_table = torch.from_numpy(init.view(np.uint8).reshape(-1)).to(dev)
    d_cnt = torch.zeros(8, dtype=torch.int64, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(L.mc_diffs_aggregate(C.c_void_p(d_text.data_ptr()), n, C.c_void_p(d_table.data_ptr()), size, C.c_void_p(d_cnt.data_ptr()), st))
    cnt = d_cnt.cpu().numpy()
    if cnt[1]:
        raise ValueError("%d rows of %s do not have 7 or 8 tab-separated fields" % (cnt[1], meth_fi))
    if cnt[3]:
        raise RuntimeError("locus table overflow")
    table = d_table.cpu().numpy().view(LOCUS_DTYPE)
    table = table[table["hash"] != 0]
    table = table[np.argsort(table["first_off"], kind="stable")]
    out = []
    for e in table:
        off = int(e["first_off"])
        end = data.find(b"\n", off)
        f = data[off:end if end >= 0 else n].split(b"\t")
        out.append((f[0].decode(), f[2].decode(), f[3].decode(), f[5].decode(), int(e["depth"]), int(e["meth"])))
    return out


def aggregate_by_pos(meth_fi, aggfi, depth_thresh, mod_thresh, pos_list, control, verbose_results, gff, ref, plot, plotdir, plotsummary):
    """reference make_bed.py:67; default mode plus --control, --gff (without --vo) and --ref."""
    if pos_list or verbose_results or plot or plotsummary:
        raise NotImplementedError("make_bed -p/--vo/--plot/--plotsummary are outside the accelerated path")
    loci = count_loci(meth_fi)
    contexts = None
    if ref:
        from . import refmark
        seqs = {}
        name, parts = None, []
        for ln in open(ref):                         # make_bed.py:36-40 keeps the case of the file
            if ln.startswith(">"):
                if name is not None:
                    seqs[name] = "".join(parts)
                t = ln[1:].split()
                name, parts = (t[0] if t else ""), []
            elif name is not None:
                parts.append(ln.strip())
        if name is not None:
            seqs[name] = "".join(parts)
        contexts = {}
        for chrom, pos, ctx, strand, _, _ in loci:
            if chrom in seqs:
                cx = seqs[chrom][int(pos) - 20:int(pos) + 21].upper()
                contexts[(chrom, pos, ctx, strand)] = refmark.revcomp(cx) if strand == "-" else cx
    count = 0
    with open(aggfi, "w") as outfi:
        for chrom, pos, ctx, strand, depth, meth in loci:
            if not _check_counts(depth, meth, mod_thresh, depth_thresh, control):
                continue
            count += 1
            frac = np.float64(meth) / np.float64(depth)
            cx = contexts[(chrom, pos, ctx, strand)] if contexts is not None else ctx
            nextpos = str(int(pos) + 1)
            if gff:
                deets = "coverage=" + str(depth) + ";context=" + cx + ";IPDRatio=5;frac=" + str(frac)
                outfi.write("\t".join([chrom, "kinModCall", "m6A", nextpos, nextpos, "10", strand, ".", deets]) + "\n")
            else:
                outfi.write("\t".join([chrom, pos, nextpos, ctx, str(frac), strand, str(depth)]) + "\n")
    if not control:
        print(count, "methylated loci found with min depth", depth_thresh, "reads")
    else:
        print(count, "unmethylated loci found with min depth", depth_thresh, "reads")
    return count


def output_name(mCaller_file, positions=None, control=False, gff=False):
    """File naming of reference make_bed.py:184-194 (first '.' of the whole path, quirk Q8)."""
    stem = mCaller_file.split(".")[0]
    if positions:
        out = stem + ".methylation.positions.summary"
    elif not control:
        out = stem + ".methylation.summary"
    else:
        out = stem + ".methylation.control.summary"
    return out + (".gff" if gff else ".bed")
