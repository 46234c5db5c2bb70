#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals from `ncu --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_source_lines.py report.ncu-rep [file-substring] [top-n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    fsub = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr = None, None
    per = {}
    for r in rows:
        if len(r) == 2 and r[0] in ("File Name", "File Path"):
            cur_file, hdr = r[1], None
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if len(r) == 2:
            continue
        if hdr is None or len(r) != len(hdr) or not cur_file or fsub not in cur_file:
            continue
        if r[0] == "":
            continue           # SASS rows belong to the source line above; the source row already aggregates them
        d = dict(zip(hdr, r))
        try:
            inst = int(d["Instructions Executed"]) if d["Instructions Executed"] not in ("-", "") else 0
            samp = int(d["# Samples"]) if d["# Samples"] not in ("-", "") else 0
            thr = int(d["Thread Instructions Executed"]) if d["Thread Instructions Executed"] not in ("-", "") else 0
        except ValueError:
            continue
        bar = int(d.get("stall_barrier", "0") or 0) if d.get("stall_barrier", "-") != "-" else 0
        lsb = int(d.get("stall_long_sb", "0") or 0) if d.get("stall_long_sb", "-") != "-" else 0
        per[(cur_file, int(r[0]))] = (inst, thr, samp, bar, lsb, r[1].strip()[:90])
    ti = sum(v[0] for v in per.values()) or 1
    ts = sum(v[2] for v in per.values()) or 1
    print("total warp-instructions %d, samples %d" % (ti, ts))
    print("%5s %7s %7s %6s %6s %6s  %s" % ("line", "inst%", "samp%", "thr/w", "bar", "longsb", "source"))
    for (f, ln), v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5d %6.2f%% %6.2f%% %6.1f %6d %6d  %s" % (ln, 100.0 * v[0] / ti, 100.0 * v[2] / ts, (v[1] / v[0]) if v[0] else 0, v[3], v[4], v[5]))


if __name__ == "__main__":
    main()
