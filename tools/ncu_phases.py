#!/usr/bin/env python
"""Share of warp-instructions and stall samples per phase of k_scan, from an ncu --set full report taken with
--import-source on.  Phases are delimited by marker comments in mcaller_b200/csrc/scan.cu.
usage: python tools/ncu_phases.py report.ncu-rep"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARKERS = [("helpers: TMA/mbarrier", "// ---- TMA / mbarrier wrappers"), ("A  gt20 + IDP.4A packing (pack32 is shared with the newline map)", "// ---- byte classification"),
           ("D  generic field walk / next_bit", "// ---- generic field walk"), ("D  line_fields / nth_bit / load8 / parse_pos8", "// ---- field location"),
           ("D  classify_line (+ global slow path)", "enum { ST_KEPT"), ("Z  kernel prologue / staging lambda", "__global__ void __launch_bounds__(THREADS"),
           ("A  newline test (3 instructions per word)", "// byte == 0x0a in three instructions per word: a LOP3"),
           ("Z  kernel prologue / staging lambda", "// warp-uniform state"),
           ("B0 prefetch + run claim + mbarrier wait", "// ---- 0. prefetch"), ("B1 newline map loop (loads, stores)", "// ---- 1. newline map"),
           ("C  line starts per lane", "// ---- 2. line list: lane owns"), ("Q  quiet test (rounds over the lanes' own lines)", "// ---- 2a. quick look"),
           ("C  line list prefix + slots (full parse only)", "// ---- 2b. line list"), ("C  line list write (full parse only)", "// list entries"),
           ("D  full parse: field-start map + per-line driver", "// ---- 3. full parse"), ("E  ballots / emit decision", "// ---- 4. which lines"),
           ("E  record write / chunk table", "// ---- 5. raw records"), ("Z  counters / tail", "// ---- counters")]


def main():
    rep = sys.argv[1]
    src = open(os.path.join(ROOT, "mcaller_b200", "csrc", "scan.cu")).read().split("\n")
    starts = []
    for name, mark in MARKERS:
        for i, ln in enumerate(src):
            if mark in ln:
                starts.append((i + 1, name))
                break
    starts.sort()
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr, agg = None, None, {}
    for r in rows:
        if len(r) == 2 and r[0] in ("File Name", "File Path"):
            cur, hdr = r[1], None
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if len(r) == 2 or hdr is None or len(r) != len(hdr) or r[0] == "":
            continue
        d = dict(zip(hdr, r))
        try:
            inst, samp, ln = int(d["Instructions Executed"]), int(d["# Samples"]), int(r[0])
        except ValueError:
            continue
        f = os.path.basename(cur)
        if f == "scan.cu":
            g = "scan.cu (before first marker)"
            for s0, name in starts:
                if ln >= s0:
                    g = name
        else:
            g = "other: " + f
        a = agg.setdefault(g, [0, 0])
        a[0] += inst
        a[1] += samp
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    for g, (i, s) in sorted(agg.items()):
        print("%-52s inst %5.1f%%  samples %5.1f%%" % (g, 100.0 * i / ti, 100.0 * s / ts))
    print("total warp-instructions %d" % ti)


if __name__ == "__main__":
    main()
