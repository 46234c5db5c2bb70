#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time and share of the last step
(everything from the last k_scan launch on).  usage: python tools/launch_summary.py launches.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, seq = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        seq.append((d["Kernel Name"][:72], v))
idx = [i for i, (n, v) in enumerate(seq) if "k_scan(" in n or "k_scan<" in n]
step = seq[idx[-1]:] if idx else seq
tot = sum(v for n, v in step)
agg = {}
for n, v in step:
    a = agg.setdefault(n, [0.0, 0])
    a[0] += v
    a[1] += 1
for n, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-74s n=%2d %10.1f us %5.1f%%" % (n, c, v / 1e3, 100 * v / tot))
print("step total %.3f ms (ncu: serialised, cold cache)" % (tot / 1e6))
