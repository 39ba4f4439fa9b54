#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/<round>_launches.md"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        agg[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
print("| kernel | launches | mean us | share of kernel time |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("| %s | %d | %.1f | %.3f |" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / tot))
