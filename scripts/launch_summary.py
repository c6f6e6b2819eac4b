#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  Usage: launch_summary.py launches.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
t = collections.defaultdict(float); n = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ui], 1)
    name = r[ki].split("(")[0]
    t[name] += v; n[name] += 1
tot = sum(t.values())
for k in sorted(t, key=lambda k: -t[k]):
    print("%-86s launches=%4d %11.1f us %6.1f%%" % (k[:86], n[k], t[k], 100 * t[k] / tot))
