#!/usr/bin/env python3
"""Summarise an ncu report's stall samples per CUDA source line.
Usage: ncu_lines.py report.ncu-rep [kernel_index] [top]   (needs -lineinfo and --import-source on)"""
import csv, subprocess, sys, io
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
def num(x):
    try: return int(x)
    except ValueError: return 0
kernels = []; hdr = None; fname = None; path = None; line = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": path = r[1]; continue
    if r[0] == "Function Name": fname = r[1]; continue
    if r[0] == "Line No":
        hdr = r
        if not kernels or kernels[-1]["name"] != fname or path in kernels[-1]["paths"]:
            kernels.append({"name": fname, "paths": set(), "lines": {}})
        kernels[-1]["paths"].add(path)
        si = hdr.index("# Samples")
        stall_cols = [i for i in range(len(hdr)) if hdr[i].startswith("stall_") and "Not Issued" not in hdr[i]]
        continue
    if hdr is None: continue
    if r[0] != "":                       # a CUDA source line; its SASS rows follow
        line = (path.split("/")[-1], num(r[0]), r[1].strip())
        kernels[-1]["lines"].setdefault(line, [0, {}])
        continue
    ent = kernels[-1]["lines"][line]
    ent[0] += num(r[si])
    for i in stall_cols:
        ent[1][hdr[i][6:]] = ent[1].get(hdr[i][6:], 0) + num(r[i])
for k in kernels:
    k["lines"] = [(f, ln, src, v[0], list(v[1].items())) for (f, ln, src), v in k["lines"].items()]
print("kernels:", [(i, k["name"][:60], sum(l[3] for l in k["lines"])) for i, k in enumerate(kernels)])
k = kernels[which]
tot = sum(l[3] for l in k["lines"])
print("total samples", tot)
for f, ln, src, n, stall in sorted(k["lines"], key=lambda l: -l[3])[:top]:
    st = sorted(stall, key=lambda x: -x[1])[:2]
    print("%6d %5.1f%%  %s:%d  %s   %s" % (n, 100.0 * n / max(tot, 1), f, ln, src[:100], st))
if len(sys.argv) > 5:
    lo, hi = int(sys.argv[4]), int(sys.argv[5])
    print("---- lines %d..%d" % (lo, hi))
    sub = 0
    for f, ln, src, n, stall in sorted(k["lines"], key=lambda l: l[1]):
        if f == "conv_tc.cu" and lo <= ln <= hi and n:
            sub += n
            st = sorted(stall, key=lambda x: -x[1])[:3]
            print("%6d  %d  %s   %s" % (n, ln, src[:90], st))
    print("subtotal", sub)
