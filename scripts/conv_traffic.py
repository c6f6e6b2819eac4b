#!/usr/bin/env python3
"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:conv_tc` CSV of ONE
forward pass into profiles/<name>.json: total DRAM bytes moved by the conv_tc family per step (bench.py's roofline.traffic)."""
import csv, json, sys
src, dst, launches_per_step = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]; ii, mi, vi, ui = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = {}
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    unit = r[ui].lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
    per.setdefault(int(r[ii]), {})[r[mi]] = v * mult
ids = sorted(per)[:launches_per_step]
rd = sum(per[i].get("dram__bytes_read.sum", 0) for i in ids); wr = sum(per[i].get("dram__bytes_write.sum", 0) for i in ids)
ns = sum(per[i].get("gpu__time_duration.sum", 0) for i in ids)
json.dump({"launches": len(ids), "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_total": rd + wr, "sum_kernel_ns_under_ncu": ns,
           "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc (one YOLOv3-416 batch-64 forward)"},
          open(dst, "w"), indent=1)
print(open(dst).read())
