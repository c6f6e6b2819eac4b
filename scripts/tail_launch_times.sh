ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_|nms_|collect|class_count" --csv --log-file gpurun_out/tail_launches.csv python scripts/profile_membound.py v3 v3-608 > gpurun_out/tailprof.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/tail_launches.csv")) if len(r)>10]
hdr=rows[0]; ki,vi=hdr.index("Kernel Name"),hdr.index("Metric Value")
for r in rows[1:]: print(r[ki].split("(")[0][:30], r[hdr.index("Grid Size")], r[vi])
PY
