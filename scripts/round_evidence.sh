#!/bin/bash
# One B200: the records and ncu summaries committed under profiles/ for a round (usage: bash scripts/round_evidence.sh r2).
# Nothing printed by a command that runs under ncu is a bench value.
R=${1:-r2}
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 3 > $O/${R}_bench.json 2> $O/${R}_bench.err
python bench.py --impl reference --steps 4 --warmup 1 > $O/${R}_bench_reference.json 2>> $O/${R}_bench.err
# launch list of the bench command (cold-cache, serialised times: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 400 --csv --log-file $O/${R}_launches_bench_ncu.csv \
    python bench.py --steps 2 --warmup 3 --headline-only --no-cpu-baseline > /dev/null 2>> $O/${R}_bench.err
python scripts/launch_summary.py $O/${R}_launches_bench_ncu.csv > $O/${R}_launches_summary.txt
# DRAM traffic of the convolution family over one forward pass
B200_NO_COPY_OVERLAP=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_ -c 74 \
    --csv --log-file $O/${R}_traffic.csv python scripts/layer_times.py yolov3 64 416 > $O/${R}_layer_times_under_ncu.txt 2>&1
python scripts/layer_times.py yolov3 64 416 > $O/${R}_layer_times.txt 2>&1
for c in "yolov3-tiny 64 416" "yolov2 64 416" "yolov1 64 448" "yolov3 32 608"; do echo "== $c"; python scripts/layer_times.py $c 2>&1 | tail -2; done > $O/${R}_other_models.txt
tail -c 300 $O/${R}_bench.err
