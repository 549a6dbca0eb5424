#!/bin/bash
# Evidence capture for profiles/ (run on the GPU box: gpurun -- 'bash tools/capture_profiles.sh r02').
# Every ncu pass runs only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 8 --warmup 6 > $OUT/${TAG}_bench_line.json 2> $OUT/${TAG}_bench.err || { echo "bench failed"; tail -5 $OUT/${TAG}_bench.err; exit 1; }
python tools/conv_breakdown.py 64 > $OUT/${TAG}_conv_breakdown_b64.txt 2>&1
LL="bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-stress --no-latency --no-train"
python $LL > $OUT/${TAG}_launchlist_plain.json 2>/dev/null && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launch_list_step.csv python $LL > $OUT/${TAG}_launchlist_ncu.log 2>&1
python tools/ncu_capture.py 64 stf ${TAG}_full > $OUT/${TAG}_capture_plain.log 2>&1 && \
ncu --set full --clock-control none --profile-from-start off -o /tmp/${TAG}_full python tools/ncu_capture.py 64 stf ${TAG}_full > $OUT/${TAG}_capture_ncu.log 2>&1 && \
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null && \
python tools/ncu_summarise.py $OUT/${TAG}_full_raw.csv $OUT/${TAG}_full_keys.json > $OUT/${TAG}_ncu_full_per_shape.txt 2>&1
tail -12 $OUT/${TAG}_ncu_full_per_shape.txt
wc -l $OUT/${TAG}_launch_list_step.csv
