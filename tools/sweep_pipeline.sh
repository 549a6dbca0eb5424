# development aid: pipeline parameter sweep (device-resident value only); usage: bash tools/sweep_pipeline.sh
run() { python bench.py --steps 10 --warmup 6 --no-e2e --no-stress --no-latency --no-train --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', '->', d['value'], d['ms_per_step'])"; }
run --streams 14 --lag 10
run --streams 16 --lag 12
run --streams 16 --lag 14
run --streams 14 --lag 12
run --streams 16 --lag 10
run --streams 20 --lag 16 --decode-priority 0
run --streams 16 --lag 12 --dec-per-cta 4
run --streams 12 --lag 10
