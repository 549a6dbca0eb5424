# development aid: cost of the coders inside the pipelined step (same box, same build)
run() { python bench.py --steps 12 --warmup 6 --no-e2e --no-stress --no-latency --no-train --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['families']['icm_rans_encode_batch']['ms'])"; }
python -m pytest tests/test_gpu_coder.py tests/test_gpu_stf.py -x -q 2>&1 | tail -1
for w in 1 4 8; do ICM_ENC_WARPS=$w run enc_warps=$w; done
