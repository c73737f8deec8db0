#!/bin/bash
# seed-pass stride sweep of the symmetric schedule (tuning bits 22-27), C3 on 1 GPU
mkdir -p gpurun_out
for s in 0 50 52 54 58 62; do
t=$(( s << 22 ))
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tuning $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stride $s', 'ms/step', round(d['ms_per_step'],2), 'k1', round(d['roofline']['kernel_ms'],2), 'flagged', d['flagged_rows'], 'clk', d['clocks']['sm_mhz'])"
done | tee gpurun_out/seed_stride.txt
