#!/bin/bash
# producer pacing sweep (tuning bits 4-11 chunk in k-blocks, bits 12-15 slack) for the default schedule, C3
mkdir -p gpurun_out
for cfg in "0 0" "8 2" "32 2" "16 1" "16 3" "16 4" "24 3" "255 0"; do
set -- $cfg
t=$(( ($1 << 4) | ($2 << 12) ))
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --tuning $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk $1 slack $2', 'ms/step', round(d['ms_per_step'],2), 'k1', round(d['roofline']['kernel_ms'],2), 'clk', d['clocks']['sm_mhz'])"
done | tee gpurun_out/pacing.txt
