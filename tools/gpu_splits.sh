#!/bin/bash
# K1 column-split sweep for the symmetric schedule (C3, 1 GPU)
mkdir -p gpurun_out
for s in 0 2 3 4 6 8; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --splits $s 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('splits $s', 'ms/step', round(d['ms_per_step'],2), 'k1', round(d['roofline']['kernel_ms'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'flagged', d['flagged_rows'], 'clk', d['clocks']['sm_mhz'])"
done | tee gpurun_out/splits.txt
