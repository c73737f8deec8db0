#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "exact or degenerate or tie_break or 2k" 2>&1 | tail -2
for cfg in C3 C4; do
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tuning 0x40000000 --config $cfg"
$CMD > gpurun_out/plain_$cfg.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:exact_rows -c 6 --csv --log-file gpurun_out/launches_$cfg.csv $CMD > gpurun_out/ncu_launches_$cfg.log 2>&1
grep exact_rows gpurun_out/launches_$cfg.csv | tail -2 | cut -c1-60,200-400
done
