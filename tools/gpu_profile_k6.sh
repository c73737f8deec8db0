#!/bin/bash
# one full ncu capture of the exact-rows kernel (K6) inside the bench command (1 GPU)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tuning 0x40000000"
$CMD > gpurun_out/plain_k6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:exact_rows -s 3 -c 1 -o gpurun_out/prof_k6 $CMD > gpurun_out/ncu_k6.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_k6.log
