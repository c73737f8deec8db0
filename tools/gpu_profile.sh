#!/bin/bash
# ncu launch list + one full capture of the top kernel for the bench command (1 GPU).
# bit 30 of --tuning: plain (non-cooperative) launch, because ncu's SASS-patching passes change the
# kernel's resource footprint and a cooperative launch of the patched kernel is refused.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --tuning 0x40000000"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hybrid_topk -s 7 -c 1 -o gpurun_out/prof_k1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
cut -c1-300 gpurun_out/plain.log
