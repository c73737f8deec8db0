#!/bin/bash
mkdir -p gpurun_out
for v in "cg1 0x40000001 4" "cg2 0x40000002 2"; do
  set -- $v
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --tuning $2 --splits $3"
  $CMD > gpurun_out/plain_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:hybrid_topk -s 3 -c 1 -o gpurun_out/prof_k1_$1 $CMD > gpurun_out/ncu_full_$1.log 2>&1
  echo "$1 full capture exit $?"
done
