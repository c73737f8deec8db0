#!/bin/bash
mkdir -p gpurun_out
for cfg in C3 C4; do
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tuning 0x40000000 --config $cfg"
$CMD > gpurun_out/plain_$cfg.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$cfg.csv $CMD > gpurun_out/ncu_launches_$cfg.log 2>&1
echo "$cfg launch list exit $?"
done
