#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -4 gpurun_out/pytest_gpu.log
for cfg in C3 C4; do
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tuning 0x40000000 --config $cfg"
$CMD > gpurun_out/plain_$cfg.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:exact_rows -c 6 --csv --log-file gpurun_out/launches_$cfg.csv $CMD > gpurun_out/ncu_launches_$cfg.log 2>&1
grep exact_rows gpurun_out/launches_$cfg.csv | tail -2 | cut -c1-60,200-400
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/plain_$cfg.log") if l.startswith("{")][-1])
print("$cfg", "ms/step", round(d["ms_per_step"],1), "k1_ms", round(d["roofline"]["kernel_ms"],1), "e2e", round(d["e2e"]["ms_per_step"],1))
PY
done
