#!/bin/bash
# parity suite, then K6 (exact rows) launch times inside the bench command for C3 and C4
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -4 gpurun_out/pytest_gpu.log
for cfg in ${CFGS:-C3 C4}; do
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --tuning 0x40000000 --config $cfg"
$CMD > gpurun_out/plain_$cfg.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"exact_rows|rescore" -c 10 --csv --log-file gpurun_out/launches_$cfg.csv $CMD > gpurun_out/ncu_launches_$cfg.log 2>&1
python - <<PY
import csv, json
rows = [r for r in csv.reader(open("gpurun_out/launches_$cfg.csv")) if len(r) > 5 and r[0].isdigit()]
for r in rows[-4:]:
    print("$cfg", r[4][:40], r[-1], r[-2])
d=json.loads([l for l in open("gpurun_out/plain_$cfg.log") if l.startswith("{")][-1])
print("$cfg", "ms/step", round(d["ms_per_step"],2), "k1_ms", round(d["roofline"]["kernel_ms"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "flagged", d["flagged_rows"])
PY
done
