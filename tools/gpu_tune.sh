#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -15 gpurun_out/pytest_gpu.log
run() {
  tag=$1; shift
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/tune_$tag.json 2> gpurun_out/tune_$tag.err || { echo "$tag FAILED"; tail -5 gpurun_out/tune_$tag.err; return; }
  python - <<PY
import json
d=json.load(open("gpurun_out/tune_$tag.json"))
print("$tag", "ms/step", round(d["ms_per_step"],1), "k1_ms", round(d["roofline"]["kernel_ms"],1), "TF", round(d["roofline"]["achieved"]), "flagged", d["flagged_rows"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
# cta_group 1, pacing off (previous kernel) / on
run cg1_nosync_s4 --tuning 0xff1 --splits 4
run cg1_sync_s4 --tuning 0x1 --splits 4
run cg1_sync_s12 --tuning 0x1 --splits 12
run cg2_nosync_s2 --tuning 0xff2 --splits 2
run cg2_sync_s2 --tuning 0x2 --splits 2
run cg2_sync_s4 --tuning 0x2 --splits 4
run cg2_sync_s8 --tuning 0x2 --splits 8
run cg2_sync8_s4 --tuning 0x82 --splits 4
run cg2_sync2_s4 --tuning 0x22 --splits 4
