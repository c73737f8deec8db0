#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_scale.py -x -q -m gpu --durations=8 > gpurun_out/pytest_scale.log 2>&1; echo "pytest scale exit $?" ) | tee gpurun_out/summary.txt
tail -16 gpurun_out/pytest_scale.log
run() {
  tag=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/tune_$tag.json 2> gpurun_out/tune_$tag.err || { echo "$tag FAILED"; tail -5 gpurun_out/tune_$tag.err; return; }
  python - <<PY
import json
d=json.load(open("gpurun_out/tune_$tag.json"))
print("$tag", "ms/step", round(d["ms_per_step"],1), "k1_ms", round(d["roofline"]["kernel_ms"],1), "TF(alg)", round(d["roofline"]["achieved"]), "e2e_ms", round(d["e2e"]["ms_per_step"],1), "flagged", d["flagged_rows"], "pairs", d["rescored_pairs"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run seed48
run seed62 --tuning 0xF800000
run seed32 --tuning 0x8000000
run seed16 --tuning 0x4000000
