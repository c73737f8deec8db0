#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -15 gpurun_out/pytest_gpu.log
run() {
  tag=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/tune_$tag.json 2> gpurun_out/tune_$tag.err || { echo "$tag FAILED"; tail -5 gpurun_out/tune_$tag.err; return; }
  python - <<PY
import json
d=json.load(open("gpurun_out/tune_$tag.json"))
print("$tag", "ms/step", round(d["ms_per_step"],1), "k1_ms", round(d["roofline"]["kernel_ms"],1), "TF(alg)", round(d["roofline"]["achieved"]), "e2e_ms", round(d["e2e"]["ms_per_step"],1), "flagged", d["flagged_rows"], "pairs", d["rescored_pairs"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run default
run s6 --splits 6
run s12 --splits 12
run seed24 --tuning 0x6000000
run seed96 --tuning 0x18000000
run c2 --config C2
run c4 --config C4
