#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
run() {
  tag=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/tune_$tag.json 2> gpurun_out/tune_$tag.err || { echo "$tag FAILED"; tail -5 gpurun_out/tune_$tag.err; return; }
  python - <<PY
import json
d=json.load(open("gpurun_out/tune_$tag.json"))
print("$tag", "ms/step", round(d["ms_per_step"],1), "k1_ms", round(d["roofline"]["kernel_ms"],1), "TF", round(d["roofline"]["achieved"]), "e2e_ms", round(d["e2e"]["ms_per_step"],1), "flagged", d["flagged_rows"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run cg1_default
run cg1_s4 --splits 4
run cg2_s2 --tuning 0x2 --splits 2
run cg2_s4 --tuning 0x2 --splits 4
run cg2_s4_sync16 --tuning 0x102 --splits 4
run cg2_nosync --tuning 0xff2 --splits 2
run c2_cg1 --config C2
run c2_cg2 --config C2 --tuning 0x2
run c4_cg2 --config C4 --tuning 0x2
