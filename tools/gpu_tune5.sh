#!/bin/bash
mkdir -p gpurun_out
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
run cg1_s2 --splits 2
run cg1_s3 --splits 3
run cg1_s6 --splits 6
run cg1_sync4 --tuning 0x41 --splits 4
run cg1_sync16 --tuning 0x101 --splits 4
run cg1_slack1 --tuning 0x1081 --splits 4
run cg1_slack4 --tuning 0x4081 --splits 4
run cg1_stages3 --tuning 0x30001 --splits 4
run c4_cg1 --config C4
run c4_cg2 --config C4 --tuning 0x2
