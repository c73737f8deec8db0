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
run default
run s3 --splits 3
run s4 --splits 4
run s6 --splits 6
run s8 --splits 8
run s4_sync32 --tuning 0x202 --splits 4
run s4_stages5 --tuning 0x50002 --splits 4
run s4_stages4 --tuning 0x40002 --splits 4
run c2_s4 --config C2 --splits 4
run c2_s8 --config C2 --splits 8
run c2_s2 --config C2 --splits 2
