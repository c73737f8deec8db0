#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/tune_$tag.json 2> gpurun_out/tune_$tag.err || { echo "$tag FAILED"; tail -5 gpurun_out/tune_$tag.err; return; }
  python - <<PY
import json
d=json.load(open("gpurun_out/tune_$tag.json"))
print("$tag", "ms/step", round(d["ms_per_step"],1), "k1_ms", round(d["roofline"]["kernel_ms"],1), "TF", round(d["roofline"]["achieved"]), "flagged", d["flagged_rows"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
# tuning bits: [0:4) cg, [4:12) sync_kb, [12:16) slack, [16:20) stages, [20:28) prefetch kb, [28:30) pf mode
run cg2_base        --tuning 0x00000082 --splits 2
run cg2_stages4     --tuning 0x00040082 --splits 2
run cg2_stages3     --tuning 0x00030082 --splits 2
run cg1_stages3     --tuning 0x00030081 --splits 4
run cg1_stages2     --tuning 0x00020081 --splits 4
run cg2_pf8_m2      --tuning 0x00800082 --splits 2
run cg2_pf16_m2     --tuning 0x01000082 --splits 2
run cg2_pf32_m2     --tuning 0x02000082 --splits 2
run cg2_pf16_m1     --tuning 0x11000082 --splits 2
run cg2_pf16_m2_s4  --tuning 0x01000082 --splits 4
run cg1_pf16_m2_s4  --tuning 0x01000081 --splits 4
run cg2_pf16_nosync --tuning 0x01000ff2 --splits 2
