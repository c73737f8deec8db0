#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
run() {
  tag=$1; shift
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/tune_$tag.json 2> gpurun_out/tune_$tag.err || { echo "$tag FAILED"; tail -5 gpurun_out/tune_$tag.err; return; }
  python - <<PY
import json
d=json.load(open("gpurun_out/tune_$tag.json"))
print("$tag", "ms/step", round(d["ms_per_step"],1), "k1_ms", round(d["roofline"]["kernel_ms"],1), "TF", round(d["roofline"]["achieved"]), "flagged", d["flagged_rows"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run cg2_default
run cg1_s4 --tuning 0x81 --splits 4
run cg2_s1 --splits 1
run cg2_s3 --splits 3
run c2_default --config C2
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hybrid_topk -s 3 -c 1 -o gpurun_out/prof_k1v2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"
