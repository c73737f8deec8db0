#!/bin/bash
mkdir -p gpurun_out
for s in 4 8 12 16 24; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --splits $s > gpurun_out/splits_$s.json 2> gpurun_out/splits_$s.err
  python - <<PY
import json
d=json.load(open("gpurun_out/splits_$s.json"))
print("splits $s", "ms/step", round(d["ms_per_step"],1), "k1_ms", round(d["roofline"]["kernel_ms"],1), "TF", round(d["roofline"]["achieved"]), "flagged", d["flagged_rows"], "pairs", d["rescored_pairs"], d["clocks"])
PY
done
