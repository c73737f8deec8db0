#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -12 gpurun_out/pytest_gpu.log
( timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench exit $?" ) | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/bench_c3.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_c3.json") if l.startswith("{")][-1])
print("C3 ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "phases", d["phases_ms"])
print("api_e2e", d.get("api_e2e"))
print("extra", json.dumps(d.get("extra"), indent=1))
print("parity", d.get("parity_check"))
PY
