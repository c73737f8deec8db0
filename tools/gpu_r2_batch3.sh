#!/bin/bash
# round 2, GPU call 3 (1 GPU): test suite, default bench with extras, ncu full captures of K1 (C3 and P80k)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -14 gpurun_out/pytest_gpu.log
( timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench exit $?" ) | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/bench_c3.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_c3.json") if l.startswith("{")][-1])
print("C3 ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "phases", d["phases_ms"])
print("dense", d.get("dense_text"))
print("extra", json.dumps(d.get("extra"), indent=1))
print("parity", d.get("parity_check"))
PY
( timeout 600 python bench.py --config P80k --steps 5 --warmup 3 --extra '' --no-cpu-baseline > gpurun_out/bench_p80k.json 2> gpurun_out/bench_p80k.err; echo "bench p80k exit $?" ) | tee -a gpurun_out/summary.txt
for cfg in C3 P80k; do
  ( timeout 900 ncu --set full --clock-control none --import-source on -k regex:hybrid_topk_kernel --launch-skip 7 --launch-count 1 -f -o gpurun_out/r2_k1_${cfg} python bench.py --config $cfg --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --no-dense-probe --extra '' > gpurun_out/ncu_full_${cfg}.log 2>&1; echo "ncu full $cfg exit $?" ) | tee -a gpurun_out/summary.txt
done
ls -la gpurun_out/*.ncu-rep
