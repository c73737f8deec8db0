#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
BENCH_DEBUG=1 timeout 600 python bench.py --config P80k --steps 6 --warmup 3 --extra '' --no-cpu-baseline --no-dense-probe > gpurun_out/bench_p80k_$i.json 2> gpurun_out/bench_p80k_$i.err
grep "^step" gpurun_out/bench_p80k_$i.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_p80k_$i.json") if l.startswith("{")][-1])
print("P80k run $i: ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "k1", round(d["roofline"]["kernel_ms"],2), "api", d.get("api_e2e",{}).get("ms"), d["clocks"])
PY
done
BENCH_DEBUG=1 timeout 600 python bench.py --steps 4 --warmup 3 --extra 'P80k' --no-cpu-baseline --no-dense-probe > gpurun_out/bench_c3_x.json 2> gpurun_out/bench_c3_x.err
grep "^step" gpurun_out/bench_c3_x.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_c3_x.json") if l.startswith("{")][-1])
print("C3: ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "api", d.get("api_e2e"), "extra", d["extra"])
PY
