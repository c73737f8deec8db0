#!/bin/bash
NG=${1:-8}
mkdir -p gpurun_out
for steps in 5 20; do
BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --steps $steps --warmup 3 --extra '' --no-cpu-baseline --no-dense-probe --no-parity-check > gpurun_out/hb_${NG}_$steps.json 2> gpurun_out/hb_${NG}_$steps.err
grep "^step" gpurun_out/hb_${NG}_$steps.err | head -20
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/hb_${NG}_$steps.json") if l.startswith("{")][-1])
print("$NG gpus steps $steps:", round(d["ms_per_step"],2), "ms/step; host enqueue", d["host_enqueue_ms_per_step"], "e2e", round(d["e2e"]["ms_per_step"],2), "phases", d["phases_ms"])
PY
done
