#!/bin/bash
# N GPUs: the multi-GPU unit test, the distributed drivers' self-check, the scaling bench line with P80k
NG=${1:-2}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -x -q -m gpu -k "multi_gpu or emulated" > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_multi.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$NG.log 2>&1; echo "dist_check exit $?" ) | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/dist_check_$NG.log
for tag in default; do
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --steps 5 --warmup 3 --extra "${EXTRA:-P80k}" --no-cpu-baseline --no-dense-probe > gpurun_out/scale_${NG}.json 2> gpurun_out/scale_${NG}.err; echo "bench $NG exit $?" ) | tee -a gpurun_out/summary.txt
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_${NG}.json") if l.startswith("{")][-1])
    print("$NG gpus:", round(d["value"]), "shows/s", round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["ms_per_step"],2), "ms; K1", round(d["roofline"]["kernel_ms"],2), "frac", round(d["roofline"]["frac"],3))
    print("   phases max", d["phases_ms"]); print("   phases min", d["phases_ms_min_over_ranks"]); print("   parity", d.get("parity_check"))
    for k,v in (d.get("extra") or {}).items(): print("   extra", k, {a: (round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ("ms_per_step","k1_ms","frac","flagged_rows","error","tile")})
except Exception as e: print("failed", e)
PY
tail -3 gpurun_out/scale_${NG}.err
done
