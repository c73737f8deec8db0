#!/bin/bash
NG=${1:-2}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -4 gpurun_out/pytest_gpu.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$NG.log 2>&1; echo "dist_check exit $?" ) | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/dist_check_$NG.log
( BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --steps 5 --warmup 3 --extra '' --no-cpu-baseline --no-dense-probe > gpurun_out/scale_$NG.json 2> gpurun_out/scale_$NG.err; echo "bench $NG exit $?" ) | tee -a gpurun_out/summary.txt
grep "^step" gpurun_out/scale_$NG.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/scale_$NG.json") if l.startswith("{")][-1])
print("$NG gpus:", round(d["value"]), "shows/s", round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["ms_per_step"],2), "ms; K1", round(d["roofline"]["kernel_ms"],2), "frac", round(d["roofline"]["frac"],3))
print("   phases max", d["phases_ms"]); print("   phases min", d["phases_ms_min_over_ranks"]); print("   parity", d.get("parity_check"))
PY
( TVBF_PEER_EXCHANGE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $NG --steps 5 --warmup 3 --extra '' --no-cpu-baseline --no-dense-probe > gpurun_out/scale_${NG}_a2a.json 2> gpurun_out/scale_${NG}_a2a.err; echo "bench $NG a2a exit $?" ) | tee -a gpurun_out/summary.txt
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/scale_${NG}_a2a.json") if l.startswith("{")][-1])
print("$NG gpus (NCCL all_to_all):", round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["ms_per_step"],2))
print("   phases max", d["phases_ms"]); print("   phases min", d["phases_ms_min_over_ranks"]); print("   parity", d.get("parity_check",{}).get("ok"))
PY
grep -h "peer exchange" gpurun_out/*.err gpurun_out/*.log | head -3
( BENCH_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --extra '' --no-cpu-baseline --no-dense-probe > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; echo "bench 1 exit $?" ) | tee -a gpurun_out/summary.txt
grep "^step" gpurun_out/scale_1.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/scale_1.json") if l.startswith("{")][-1])
print("1 gpu:", round(d["value"]), "shows/s", round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["ms_per_step"],2), "api", d["api_e2e"]["ms"], "phases", d["phases_ms"], "parity ok", d["parity_check"]["ok"])
PY
