#!/bin/bash
# 2-GPU check: distributed driver correctness, single-process multi-GPU test, scaling bench at N=1,2
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check exit $?" ) | tee gpurun_out/summary.txt
tail -3 gpurun_out/dist_check.log
( timeout 600 python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "multi_gpu" > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi exit $?" ) | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest_multi.log
( timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; echo "bench 1 exit $?" ) | tee -a gpurun_out/summary.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/scale_2.json 2> gpurun_out/scale_2.err; echo "bench 2 exit $?" ) | tee -a gpurun_out/summary.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/ref_2.json 2> gpurun_out/ref_2.err; echo "ref arm exit $?" ) | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.loads([l for l in open(f"gpurun_out/scale_{n}.json") if l.startswith("{")][-1])
        print(n, "gpus:", round(d["value"]), "shows/s", round(d["ms_per_step"],1), "ms/step; e2e", round(d["e2e"]["ms_per_step"],1), "ms; K1", round(d["roofline"]["kernel_ms"],1), "ms", round(d["roofline"]["achieved"]), "TF; cpu", d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(n, "failed", e)
try:
    d=json.loads([l for l in open("gpurun_out/ref_2.json") if l.startswith("{")][-1]); print("reference arm:", d["value"], d["unit"], d["cpu_baseline"]["cores"], "threads")
except Exception as e: print("ref failed", e)
PY
tail -3 gpurun_out/scale_2.err
