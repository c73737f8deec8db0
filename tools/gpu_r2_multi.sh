#!/bin/bash
# round 2: N-GPU check -- GPU test suite, distributed drivers, scaling bench at 1 and N, ncu launch list
NG=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
( timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$NG.log 2>&1; echo "dist_check exit $?" ) | tee -a gpurun_out/summary.txt
tail -4 gpurun_out/dist_check_$NG.log
( timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; echo "bench 1 exit $?" ) | tee -a gpurun_out/summary.txt
for n in $NG; do
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; echo "bench $n exit $?" ) | tee -a gpurun_out/summary.txt
  tail -3 gpurun_out/scale_$n.err
done
python - <<PY
import json
for n in ("1","$NG"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/scale_{n}.json") if l.startswith("{")][-1])
        print(n, "gpus:", round(d["value"]), "shows/s", round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["ms_per_step"],2), "ms; K1", round(d["roofline"]["kernel_ms"],2), "ms frac", round(d["roofline"]["frac"],3), "phases", d["phases_ms"], "parity", d.get("parity_check"))
    except Exception as e: print(n, "failed", e)
PY
# driver-style ncu launch list (no tuning bits): the profiler is detected and K1 launched plainly
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --no-dense-probe > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" ) | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/ncu_launches.log
grep -c hybrid_topk gpurun_out/r2_launches_c3.csv
ncu --metrics gpu__time_duration.sum --clock-control none bash -c 'env | sort > gpurun_out/env_under_ncu.txt; python -c "print(open(\"/proc/self/maps\").read())" > gpurun_out/maps_under_ncu.txt' > /dev/null 2>&1
env | sort > gpurun_out/env_plain.txt
