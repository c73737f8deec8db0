#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check8.log 2>&1; echo "dist_check8 exit $?" ) | tee gpurun_out/summary.txt
tail -2 gpurun_out/dist_check8.log
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    ( timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; echo "bench 1 exit $?" ) | tee -a gpurun_out/summary.txt
  else
    ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; echo "bench $n exit $?" ) | tee -a gpurun_out/summary.txt
  fi
done
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 3 --warmup 3 --config C4 > gpurun_out/scale_8_c4.json 2> gpurun_out/scale_8_c4.err; echo "bench 8 C4 exit $?" ) | tee -a gpurun_out/summary.txt
python - <<PY
import json
for n in ("1","2","4","8","8_c4"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/scale_{n}.json") if l.startswith("{")][-1])
        print(n, "gpus:", round(d["value"]), "shows/s", round(d["ms_per_step"],1), "ms/step; e2e", round(d["e2e"]["ms_per_step"],1), "ms; K1", round(d["roofline"]["kernel_ms"],1), "ms", round(d["roofline"]["achieved"]), "TF/gpu alg,", round(d["roofline"]["executed_tflops"]), "exec; flagged", d["flagged_rows"])
    except Exception as e: print(n, "failed", e)
PY
tail -3 gpurun_out/scale_8.err
