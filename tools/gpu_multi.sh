#!/bin/bash
# N-GPU check: distributed drivers (row-sharded one-sided and tile-sharded symmetric), scaling bench
NG=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check exit $?" ) | tee gpurun_out/summary.txt
tail -3 gpurun_out/dist_check.log
for n in 1 $NG; do
  if [ $n -eq 1 ]; then
    ( timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; echo "bench 1 exit $?" ) | tee -a gpurun_out/summary.txt
  else
    ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; echo "bench $n exit $?" ) | tee -a gpurun_out/summary.txt
    ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 --one-sided > gpurun_out/scale_${n}_onesided.json 2> gpurun_out/scale_${n}_onesided.err; echo "bench $n one-sided exit $?" ) | tee -a gpurun_out/summary.txt
  fi
done
python - <<PY
import json
for n in ("1","$NG","${NG}_onesided"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/scale_{n}.json") if l.startswith("{")][-1])
        print(n, "gpus:", round(d["value"]), "shows/s", round(d["ms_per_step"],1), "ms/step; e2e", round(d["e2e"]["ms_per_step"],1), "ms; K1", round(d["roofline"]["kernel_ms"],1), "ms", round(d["roofline"]["achieved"]), "TF/gpu; flagged", d["flagged_rows"])
    except Exception as e: print(n, "failed", e)
PY
tail -3 gpurun_out/scale_$NG.err
