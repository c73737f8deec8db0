#!/bin/bash
# N-GPU check: distributed drivers vs single GPU, then the tile-sharded symmetric bench
NG=${1:-2}
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$NG.log 2>&1; echo "dist_check exit $?" ) | tee gpurun_out/summary.txt
tail -2 gpurun_out/dist_check_$NG.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2951$NG bench.py --gpus $NG --steps 8 --warmup 3 > gpurun_out/scale_$NG.json 2> gpurun_out/scale_$NG.err; echo "bench $NG exit $?" ) | tee -a gpurun_out/summary.txt
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/scale_$NG.json") if l.startswith("{")][-1])
print("$NG gpus:", round(d["value"]), "shows/s", round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["ms_per_step"],2), "ms; K1", round(d["roofline"]["kernel_ms"],2), "ms; flagged", d["flagged_rows"])
PY
tail -2 gpurun_out/scale_$NG.err
