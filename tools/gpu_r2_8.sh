#!/bin/bash
# round 2: 8-GPU run -- distributed correctness check, scaling bench with extras, tuning variants
NG=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
run() {  # tag, extra args
  tag=$1; shift
  ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NG --steps 5 --warmup 3 "$@" > gpurun_out/scale_${NG}_$tag.json 2> gpurun_out/scale_${NG}_$tag.err; echo "bench $NG $tag exit $?" ) | tee -a gpurun_out/summary.txt
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_${NG}_$tag.json") if l.startswith("{")][-1])
    print("$tag:", round(d["value"]), "shows/s", round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["ms_per_step"],2), "ms; K1", round(d["roofline"]["kernel_ms"],2), "frac", round(d["roofline"]["frac"],3))
    print("   phases max", d["phases_ms"]); print("   phases min", d["phases_ms_min_over_ranks"]); print("   parity", d.get("parity_check")); 
    for k,v in (d.get("extra") or {}).items(): print("   extra", k, {a: (round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ("ms_per_step","k1_ms","frac","flagged_rows","error")})
except Exception as e: print("$tag failed", e)
PY
  tail -2 gpurun_out/scale_${NG}_$tag.err
}
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$NG.log 2>&1; echo "dist_check exit $?" ) | tee gpurun_out/summary.txt
tail -2 gpurun_out/dist_check_$NG.log
run default --extra 'P80k,C4,C5' --no-cpu-baseline --no-dense-probe
run s8 --splits 8 --extra '' --no-cpu-baseline --no-dense-probe --no-parity-check
run seed96 --tuning 0xd800000 --extra '' --no-cpu-baseline --no-dense-probe --no-parity-check
