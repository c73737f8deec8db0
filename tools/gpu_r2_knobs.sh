#!/bin/bash
# experiment knobs of the symmetric sweep (environment): refresh schedule, sleeping waits
mkdir -p gpurun_out
export TVBF_DEBUG_COUNTS=1
for cfg in P80k C3; do
for v in "TVBF_REFRESH_PERIOD=0 TVBF_WAIT_NS=64" "TVBF_REFRESH_PERIOD=32 TVBF_WAIT_NS=64" "TVBF_REFRESH_PERIOD=64 TVBF_WAIT_NS=64" "TVBF_REFRESH_PERIOD=0 TVBF_WAIT_NS=16" "TVBF_REFRESH_PERIOD=32 TVBF_WAIT_NS=16" "TVBF_REFRESH_PERIOD=0 TVBF_WAIT_NS=256"; do
  echo "== $cfg $v"
  env $v timeout 600 python tools/time_k1.py $cfg 0x0 2>&1 | tail -1
done
done | tee gpurun_out/knobs.txt
