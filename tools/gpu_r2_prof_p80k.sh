#!/bin/bash
# ncu --set full capture of the P80k sweep launch (the 6th hybrid_topk launch: seed + sweep per call)
mkdir -p gpurun_out
timeout 600 python tools/time_k1.py P80k 0x0 > gpurun_out/time_k1_p80k.txt 2>&1; cat gpurun_out/time_k1_p80k.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hybrid_topk -s 5 -c 1 -f -o gpurun_out/r2_k1_P80k_${1:-wide} python tools/time_k1.py P80k 0x0 > gpurun_out/ncu_full_p80k.log 2>&1
echo "full capture exit $?"; tail -3 gpurun_out/ncu_full_p80k.log
