#!/bin/bash
# ncu --set full captures of the sweep launch (2nd hybrid_topk launch of a call) on P80k and C3, plus the failing-test rerun
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu -k "emulated or multi_gpu or folded" > gpurun_out/pytest_sel.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_sel.log
for CFG in P80k C3; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hybrid_topk -s 5 -c 1 -f -o gpurun_out/r2_k1_${CFG}_final python tools/time_k1.py $CFG 0x0 > gpurun_out/ncu_full_$CFG.log 2>&1
echo "full capture $CFG exit $?"; tail -2 gpurun_out/ncu_full_$CFG.log
done
