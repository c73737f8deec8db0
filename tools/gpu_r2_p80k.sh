#!/bin/bash
mkdir -p gpurun_out
# pacing default / off, wide epilogue auto(on) / off, combos
( timeout 600 python tools/time_k1.py P80k 0x0 0xff0 0x10000000 0x10000ff0 0x0 > gpurun_out/time_k1_p80k.txt 2>&1; echo "exit $?" ) | tee gpurun_out/summary.txt
cat gpurun_out/time_k1_p80k.txt
( timeout 600 python tools/time_k1.py C3 0x0 0x20000000 > gpurun_out/time_k1_c3.txt 2>&1; echo "exit $?" ) | tee -a gpurun_out/summary.txt
cat gpurun_out/time_k1_c3.txt
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
