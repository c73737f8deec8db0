#!/bin/bash
# one development iteration: K1 timings on P80k / C3, the GPU parity suite, the weight sweep
mkdir -p gpurun_out
timeout 600 python tools/time_k1.py P80k 0x0 > gpurun_out/time_k1_p80k.txt 2>&1; cat gpurun_out/time_k1_p80k.txt
timeout 600 python tools/time_k1.py C3 0x0 > gpurun_out/time_k1_c3.txt 2>&1; cat gpurun_out/time_k1_c3.txt
( timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/time_sweep.py > gpurun_out/time_sweep.txt 2>&1; tail -5 gpurun_out/time_sweep.txt
