#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "streaming" --durations=4 > gpurun_out/pytest_stats.log 2>&1; echo "pytest stats exit $?" ) | tee gpurun_out/summary.txt
tail -25 gpurun_out/pytest_stats.log
( timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all exit $?" ) | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
