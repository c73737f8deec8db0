#!/bin/bash
# round 2, GPU call 1: full GPU test suite, default bench, ncu launch list (driver-style: no tuning bits)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
( timeout 1500 python -m pytest tests -x -q -m gpu -s --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -30 gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench exit $?" ) | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/bench_c3.err
cat gpurun_out/bench_c3.json | head -c 6000
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --no-dense-probe > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" ) | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/ncu_launches.log
grep -c hybrid_topk gpurun_out/r2_launches_c3.csv
