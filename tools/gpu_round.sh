#!/bin/bash
# What the driver runs at round end, in one call: GPU parity tests, smoke, both bench arms.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" ) | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/smoke.log
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?" ) | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_reference.json | cut -c1-600
( timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" ) | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
