#!/bin/bash
# Full GPU check: parity tests, smoke, bench (C2 + C3) and the ncu launch list of the bench command.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" ) | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/smoke.log
( timeout 600 python bench.py --config C2 --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $?" ) | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_c2.json
( timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 exit $?" ) | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
