#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" ) | tee gpurun_out/summary.txt
tail -12 gpurun_out/pytest_gpu.log
# memcheck on a tiny job (both sweeps): one compute-sanitizer tool per call
cat > /tmp/tiny.py <<'PY'
import sys; sys.path.insert(0, ".")
from tvbingefriend_recommendation_service_b200.engine import HybridTopKEngine
from tvbingefriend_recommendation_service_b200.synthetic import make_catalogue
cat = make_catalogue(700, 256, nnz=10, seed=1)
eng = HybridTopKEngine(0)
a = eng.compute_top_k(cat.features(), (0.4, 0.5, 0.1), 20, 0.1, tuning=(1 << 20) | (1 << 30))
b = eng.compute_top_k(cat.features(), (0.4, 0.5, 0.1), 20, 0.1, tuning=(2 << 20) | (1 << 30))
c = eng.compute_top_k(cat.features(), (0.4, 0.5, 0.1), 20, 0.1, force_exact=True)
import numpy as np
assert np.array_equal(a.indices, c.indices) and np.array_equal(b.indices, c.indices)
print("tiny ok", a.flagged_rows, b.flagged_rows)
PY
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/tiny.py > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?" ) | tee -a gpurun_out/summary.txt
tail -6 gpurun_out/memcheck.log
