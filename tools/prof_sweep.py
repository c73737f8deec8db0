"""Workload for an ncu capture of the weight-sweep kernel (kMode 5): five triples sharing one
symmetric sweep on a 50 k-show slice of C3.  The 6th hybrid_topk launch of a call is the sweep
(5 seed passes precede it):
  ncu --set full --clock-control none --import-source on -k regex:hybrid_topk -s 5 -c 1 -o prof python tools/prof_sweep.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tvbingefriend_recommendation_service_b200.engine import HybridTopKEngine, stage
from tvbingefriend_recommendation_service_b200.synthetic import WEIGHT_SWEEP, make_config
cat = make_config("C3", 50000); eng = HybridTopKEngine(0)
dc = eng.upload(stage(cat.features()))
for _ in range(2):
    out = eng.top_k_sweep_device(dc, WEIGHT_SWEEP, 20, 0.1, shared=True, tuning=0x40000000)
torch.cuda.synchronize()
print("ok", [int(t["stats"][0]) for t in out])
