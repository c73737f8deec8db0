"""Weight sweep (the five schemes of SURVEY section 8d / notebooks/03 cell 6) with the shared
symmetric tensor-core sweep against five separate jobs:  python tools/time_sweep.py [CONFIG] [K] [N]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tvbingefriend_recommendation_service_b200.engine import HybridTopKEngine, stage
from tvbingefriend_recommendation_service_b200.synthetic import WEIGHT_SWEEP, make_config

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = int(sys.argv[3]) if len(sys.argv) > 3 else None
cat = make_config(cfg, n); eng = HybridTopKEngine(0)
dc = eng.upload(stage(cat.features()))


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return out, min(ts)


sep, t_sep = timed(lambda: eng.top_k_sweep_device(dc, WEIGHT_SWEEP, k, 0.1, shared=False))
swp, t_swp = timed(lambda: eng.top_k_sweep_device(dc, WEIGHT_SWEEP, k, 0.1, shared=True))
same = all(torch.equal(a["indices"], b["indices"]) and torch.equal(a["counts"], b["counts"]) and
           torch.equal(a["hybrid"].nan_to_num(), b["hybrid"].nan_to_num()) for a, b in zip(sep, swp))
print(f"{cfg} n={cat.n_shows} k={k}: separate {t_sep:.1f} ms, shared sweep {t_swp:.1f} ms, speed-up {t_sep / t_swp:.2f}x, "
      f"tables identical: {same}; flagged separate {[int(t['stats'][0]) for t in sep]} shared {[int(t['stats'][0]) for t in swp]}")
