"""top-100 on C3 / C5-sized catalogues: symmetric against one-sided sweep.
python tools/time_k100.py [CONFIG] [N]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tvbingefriend_recommendation_service_b200.engine import HybridTopKEngine, stage
from tvbingefriend_recommendation_service_b200.synthetic import make_config

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
cat = make_config(cfg, n); eng = HybridTopKEngine(0); w = (0.4, 0.5, 0.1)
dc = eng.upload(stage(cat.features()), w)
res = {}
for name, tun in (("one-sided", 1 << 20), ("symmetric", 2 << 20)):
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = eng.top_k_device(dc, w, 100, 0.1, tuning=tun); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    res[name] = out
    print(f"{cfg} n={cat.n_shows} k=100 {name}: {min(ts):.1f} ms, flagged {int(out['stats'][0])}, rescored pairs {int(out['stats'][1])}", flush=True)
a, b = res["one-sided"], res["symmetric"]
print("tables identical:", torch.equal(a["indices"], b["indices"]) and torch.equal(a["counts"], b["counts"]))
