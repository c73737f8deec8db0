"""Per-step K1 (candidate pass) times of one config under several tuning words: exposes run-to-run
variance that an average hides.   python tools/time_k1.py P80k 0x0 0xff0 0x20000000 ..."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tvbingefriend_recommendation_service_b200.engine import HybridTopKEngine, stage  # noqa: E402
from tvbingefriend_recommendation_service_b200.synthetic import CONFIGS, make_config  # noqa: E402

name = sys.argv[1]
tunings = [int(x, 0) for x in sys.argv[2:]] or [0]
cfg = CONFIGS[name]
cat = make_config(name)
eng = HybridTopKEngine(0)
w = (0.4, 0.5, 0.1)
dc = eng.upload(stage(cat.features()), w)
k = cfg["k"]
for tun in tunings:
    times, full = [], []
    for it in range(8):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        t = eng.top_k_device(dc, w, k, 0.1, True, phases=1, tuning=tun)
        e1.record()
        eng.top_k_device(dc, w, k, 0.1, True, phases=6, out=t, tuning=tun)
        e2.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        full.append(e0.elapsed_time(e2))
    st = t["stats"].cpu().numpy()
    print(f"{name} tuning {tun:#x}: K1 ms {np.round(times, 2).tolist()}  step ms {np.round(full[2:], 2).tolist()} "
          f"flagged {int(st[0])} pairs {int(st[1])} list entries ~{int(st[4]) * 16} (TVBF_DEBUG_COUNTS)", flush=True)
