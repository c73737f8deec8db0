"""Where does folding the packed groups into the operand stop paying?  K1 time with and without the
folded operand over a range of vocabularies (N = 80 000):  python tools/time_fold_crossover.py"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tvbingefriend_recommendation_service_b200.engine import HybridTopKEngine, stage  # noqa: E402
from tvbingefriend_recommendation_service_b200.synthetic import make_catalogue  # noqa: E402

eng = HybridTopKEngine(0)
w = (0.4, 0.5, 0.1)
for vocab in [int(x) for x in sys.argv[1:]] or [500, 1000, 1900, 3000, 4500]:
    cat = make_catalogue(80_000, vocab, nnz=20, meta=(21, 5, 6), seed=7)
    st = stage(cat.features())
    res = {}
    for mode, max_k in (("popcount", 0), ("folded", 1 << 20)):
        eng.fold_max_k = max_k
        dc = eng.upload(st, w)
        ts = []
        for it in range(6):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            t = eng.top_k_device(dc, w, 20, 0.1, True, phases=1)
            e1.record()
            eng.top_k_device(dc, w, 20, 0.1, True, phases=6, out=t)
            e2.record()
            torch.cuda.synchronize()
            ts.append((e0.elapsed_time(e1), e0.elapsed_time(e2)))
        res[mode] = (np.median([a for a, _ in ts[2:]]), np.median([b for _, b in ts[2:]]), int(t["stats"][0]))
        del dc
    print(f"V={vocab}: K1 popcount {res['popcount'][0]:.2f} ms folded {res['folded'][0]:.2f} ms | step {res['popcount'][1]:.2f} / "
          f"{res['folded'][1]:.2f} ms | flagged {res['popcount'][2]} / {res['folded'][2]}", flush=True)
    eng.release()
