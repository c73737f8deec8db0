"""Per-phase times of the tile-sharded symmetric job for every rank of WORLD, emulated on ONE GPU
(no collectives; the MAX all-reduce and the all-gather are done with torch ops outside the timed
regions):  python tools/time_sym_phases.py WORLD [CONFIG]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tvbingefriend_recommendation_service_b200.engine import HybridTopKEngine, stage
from tvbingefriend_recommendation_service_b200.sharding import row_shard
from tvbingefriend_recommendation_service_b200.synthetic import make_config

world = int(sys.argv[1]); cfg = sys.argv[2] if len(sys.argv) > 2 else "C3"
cat = make_config(cfg); eng = HybridTopKEngine(0); w = (0.4, 0.5, 0.1); k, ms = 20, 0.1
st = stage(cat.features())


def timed(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = fn(); b.record(); torch.cuda.synchronize()
    return out, a.elapsed_time(b)


dc = eng.upload(st, w)
for it in range(2):
    _, t_prep = timed(lambda: eng.upload(st, w))
    seeds, t_seed = zip(*[timed(lambda r=r: eng.sym_seed(dc, w, k, ms, r, world)) for r in range(world)])
    theta = torch.stack(list(seeds)).amax(dim=0)
    parts, t_sweep = zip(*[timed(lambda r=r: eng.sym_sweep(dc, w, k, ms, r, world, theta.clone())) for r in range(world)])
    cand = torch.stack([p[0] for p in parts]); cnt = torch.stack([p[1] for p in parts]); bound = torch.stack([p[2] for p in parts])
    t_res, flagged = [], []
    for r in range(world):
        b, e = row_shard(cat.n_shows, world, r)
        sl = (cand[:, b:e].contiguous(), cnt[:, b:e].contiguous(), bound[:, b:e].contiguous())
        out, t = timed(lambda: eng.sym_rescore(dc, w, k, ms, *sl, b, e, table_row0=b))
        t_res.append(t); flagged.append(int(out["stats"][0]))
f = lambda xs: " ".join(f"{x:6.2f}" for x in xs)
print(f"{cfg} world {world}: prep(h2d+K0) {t_prep:.2f} ms")
print("seed    ", f(t_seed)); print("sweep+K4", f(t_sweep)); print("rescore ", f(t_res)); print("flagged ", flagged)
print(f"critical path without collectives: {t_prep + max(t_seed) + max(t_sweep) + max(t_res):.2f} ms; "
      f"gathered bytes per rank: {cand[0].numel() * 4 + cnt[0].numel() * 8}")
