"""Time one rank's share of the tile-sharded symmetric sweep on a single GPU (no collectives):
python tools/time_sym_rank.py WORLD [CONFIG]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tvbingefriend_recommendation_service_b200.engine import HybridTopKEngine, stage
from tvbingefriend_recommendation_service_b200.synthetic import make_config

world = int(sys.argv[1]); cfg = sys.argv[2] if len(sys.argv) > 2 else "C3"
cat = make_config(cfg); eng = HybridTopKEngine(0); w = (0.4, 0.5, 0.1)
dc = eng.upload(stage(cat.features()), w)
for splits in (4, 6, 8, 10, 12, 16):
    for rank in (0, world - 1):
        ts = []
        for it in range(3):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
            thetas = [eng.sym_seed(dc, w, 20, 0.1, r, world, splits=splits) for r in range(world) if r != rank]
            e0.record()
            theta = eng.sym_seed(dc, w, 20, 0.1, rank, world, splits=splits)
            e1.record()
            theta = torch.stack(thetas + [theta]).amax(dim=0)   # what the MAX all-reduce delivers
            torch.cuda.synchronize(); e1 = torch.cuda.Event(enable_timing=True); e1.record()
            eng.sym_sweep(dc, w, 20, 0.1, rank, world, theta, splits=splits)
            e2.record(); torch.cuda.synchronize()
            ts.append((0.0, e1.elapsed_time(e2)))
        print(f"world {world} rank {rank} splits {splits}: seed {ts[-1][0]:.2f} ms sweep+compact {ts[-1][1]:.2f} ms", flush=True)
