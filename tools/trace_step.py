"""Kernel timeline of one distributed step on rank 0 (torch.profiler, CUDA activities):
   torchrun --nproc-per-node N tools/trace_step.py [CONFIG]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile
from tvbingefriend_recommendation_service_b200.engine import HybridTopKEngine, stage
from tvbingefriend_recommendation_service_b200.multi_gpu import top_k_device_distributed
from tvbingefriend_recommendation_service_b200.synthetic import CONFIGS, make_config

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = HybridTopKEngine(local)
cat = make_config(name)
k = CONFIGS[name]["k"]
w = (0.4, 0.5, 0.1)
raw = eng.h2d(stage(cat.features(), "mean3", pin=True))
prev = None


def step():
    global prev
    dc = eng.prepare(raw, w, recycle=prev)
    prev = dc
    if world > 1:
        return top_k_device_distributed(eng, dc, w, k, 0.1, True)
    return eng.top_k_device(dc, w, k, 0.1, True)


for _ in range(4):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    last_end = t0
    print(f"{'start_us':>10s} {'dur_us':>9s} {'gap_us':>8s}  kernel")
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        print(f"{s:10.1f} {d:9.1f} {e.time_range.start - last_end:8.1f}  {e.name[:90]}")
        last_end = max(last_end, e.time_range.end)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
