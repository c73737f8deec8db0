"""Several GPUs of one box: features replicated, source rows sharded in 128-row tiles, result
tables gathered.  Two drivers over the same shard arithmetic:

* ``compute_top_k_distributed`` -- one process per GPU under ``torch.distributed`` (NCCL over
  NVLink); the gather is ``sharding.gather_tables``.  This is what ``bench.py --gpus N`` runs.
* ``compute_top_k_multi_gpu``   -- a single process driving ``device_ids`` (kernels are launched
  asynchronously on every device, tables are read back per shard); convenience for the drop-in
  API's ``device_ids`` argument.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .engine import HybridTopKEngine, TopK, stage
from .sharding import empty_tables, exchange_row_shards, gather_tables, row_shard


def top_k_device_distributed(eng: HybridTopKEngine, cat, weights, k: int, min_similarity: float,
                             exclude_self: bool = True, group=None, symmetric: bool | None = None,
                             splits: int = 0, tuning: int = 0, k1_events: list | None = None) -> dict:
    """One job over all ranks of ``group``; every rank returns the full gathered device table.

    * symmetric (default when eligible): tile sharding -- every rank sweeps the tiles on/above the
      diagonal of its zigzag-dealt 256-row super blocks and feeds both shows of each score; the
      partial candidate lists go through one all-to-all over the row shards (N x 32 x 8 B sent per
      rank) and each rank rescores its row shard.  Halves the tensor-core work.
    * one-sided: row sharding, no exchange before the final gather.
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = cat.n_shows
    b, e = row_shard(n, world, rank)
    if symmetric is None:
        symmetric = exclude_self and eng.sym_eligible(cat, weights, k, min_similarity) and n >= 40_000
    if symmetric:
        def all_reduce_max(t):
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)

        def exchange(t):
            return exchange_row_shards(t[:n], n, group)

        local = eng.top_k_device_sym_sharded(cat, weights, k, min_similarity, rank, world, all_reduce_max,
                                             exchange, (b, e), splits=splits, tuning=tuning, k1_events=k1_events)
    elif e > b:
        tun = tuning | (1 << 20)
        if k1_events is None:
            local = eng.top_k_device(cat, weights, k, min_similarity, exclude_self, row_begin=b, row_end=e,
                                     splits=splits, tuning=tun)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            local = eng.top_k_device(cat, weights, k, min_similarity, exclude_self, row_begin=b, row_end=e,
                                     splits=splits, tuning=tun, phases=1)
            e1.record()
            eng.top_k_device(cat, weights, k, min_similarity, exclude_self, row_begin=b, row_end=e,
                             splits=splits, tuning=tun, phases=6, out=local)
            k1_events.append(((e0, e1),))
    else:
        local = empty_tables(k, eng.device)
    full = gather_tables(local, n, k, group)
    full["row_begin"] = 0
    return full


def compute_top_k_distributed(features: dict, weights=(0.4, 0.5, 0.1), k: int = 20,
                              min_similarity: float = 0.1, metadata_mode: str = "mean3",
                              exclude_self: bool = True, engine: HybridTopKEngine | None = None,
                              group=None, staged=None, **kw) -> TopK:
    """Every rank calls this with the same features; every rank returns the full table."""
    eng = engine or HybridTopKEngine(torch.cuda.current_device())
    st = staged or stage(features, metadata_mode)
    cat = eng.upload(st, weights)
    return eng.to_host(top_k_device_distributed(eng, cat, weights, k, min_similarity, exclude_self, group, **kw))


def compute_top_k_multi_gpu(features: dict, weights=(0.4, 0.5, 0.1), k: int = 20,
                            min_similarity: float = 0.1, metadata_mode: str = "mean3",
                            exclude_self: bool = True, device_ids=(0,), **kw) -> TopK:
    st = stage(features, metadata_mode)
    world = len(device_ids)
    pending = []
    for rank, dev in enumerate(device_ids):
        b, e = row_shard(st.n_shows, world, rank)
        if e <= b:
            continue
        eng = HybridTopKEngine(dev)
        cat = eng.upload(st, weights)
        pending.append((eng, cat, eng.top_k_device(cat, weights, k, min_similarity, exclude_self,
                                                   row_begin=b, row_end=e, **kw)))
    parts = [eng.to_host(t) for eng, _cat, t in pending]
    cat_ = lambda name: np.concatenate([getattr(p, name) for p in parts], axis=0)  # noqa: E731
    return TopK(indices=cat_("indices"), counts=cat_("counts"), hybrid=cat_("hybrid"), genre=cat_("genre"),
                text=cat_("text"), metadata=cat_("metadata"), row_begin=0,
                flagged_rows=sum(p.flagged_rows for p in parts),
                rescored_pairs=sum(p.rescored_pairs for p in parts))
