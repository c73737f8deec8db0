"""Several GPUs of one box: features replicated, result rows sharded, tables gathered
(SURVEY.md section 8e).  Drivers over the shard arithmetic of ``sharding.py``:

* ``top_k_device_distributed``  -- one process per GPU under ``torch.distributed`` (NCCL over
  NVLink), device-resident features in, gathered device tables out.  Per job: one all-reduce(MAX) of
  the seeded thresholds, one all-to-all of the packed candidate rows, one coalesced in-place
  all-gather of the result tables.  This is what ``bench.py --gpus N`` times as ``value``.
* ``DistributedTopK``           -- the end-to-end form of the same job (host features in, host table
  out): every rank uploads 1/world of the feature bytes and NVLink replicates them; every rank copies
  its shard of the result into ONE pinned host table in shared memory.  ``bench.py``'s ``e2e``.
* ``compute_top_k_multi_gpu``   -- a single process driving ``device_ids`` (kernels are launched
  asynchronously on every device, tables are read back per shard); convenience for the drop-in
  API's ``device_ids`` argument.
"""

from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from .engine import HybridTopKEngine, StagedCatalogue, TopK, stage
import time

from ._lib import check
from .sharding import (PeerCandidateBuffers, PeerTables, ShardedUpload, SharedHostTable, alloc_full_tables,
                       exchange_packed, gather_full_tables, peer_table_layout, row_shard, shard_rows, shard_views)

_peer_cache: dict = {}


def peer_buffers(eng: HybridTopKEngine, cat, weights, k: int, min_similarity: float, group=None):
    """The fused exchange's receive buffers for this job shape (created collectively on first use,
    then cached), or None when torch symmetric memory is unavailable or TVBF_PEER_EXCHANGE=0 -- the
    driver then falls back to one NCCL all-to-all.  Every rank takes the same decision."""
    if os.environ.get("TVBF_PEER_EXCHANGE", "1") == "0":
        return None
    world = dist.get_world_size(group)
    p = eng._params(cat, weights, k, min_similarity)
    import ctypes as C

    L = int(eng.lib.tvbf_sym_list_len(C.byref(cat.c), C.byref(p)))
    key = (eng.device.index, cat.n_shows, L, world, id(group))
    if key not in _peer_cache:
        ok = torch.ones((1,), dtype=torch.int32, device=eng.device)
        bufs = None
        try:
            with torch.cuda.device(eng.device):
                bufs = PeerCandidateBuffers(cat.n_shows, L, eng.device, group)
        except Exception as exc:      # no symmetric memory on this system / build
            ok.zero_()
            print(f"tvbf: peer exchange unavailable ({type(exc).__name__}: {exc}); using all_to_all", flush=True)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        _peer_cache[key] = bufs if int(ok.item()) == 1 else None
    return _peer_cache[key]


_peer_table_cache: dict = {}


def peer_tables(eng: HybridTopKEngine, n_shows: int, k: int, group=None):
    """Result tables in symmetric memory for this job shape (created collectively on first use, then
    cached), or None when torch symmetric memory is unavailable or TVBF_PEER_GATHER=0 -- the driver
    then gathers with one coalesced NCCL all-gather.  Every rank takes the same decision."""
    if os.environ.get("TVBF_PEER_GATHER", "1") == "0":
        return None
    world = dist.get_world_size(group)
    # two sets of symmetric memory per job shape stay allocated: only where the gather is a visible share
    # of the job (72 MB at C3, 181 MB at C4; C5's 720 MB tables go through NCCL -- 0.5 % of a 170 ms job)
    if peer_table_layout(n_shows, k, world)[1] > (256 << 20):
        return None
    key = (eng.device.index, n_shows, k, world, id(group))
    if key not in _peer_table_cache:
        ok = torch.ones((1,), dtype=torch.int32, device=eng.device)
        tabs = None
        try:
            with torch.cuda.device(eng.device):
                tabs = PeerTables(n_shows, k, eng.device, group)
        except Exception as exc:      # no symmetric memory on this system / build
            ok.zero_()
            print(f"tvbf: peer gather unavailable ({type(exc).__name__}: {exc}); using all_gather", flush=True)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        _peer_table_cache[key] = tabs if int(ok.item()) == 1 else None
    return _peer_table_cache[key]


def _local_tables(eng: HybridTopKEngine, cat, weights, k: int, min_similarity: float, exclude_self: bool, group,
                  symmetric: bool | None, splits: int, tuning: int, events: dict | None, mine: dict) -> dict:
    """This rank's rows of the job, written into ``mine`` (views of the gather buffer)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = cat.n_shows
    b, e = row_shard(n, world, rank)
    if symmetric is None:
        symmetric = exclude_self and eng.sym_eligible(cat, weights, k, min_similarity) and n >= 40_000
    if symmetric:
        def all_reduce_max(t):
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)

        def exchange(packed):
            return exchange_packed(packed, group)

        return eng.top_k_device_sym_sharded(cat, weights, k, min_similarity, rank, world, all_reduce_max, exchange,
                                            (b, e), splits=splits, tuning=tuning, events=events, out=mine,
                                            padded_rows=world * shard_rows(n, world),
                                            peer=peer_buffers(eng, cat, weights, k, min_similarity, group))
    if e > b:
        tun = tuning | (1 << 20)
        if events is None:
            return eng.top_k_device(cat, weights, k, min_similarity, exclude_self, row_begin=b, row_end=e,
                                    splits=splits, tuning=tun, out=mine)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        eng.top_k_device(cat, weights, k, min_similarity, exclude_self, row_begin=b, row_end=e, splits=splits,
                         tuning=tun, phases=1, out=mine)
        e1.record()
        eng.top_k_device(cat, weights, k, min_similarity, exclude_self, row_begin=b, row_end=e, splits=splits,
                         tuning=tun, phases=6, out=mine)
        e2.record()
        for name, ev in (("seed0", e0), ("seed1", e0), ("reduce1", e0), ("sweep1", e1), ("exchange1", e1),
                         ("rescore1", e2)):
            events.setdefault(name, []).append(ev)
        return mine
    eng._zero(mine["stats"])
    return mine


def top_k_device_distributed(eng: HybridTopKEngine, cat, weights, k: int, min_similarity: float,
                             exclude_self: bool = True, group=None, symmetric: bool | None = None,
                             splits: int = 0, tuning: int = 0, events: dict | None = None,
                             tables: dict | None = None) -> dict:
    """One job over all ranks of ``group``; every rank returns the full gathered device table.

    * symmetric (default when eligible): tile sharding -- every rank sweeps the tiles on/above the
      diagonal of the super-block groups dealt to it and feeds both shows of each score; the packed
      partial candidate lists go through ONE all-to-all over the row shards (N x 33 x 8 B sent per
      rank) and each rank rescores its row shard.  Halves the tensor-core work.
    * one-sided: row sharding, no exchange before the final gather.

    The tables live in torch symmetric memory and are replicated by ``tvbf_peer_push`` (NVLink stores
    + a device-side barrier; two sets used in turn, so a result stays valid until the job after the
    next one).  Without symmetric memory (or ``TVBF_PEER_GATHER=0``): one coalesced NCCL all-gather;
    ``tables`` then names the padded gather buffers of an earlier call (``out["_full"]``) to reuse.
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = cat.n_shows
    dbg = os.environ.get("TVBF_HOST_TIMES") and rank == 0
    with torch.cuda.device(eng.device):
        t0 = time.perf_counter()
        ptab = peer_tables(eng, n, k, group)
        if ptab is not None:
            full, ptrs, barrier = ptab.next()
        else:
            full = tables if tables is not None and tuple(tables["indices"].shape) == (world * shard_rows(n, world), k) \
                else alloc_full_tables(n, k, world, eng.device)
        mine = shard_views(full, n, world, rank)
        t1 = time.perf_counter()
        _local_tables(eng, cat, weights, k, min_similarity, exclude_self, group, symmetric, splits, tuning, events, mine)
        t2 = time.perf_counter()
        if ptab is not None:
            # this rank's shard of every field -> the seven other copies over NVLink, then a device-side barrier
            check(eng.lib.tvbf_peer_push(ptrs, world, rank, ptab.offsets, ptab.sizes, ptab.n_fields, eng._stream()),
                  "tvbf_peer_push")
            barrier()
            out = {name: full[name][:n] for name in ("indices", "counts", "hybrid", "genre", "text", "metadata")}
            out["stats"] = full["stats"]
        else:
            out = gather_full_tables(full, n, group)
        if dbg:
            print(f"host: alloc {1e3 * (t1 - t0):.2f} local {1e3 * (t2 - t1):.2f} gather {1e3 * (time.perf_counter() - t2):.2f} ms",
                  flush=True)
        if events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            events.setdefault("gather1", []).append(ev)
    out["row_begin"] = 0
    out["_full"] = full
    return out


class DistributedTopK:
    """End-to-end multi-GPU job with its host-side resources (a shared pinned result table) set up
    once and reused per call: ``run(staged)`` -> full host ``TopK`` on every rank."""

    def __init__(self, eng: HybridTopKEngine, n_shows: int, k: int, group=None):
        self.eng, self.group, self.n, self.k = eng, group, n_shows, k
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.upload = ShardedUpload(group)
        self.table = SharedHostTable(n_shows, k, group)
        self._prev = None
        self._full = None          # padded device tables, reused from job to job

    def h2d(self, st: StagedCatalogue) -> dict:
        """1/world of every feature buffer over this rank's PCIe link + one all-gather over NVLink."""
        dev = self.eng.device
        with torch.cuda.device(dev):
            host = [st.text_indptr, st.text_indices, st.text_values, st.genre, *st.meta]
            d = self.upload(host, dev)
        return {"st": st, "indptr": d[0], "indices": d[1], "values": d[2], "genre": d[3], "meta": d[4:]}

    def run(self, st: StagedCatalogue, weights=(0.4, 0.5, 0.1), min_similarity: float = 0.1,
            exclude_self: bool = True, symmetric: bool | None = None, splits: int = 0, tuning: int = 0) -> TopK:
        eng, n, k = self.eng, self.n, self.k
        assert st.n_shows == n
        with torch.cuda.device(eng.device):
            cat = eng.prepare(self.h2d(st), weights, recycle=self._prev)
            self._prev = cat
            if self._full is None:
                self._full = alloc_full_tables(n, k, self.world, eng.device)
            full = self._full
            mine = shard_views(full, n, self.world, self.rank)
            _local_tables(eng, cat, weights, k, min_similarity, exclude_self, self.group, symmetric, splits, tuning,
                          None, mine)
            self.table.store_shard(mine)
            torch.cuda.current_stream(eng.device).synchronize()
        dist.barrier(group=self.group)          # every shard has landed in the shared table
        h = self.table.numpy()
        stats = h["stats"].sum(axis=0)
        return TopK(indices=h["indices"], counts=h["counts"], hybrid=h["hybrid"], genre=h["genre"], text=h["text"],
                    metadata=h["metadata"], row_begin=0, flagged_rows=int(stats[0]), rescored_pairs=int(stats[1]))

    def close(self) -> None:
        self.table.close()


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
