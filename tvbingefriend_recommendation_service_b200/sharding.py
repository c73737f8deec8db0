"""Row-block sharding of the top-K job across the GPUs of one box.

Rows are independent (each row's top-K needs row i's features and ALL columns' features,
nothing else -- the reference computes them one at a time, scripts/populate_database.py:170), so
the features are replicated on every GPU, each rank owns a contiguous block of source rows
(multiples of the 128-row tile), and the only exchange is one all-gather of the small
``[rows, k]`` result tables (NCCL over NVLink on GPUs; gloo in the CPU tests of this logic).
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

ROW_TILE = 128
_FIELDS = ("indices", "counts", "hybrid", "genre", "text", "metadata")


def row_shard(n_shows: int, world_size: int, rank: int) -> tuple[int, int]:
    """[row_begin, row_end) of ``rank``: whole 128-row tiles, as even as possible; ranks past the
    last tile get an empty range."""
    tiles = (n_shows + ROW_TILE - 1) // ROW_TILE
    b = tiles * rank // world_size
    e = tiles * (rank + 1) // world_size
    return min(b * ROW_TILE, n_shows), min(e * ROW_TILE, n_shows)


def max_shard_rows(n_shows: int, world_size: int) -> int:
    return max(row_shard(n_shows, world_size, r)[1] - row_shard(n_shows, world_size, r)[0]
               for r in range(world_size))


def exchange_row_shards(t: torch.Tensor, n_shows: int, group=None) -> torch.Tensor:
    """All-to-all over the row shards: every rank holds per-show entries ``t[N, ...]`` for ALL
    shows (its partial candidate lists); rank r receives every rank's entries for ITS rows and
    returns ``[world, rows_r, ...]``.  Moves 1/world of the bytes an all-gather would."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    spans = [row_shard(n_shows, world, r) for r in range(world)]
    rows = spans[rank][1] - spans[rank][0]
    per_row = 1
    for d in t.shape[1:]:
        per_row *= int(d)
    out = torch.empty((world, rows) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_to_all_single(out.view(-1), t.contiguous().view(-1), output_split_sizes=[rows * per_row] * world,
                           input_split_sizes=[(e - b) * per_row for b, e in spans], group=group)
    return out


def empty_tables(k: int, device) -> dict:
    """Local tables of a rank that owns no rows."""
    t = {"indices": torch.empty((0, k), dtype=torch.int32, device=device),
         "counts": torch.empty((0,), dtype=torch.int32, device=device),
         "stats": torch.zeros((8,), dtype=torch.int32, device=device)}
    for name in ("hybrid", "genre", "text", "metadata"):
        t[name] = torch.empty((0, k), dtype=torch.float64, device=device)
    return t


def gather_tables(local: dict, n_shows: int, k: int, group=None) -> dict:
    """All-gather the per-rank tables into full ``[n_shows, k]`` tensors (on every rank).

    ``local`` holds this rank's tensors for its ``row_shard`` (device or CPU tensors; the
    collective runs on whatever backend the group has).  Shards are padded to a common row count
    so that one all_gather per field suffices."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    pad_rows = max_shard_rows(n_shows, world)
    b, e = row_shard(n_shows, world, rank)
    out = {}
    for name in _FIELDS:
        t = local[name]
        shape = (pad_rows,) + tuple(t.shape[1:])
        padded = torch.zeros(shape, dtype=t.dtype, device=t.device)
        padded[: e - b] = t
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        full = []
        for r in range(world):
            rb, re_ = row_shard(n_shows, world, r)
            full.append(parts[r][: re_ - rb])
        out[name] = torch.cat(full, dim=0)
        assert out[name].shape[0] == n_shows
    stats = local.get("stats")
    if stats is not None:
        s = stats.clone()
        dist.all_reduce(s, group=group)
        out["stats"] = s
    return out


def tables_to_numpy(t: dict) -> dict:
    return {n: (v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for n, v in t.items()}
