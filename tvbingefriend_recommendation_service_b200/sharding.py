"""Row-block sharding of the top-K job across the GPUs of one box, and the collectives around it.

Rows are independent (each row's top-K needs row i's features and ALL columns' features,
nothing else -- the reference computes them one at a time, scripts/populate_database.py:170), so
the features are replicated on every GPU and each rank owns a contiguous block of source rows.
All shards have the SAME padded size (whole 128-row tiles), so every collective is the plain
equal-split form and the gathers run in place:

* ``exchange_packed``   one ``all_to_all_single`` of the packed candidate rows (symmetric sweep);
* ``gather_tables``     the ``[rows, k]`` result tables, all fields in ONE coalesced in-place
                        all-gather (NCCL; one call per field on backends without coalescing);
* ``ShardedUpload``     each rank uploads 1/world of the feature bytes over its own PCIe link and
                        the GPUs all-gather them over NVLink (instead of ``world`` full uploads);
* ``SharedHostTable``   one pinned host table in shared memory: every rank copies its shard D2H
                        over its own PCIe link (instead of ``world`` x the table through rank 0's).

NCCL over NVLink on GPUs; gloo in the CPU tests of this logic.
"""

from __future__ import annotations

import mmap
import os

import numpy as np
import torch
import torch.distributed as dist

ROW_TILE = 128
_FIELDS = ("indices", "counts", "hybrid", "genre", "text", "metadata")


def shard_rows(n_shows: int, world_size: int) -> int:
    """Rows per shard: whole 128-row tiles, the same for every rank (the last shards may be short
    or empty)."""
    tiles = (n_shows + ROW_TILE - 1) // ROW_TILE
    return (tiles + world_size - 1) // world_size * ROW_TILE


def row_shard(n_shows: int, world_size: int, rank: int) -> tuple[int, int]:
    """[row_begin, row_end) of ``rank``; ranks past the last tile get an empty range."""
    rows = shard_rows(n_shows, world_size)
    b = min(rank * rows, n_shows)
    return b, min(b + rows, n_shows)


def max_shard_rows(n_shows: int, world_size: int) -> int:
    return shard_rows(n_shows, world_size)


def _coalescing(group, device):
    """One NCCL launch for several collectives; a no-op context on backends without support."""
    backend = dist.get_backend(group)
    if backend == "nccl" and hasattr(dist, "_coalescing_manager"):
        return dist._coalescing_manager(group=group, device=device, async_ops=False)

    class _Null:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    return _Null()


def exchange_row_shards(t: torch.Tensor, n_shows: int, group=None) -> torch.Tensor:
    """All-to-all over the row shards: every rank holds per-show entries ``t[>= N, ...]`` for ALL
    shows (its partial candidate lists); rank r receives every rank's entries for ITS rows and
    returns ``[world, shard_rows, ...]`` (rows past the catalogue are padding).  Moves 1/world of
    the bytes an all-gather would.  ``t`` may already be padded to ``world * shard_rows`` rows
    (no copy then)."""
    world = dist.get_world_size(group)
    rows = shard_rows(n_shows, world)
    total = world * rows
    if t.shape[0] != total:
        padded = torch.empty((total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        padded[:t.shape[0]] = t
        t = padded
    out = torch.empty((world, rows) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_to_all_single(out.view(-1), t.contiguous().view(-1), group=group)
    return out


def exchange_packed(packed: torch.Tensor, group=None) -> torch.Tensor:
    """The symmetric sweep's exchange: ``packed[world * shard_rows, L + 1, 2]`` int32 (candidate
    entries + {count, bound} per show) -> ``[world, shard_rows, L + 1, 2]``: every rank's lists of
    this rank's rows, in ONE collective."""
    world = dist.get_world_size(group)
    rows = packed.shape[0] // world
    out = torch.empty((world, rows) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    dist.all_to_all_single(out.view(-1), packed.view(-1), group=group)
    return out


class PeerCandidateBuffers:
    """Receive buffers of the FUSED candidate exchange: ``[world, shard_rows, L + 1, 2]`` int32 on every
    GPU, allocated as torch symmetric memory so that each rank holds a device mapping of every peer's
    buffer.  The compaction kernel (K4s) of rank r stores each finished candidate row straight into
    slice r of the owner's buffer over NVLink (``tvbf_sym_sweep_peer``); ``barrier()`` (a device-side
    signal exchange on the current stream) replaces the all-to-all.  Two buffers are used in turn: a
    fast rank may already be writing the next job's rows while a slow one still rescoring this job's."""

    def __init__(self, n_shows: int, list_len: int, device, group=None):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.rows = shard_rows(n_shows, self.world)
        self.key = (n_shows, list_len, self.world)
        grp = group if group is not None else dist.group.WORLD
        self.bufs, self.handles, self.ptrs = [], [], []
        for _ in range(2):
            t = symm_mem.empty((self.world * self.rows, list_len + 1, 2), dtype=torch.int32, device=device)
            h = symm_mem.rendezvous(t, grp)
            ptrs = [int(x) for x in h.buffer_ptrs]
            if len(ptrs) != self.world or any(x == 0 for x in ptrs):
                raise RuntimeError("symmetric memory rendezvous returned no peer mappings")
            self.bufs.append(t)
            self.handles.append(h)
            self.ptrs.append((C.c_uint64 * self.world)(*ptrs))
        self.turn = 0

    def next(self):
        """(peer pointer array for tvbf_sym_sweep_peer, local receive view [world, rows, L + 1, 2],
        barrier callable) of the next job."""
        i = self.turn & 1
        self.turn += 1
        buf, h = self.bufs[i], self.handles[i]
        return self.ptrs[i], buf.view(self.world, self.rows, buf.shape[1], 2), (lambda: h.barrier(channel=i))


def peer_table_layout(n_shows: int, k: int, world: int) -> tuple[list, int]:
    """Byte layout of the result tables inside ONE buffer: ``[(name, dtype, shape, offset, nbytes,
    shard_bytes)]`` and the total size.  Every field is padded to ``world * shard_rows`` rows, so rank
    r's rows of a field are the ``shard_bytes`` at ``offset + r * shard_bytes``; offsets are multiples
    of 256."""
    rows = shard_rows(n_shows, world)
    total = world * rows
    specs = [("indices", torch.int32, (total, k), rows * k * 4), ("counts", torch.int32, (total,), rows * 4)]
    specs += [(name, torch.float64, (total, k), rows * k * 8) for name in ("hybrid", "genre", "text", "metadata")]
    specs += [("stats", torch.int32, (world, 8), 32)]
    off, layout = 0, []
    for name, dt, shape, shard_bytes in specs:
        nbytes = shard_bytes * world
        layout.append((name, dt, shape, off, nbytes, shard_bytes))
        off = (off + nbytes + 255) // 256 * 256
    return layout, off


class PeerTables:
    """The ``[world * shard_rows, k]`` result tables of ``alloc_full_tables`` in torch symmetric memory
    (one allocation, the fields back to back), so that every rank holds a device mapping of every
    peer's copy.  After the kernels have written a rank's row shard into ITS copy, ``tvbf_peer_push``
    stores that shard into the seven other copies over NVLink and ``barrier()`` (device-side, on the
    current stream) makes all shards visible everywhere: the all-gather without NCCL's ring.  Two sets
    are used in turn -- a fast rank may already be pushing the next job's rows."""

    def __init__(self, n_shows: int, k: int, device, group=None):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_shows, self.k = n_shows, k
        layout, self.nbytes = peer_table_layout(n_shows, k, self.world)
        grp = group if group is not None else dist.group.WORLD
        self.sets = []
        for _ in range(2):
            t = symm_mem.empty((self.nbytes,), dtype=torch.uint8, device=device)
            h = symm_mem.rendezvous(t, grp)
            ptrs = [int(x) for x in h.buffer_ptrs]
            if len(ptrs) != self.world or any(x == 0 for x in ptrs):
                raise RuntimeError("symmetric memory rendezvous returned no peer mappings")
            views = {name: t[o:o + nb].view(dt).view(shape) for name, dt, shape, o, nb, _sb in layout}
            self.sets.append((t, h, (C.c_uint64 * self.world)(*ptrs), views))
        # this rank's slice of every field: what it pushes to the peers
        self.n_fields = len(layout)
        self.offsets = (C.c_uint64 * self.n_fields)(*[o + self.rank * sb for _n, _d, _s, o, _nb, sb in layout])
        self.sizes = (C.c_uint64 * self.n_fields)(*[sb for *_x, sb in layout])
        self.turn = 0

    def next(self):
        """(tables dict like ``alloc_full_tables``, peer pointer array, barrier callable) of the next job."""
        i = self.turn & 1
        self.turn += 1
        _t, h, ptrs, views = self.sets[i]
        return views, ptrs, (lambda: h.barrier(channel=i))


def empty_tables(k: int, device) -> dict:
    """Local tables of a rank that owns no rows."""
    t = {"indices": torch.empty((0, k), dtype=torch.int32, device=device),
         "counts": torch.empty((0,), dtype=torch.int32, device=device),
         "stats": torch.zeros((8,), dtype=torch.int32, device=device)}
    for name in ("hybrid", "genre", "text", "metadata"):
        t[name] = torch.empty((0, k), dtype=torch.float64, device=device)
    return t


def alloc_full_tables(n_shows: int, k: int, world: int, device) -> dict:
    """Result tables padded to ``world * shard_rows`` rows, so that a rank's shard is a contiguous
    slice the kernels write in place and the all-gather needs no staging; ``stats`` is [world, 8]."""
    total = world * shard_rows(n_shows, world)
    t = {"indices": torch.empty((total, k), dtype=torch.int32, device=device),
         "counts": torch.empty((total,), dtype=torch.int32, device=device),
         "stats": torch.empty((world, 8), dtype=torch.int32, device=device)}
    for name in ("hybrid", "genre", "text", "metadata"):
        t[name] = torch.empty((total, k), dtype=torch.float64, device=device)
    return t


def shard_views(full: dict, n_shows: int, world: int, rank: int) -> dict:
    """This rank's slice of ``alloc_full_tables`` (views: the kernels write straight into the
    gather buffer)."""
    rows = shard_rows(n_shows, world)
    v = {name: full[name][rank * rows:(rank + 1) * rows] for name in _FIELDS}
    v["stats"] = full["stats"][rank]
    return v


def gather_full_tables(full: dict, n_shows: int, group=None) -> dict:
    """In-place all-gather of every field of ``alloc_full_tables`` (each rank has filled its own
    slice): one coalesced NCCL launch.  Returns views trimmed to ``n_shows`` rows; ``stats`` stays
    [world, 8] (summed by the reader)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    rows = full["counts"].shape[0] // world
    with _coalescing(group, full["counts"].device):
        for name in _FIELDS:
            dist.all_gather_into_tensor(full[name], full[name][rank * rows:(rank + 1) * rows], group=group)
        dist.all_gather_into_tensor(full["stats"].view(-1), full["stats"][rank], group=group)
    out = {name: full[name][:n_shows] for name in _FIELDS}
    out["stats"] = full["stats"]
    return out


def gather_tables(local: dict, n_shows: int, k: int, group=None) -> dict:
    """All-gather per-rank tables (``local`` holds this rank's ``row_shard`` rows) into full
    ``[n_shows, k]`` tensors on every rank.  Convenience form that copies ``local`` into the padded
    gather buffer first; the drivers write into ``shard_views`` directly instead."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = local["counts"].device
    full = alloc_full_tables(n_shows, k, world, dev)
    mine = shard_views(full, n_shows, world, rank)
    b, e = row_shard(n_shows, world, rank)
    for name in _FIELDS:
        mine[name][: e - b] = local[name]
    stats = local.get("stats")
    mine["stats"].copy_(stats if stats is not None else torch.zeros(8, dtype=torch.int32, device=dev))
    out = gather_full_tables(full, n_shows, group)
    out["stats"] = out["stats"].sum(dim=0).to(torch.int32)
    return out


def tables_to_numpy(t: dict) -> dict:
    return {n: (v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for n, v in t.items()
            if not n.startswith("_")}


# ---- features: each rank uploads a slice, NVLink replicates -------------------------------------------
class ShardedUpload:
    """Host -> device of a list of (pinned) host tensors, replicated on every GPU of the group:
    rank r copies only bytes [r/world, (r+1)/world) of each tensor over its PCIe link and one
    coalesced in-place all-gather over NVLink completes every copy."""

    ALIGN = 256

    def __init__(self, group=None):
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self._cache: dict = {}     # device buffers of earlier calls (two sets used in turn), reused when the sizes repeat
        self._gen = 0

    def __call__(self, host_tensors: list[torch.Tensor], device) -> list[torch.Tensor]:
        """The returned tensors alias buffers that the call AFTER the next one overwrites (two sets are
        used in turn: the previous job's CSR arrays are still read when its operand is recycled)."""
        world, rank = self.world, self.rank
        bufs, outs = [], []
        self._gen ^= 1
        for j, h in enumerate(host_tensors):
            i = (self._gen, j)
            nbytes = h.numel() * h.element_size()
            chunk = (nbytes + world * self.ALIGN - 1) // (world * self.ALIGN) * self.ALIGN
            dev = self._cache.get(i)
            if dev is None or dev.numel() != world * chunk or dev.device != torch.device(device):
                dev = self._cache[i] = torch.empty((world * chunk,), dtype=torch.uint8, device=device)
            src = h.view(-1).view(torch.uint8)
            b, e = min(rank * chunk, nbytes), min((rank + 1) * chunk, nbytes)
            if e > b:
                dev[b:e].copy_(src[b:e], non_blocking=True)
            bufs.append((dev, chunk))
            outs.append(dev[:nbytes].view(h.dtype).view(h.shape))
        with _coalescing(self.group, device):
            for dev, chunk in bufs:
                dist.all_gather_into_tensor(dev, dev[rank * chunk:(rank + 1) * chunk], group=self.group)
        return outs


# ---- result: one pinned host table shared by the ranks -----------------------------------------------
class SharedHostTable:
    """``[world * shard_rows, k]`` result table in POSIX shared memory, registered as pinned memory
    by every rank: after the kernels each rank copies ITS shard device -> host over its own PCIe
    link; a barrier later every rank can read the whole table as numpy views."""

    def __init__(self, n_shows: int, k: int, group=None):
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_shows, self.k = n_shows, k
        self.rows = shard_rows(n_shows, self.world)
        total = self.world * self.rows
        specs = [("indices", np.int32, (total, k)), ("counts", np.int32, (total,)),
                 ("hybrid", np.float64, (total, k)), ("genre", np.float64, (total, k)),
                 ("text", np.float64, (total, k)), ("metadata", np.float64, (total, k)),
                 ("stats", np.int32, (self.world, 8))]
        off, self.layout = 0, {}
        for name, dt, shape in specs:
            nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
            self.layout[name] = (off, dt, shape)
            off = (off + nbytes + 4095) // 4096 * 4096
        self.nbytes = off
        names = [None]
        if self.rank == 0:
            names[0] = f"/dev/shm/tvbf_table_{os.getpid()}_{id(self) & 0xffff:x}"
            with open(names[0], "wb") as fh:
                fh.truncate(self.nbytes)
        dist.broadcast_object_list(names, src=0, group=group)
        self.path = names[0]
        self._fh = open(self.path, "r+b")
        self._map = mmap.mmap(self._fh.fileno(), self.nbytes)
        self._all = np.frombuffer(self._map, dtype=np.uint8)
        self._tensor = torch.from_numpy(self._all)
        self._registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self._tensor.data_ptr(), self.nbytes, 0)
            self._registered = int(rc) == 0
        dist.barrier(group=group)
        if self.rank == 0:
            os.unlink(self.path)       # the mappings keep it alive; nothing is left behind on exit

    def field(self, name: str) -> torch.Tensor:
        off, dt, shape = self.layout[name]
        n = int(np.prod(shape)) * np.dtype(dt).itemsize
        return self._tensor[off:off + n].view(getattr(torch, np.dtype(dt).name)).view(shape)

    def store_shard(self, local: dict) -> None:
        """Enqueue the D2H copies of this rank's shard (``local`` = ``shard_views`` tensors or tables
        of exactly the shard's rows) on the current stream."""
        r0 = self.rank * self.rows
        for name in _FIELDS:
            src = local[name]
            self.field(name)[r0:r0 + src.shape[0]].copy_(src, non_blocking=True)
        self.field("stats")[self.rank].copy_(local["stats"], non_blocking=True)

    def numpy(self) -> dict:
        out = {name: self.field(name).numpy()[: self.n_shows] for name in _FIELDS}
        out["stats"] = self.field("stats").numpy()
        return out

    def close(self) -> None:
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self._tensor.data_ptr())
            self._registered = False
