"""Shared helpers of the GPU parity tests: run the oracle (CPU, float64) on the same inputs and
compare with the tie-aware comparator."""

from __future__ import annotations

import numpy as np

from oracle.compare import compare_topk
from oracle.reference_paths import ProductionRows


def oracle_rows(features, rows, weights=(0.4, 0.5, 0.1), k=20, min_similarity=0.1,
                metadata_mode="mean3", normalize_weights=False):
    pr = ProductionRows(features, *weights, metadata_mode=metadata_mode, normalize_weights=normalize_weights)
    idx, cnt, sc = pr.topk_arrays(rows, k, min_similarity)
    return pr, idx, cnt, sc


def assert_topk_matches(top, features, rows=None, weights=(0.4, 0.5, 0.1), k=20, min_similarity=0.1,
                        metadata_mode="mean3", normalize_weights=False, eps=1e-9, rtol=1e-5):
    """``top``: engine TopK covering all rows of ``features`` (row_begin may be non-zero)."""
    n = top.indices.shape[0]
    rows = np.arange(top.row_begin, top.row_begin + n) if rows is None else np.asarray(rows)
    pr, ridx, rcnt, rsc = oracle_rows(features, rows, weights, k, min_similarity, metadata_mode,
                                      normalize_weights)
    local = rows - top.row_begin
    rep = compare_topk(ridx, rcnt, rsc[0], top.indices[local], top.counts[local], top.hybrid[local],
                       lambda r, js: pr.pair_scores(int(rows[r]), js), k, min_similarity, eps=eps, rtol=rtol)
    assert rep.ok, rep.summary() + "\n" + "\n".join(rep.failures)
    # component scores of entries at identical positions
    same = (ridx == top.indices[local])
    for c, name in enumerate(("hybrid", "genre", "text", "metadata")):
        got = getattr(top, name)[local]
        ref = rsc[c]
        m = same & (ridx >= 0)
        if m.any():
            err = np.abs(got[m] - ref[m])
            assert np.all(err <= rtol * np.abs(ref[m]) + 1e-12), f"{name}: max err {err.max()}"
    return rep
