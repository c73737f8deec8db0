"""Tie-aware comparison of a top-K table against the reference's.  Test infrastructure only.

Why "tie-aware": the reference orders a row with ``np.argsort(hybrid_sim)[::-1]``
(scripts/populate_database.py:195; services/content_based_service.py:209), numpy's default
*unstable* sort, so the relative order of equal scores -- and which of several equal scores
survives the rank-K cut -- is implementation-defined.  Exact float64 ties are common here
because genre and metadata scores are discrete.  Everything that IS defined is checked exactly:

* wherever a reference score is separated from its neighbours by more than ``eps`` the index at
  that rank must be identical;
* inside a run of (near-)equal reference scores the same SET of indices must occupy the run;
* a run that touches the rank-K cut (or the ``min_similarity`` cut) may differ only by members
  whose reference score is within ``eps`` of the cut value;
* scores of common indices agree to ``rtol`` relative;
* the candidate's own list is sorted by (score descending, index ascending) -- the tie-break
  this repo states.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable

import numpy as np


@dataclass
class Report:
    rows: int = 0
    rows_identical_ordered: int = 0   # index list identical position by position
    rows_tie_permuted: int = 0        # differs only inside reference tie runs / at a tied cut
    max_rel_score_err: float = 0.0
    failures: list = field(default_factory=list)

    @property
    def ok(self) -> bool:
        return not self.failures

    def summary(self) -> str:
        return (f"rows={self.rows} identical={self.rows_identical_ordered} "
                f"tie_permuted={self.rows_tie_permuted} failures={len(self.failures)} "
                f"max_rel_score_err={self.max_rel_score_err:.3e}")


def compare_topk(ref_idx, ref_cnt, ref_score, got_idx, got_cnt, got_score,
                 pair_score: Callable[[int, np.ndarray], np.ndarray], k: int,
                 min_similarity: float, eps: float = 1e-9, rtol: float = 1e-5,
                 max_failures: int = 20) -> Report:
    """``*_idx`` [R,k] (-1 padded), ``*_cnt`` [R], ``*_score`` [R,k] hybrid scores.

    ``pair_score(r, js)`` returns the REFERENCE hybrid score of source row ``r`` (position in
    these arrays) against column indices ``js`` -- needed for candidates the reference list does
    not contain (legal only when tied with the cut).
    """
    rep = Report()
    R = ref_idx.shape[0]
    for r in range(R):
        rep.rows += 1
        rc, gc = int(ref_cnt[r]), int(got_cnt[r])
        ri, gi = ref_idx[r, :rc].astype(np.int64), got_idx[r, :gc].astype(np.int64)
        rs, gs = ref_score[r, :rc].astype(np.float64), got_score[r, :gc].astype(np.float64)

        def fail(msg):
            if len(rep.failures) < max_failures:
                rep.failures.append(f"row {r}: {msg}")

        if gc > k or np.any(gi < 0) or len(set(gi.tolist())) != gc:
            fail(f"malformed candidate list (count {gc}, ids {gi.tolist()})")
            continue
        # own ordering: score desc, index asc on exact ties
        bad_order = False
        for c in range(1, gc):
            if gs[c] > gs[c - 1] or (gs[c] == gs[c - 1] and gi[c] < gi[c - 1]):
                bad_order = True
        if bad_order:
            fail("candidate list not sorted by (score desc, index asc)")
            continue
        if np.any(gs < min_similarity - eps):
            fail("candidate score below min_similarity")
            continue

        # the cut value: k-th reference score when the list is full, else the threshold
        cut = rs[rc - 1] if rc == k else min_similarity
        ref_map = {int(j): float(s) for j, s in zip(ri.tolist(), rs.tolist())}
        extra = [int(j) for j in gi.tolist() if int(j) not in ref_map]
        extra_scores = pair_score(r, np.asarray(extra, dtype=np.int64)) if extra else np.zeros(0)
        for j, s in zip(extra, np.asarray(extra_scores, dtype=np.float64).tolist()):
            ref_map[j] = s
        # (2) every reference member clearly above the cut is present
        gset = set(gi.tolist())
        missing = [int(j) for j, s in zip(ri.tolist(), rs.tolist()) if s > cut + eps and int(j) not in gset]
        if missing:
            fail(f"missing reference members above the cut: {missing}")
            continue
        # (3) every candidate member has reference score >= cut - eps
        low = [j for j in gi.tolist() if ref_map[int(j)] < cut - eps]
        if low:
            fail(f"members below the reference cut {cut!r}: {low}")
            continue
        # (4) list length
        if rc == k:
            n_clear = int(np.sum(rs > min_similarity + eps))
            if gc < min(k, n_clear):
                fail(f"count {gc} < {min(k, n_clear)}")
                continue
        else:
            n_clear = int(np.sum(rs > min_similarity + eps))
            n_near = sum(1 for j in gi.tolist() if abs(ref_map[int(j)] - min_similarity) <= eps)
            if gc < n_clear or gc > rc + n_near:
                fail(f"count {gc} vs reference {rc}")
                continue
        # (5) scores of every candidate member against the reference's value for that pair
        for j, s in zip(gi.tolist(), gs.tolist()):
            refv = ref_map[int(j)]
            err = abs(s - refv) / max(abs(refv), 1e-300)
            if abs(s - refv) > rtol * abs(refv) + 1e-12:
                fail(f"score of {j}: got {s!r} reference {refv!r}")
                break
            rep.max_rel_score_err = max(rep.max_rel_score_err, err if refv != 0 else 0.0)
        else:
            # (6) order: exact where the reference order is defined, set-equal inside tie runs
            identical = rc == gc and bool(np.array_equal(ri, gi))
            if identical:
                rep.rows_identical_ordered += 1
                continue
            ok = True
            c = 0
            n = min(rc, gc)
            while c < n:
                e = c
                while e + 1 < rc and rs[e] - rs[e + 1] <= eps:
                    e += 1
                run_ref = set(ri[c:e + 1].tolist())
                touches_cut = (e == rc - 1)
                run_got = set(gi[c:min(e + 1, gc)].tolist())
                if not touches_cut:
                    if run_ref != run_got:
                        ok = False
                        fail(f"ranks {c}..{e}: reference {sorted(run_ref)} got {sorted(run_got)}")
                        break
                else:
                    # members may be swapped for others tied with the cut (checked in 2/3)
                    for j in run_got - run_ref:
                        if abs(ref_map[int(j)] - rs[e]) > eps and abs(ref_map[int(j)] - cut) > eps:
                            ok = False
                            fail(f"rank {c}..{e}: {j} is not tied with the cut")
                            break
                    if not ok:
                        break
                c = e + 1
            if ok:
                rep.rows_tie_permuted += 1
            continue
        continue
    return rep
