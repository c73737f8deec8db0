"""Randomised cross-check: every candidate path (one-sided / symmetric sweep, column splits, k,
thresholds, weights, metadata conventions, sizes down to one show) against the exact fp64 kernel
(bit-for-bit) and, on a few rows, against the numpy oracle (tie-aware)."""

import numpy as np
import pytest

from helpers import assert_topk_matches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["popcount", "folded"])
def engine(request):
    """Every test of this module runs twice: with the popcount epilogue (what large vocabularies get)
    and with the packed groups folded into the operand (what these small test vocabularies get by
    default: ``HybridTopKEngine.fold_max_k``)."""
    from tvbingefriend_recommendation_service_b200.engine import HybridTopKEngine

    eng = HybridTopKEngine(0)
    if request.param == "popcount":
        eng.fold_max_k = 0
    return eng


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_jobs_agree_with_the_exact_kernel_and_the_oracle(engine, seed):
    from tvbingefriend_recommendation_service_b200._lib import TvbfError
    from tvbingefriend_recommendation_service_b200.engine import stage
    from tvbingefriend_recommendation_service_b200.synthetic import WEIGHT_SWEEP, make_catalogue

    rng = np.random.default_rng(seed)
    ran = 0
    for case in range(25):
        n = int(rng.choice([1, 2, 100, 257, 700, 1500, 3000, 6000, 9000]))
        v = int(rng.choice([64, 300, 1024, 3000]))
        nnz = min(int(rng.choice([3, 10, 30])), v)
        k = int(rng.choice([1, 5, 20, 48, 64, 100]))
        ms = float(rng.choice([0.02, 0.1, 0.3]))
        w = WEIGHT_SWEEP[int(rng.integers(len(WEIGHT_SWEEP)))] if rng.random() < 0.6 \
            else tuple(float(x) for x in rng.uniform(0.05, 2.0, 3).round(3))
        mode = str(rng.choice(["mean3", "hstack"]))
        sym = int(rng.choice([0, 1, 2]))          # auto / one-sided / symmetric
        splits = int(rng.choice([0, 1, 2, 3, 5]))
        cat = make_catalogue(n, v, nnz=nnz, seed=int(rng.integers(1 << 30)))
        f = cat.features()
        dc = engine.upload(stage(f, mode), w)
        tag = f"seed {seed} case {case}: n={n} v={v} nnz={nnz} k={k} ms={ms} w={w} {mode} sym={sym} splits={splits}"
        try:
            got = engine.to_host(engine.top_k_device(dc, w, k, ms, splits=splits, tuning=sym << 20))
        except TvbfError:
            assert sym == 2, tag       # only an explicit symmetric request may be refused
            continue
        ref = engine.to_host(engine.top_k_device(dc, w, k, ms, force_exact=True))
        assert np.array_equal(got.indices, ref.indices) and np.array_equal(got.counts, ref.counts), tag
        m = ref.indices >= 0
        for name in ("hybrid", "genre", "text", "metadata"):
            assert np.array_equal(getattr(got, name)[m], getattr(ref, name)[m]), (tag, name)
        rows = np.unique(rng.integers(0, n, size=min(n, 12)))
        assert_topk_matches(got, f, rows, w, k, ms, metadata_mode=mode)
        ran += 1
    assert ran >= 15
