"""Parity at BASELINE.json sizes (``-m gpu``): the oracle is run on a row sample (it needs ~20 ms
per source row at N = 100 k), and the whole table is checked through size-independent properties
of the selection rule (populate_database.py:195-218): sorted by (score desc, index asc), self
excluded, scores >= min_similarity, count < k only when the threshold cuts, determinism, and
agreement of the certified tensor-core path with the exact fp64 row kernel on flagged rows."""

import numpy as np
import pytest

from helpers import assert_topk_matches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from tvbingefriend_recommendation_service_b200.engine import HybridTopKEngine

    return HybridTopKEngine(0)


def check_table_properties(top, n, k, min_similarity):
    idx, cnt, h = top.indices, top.counts, top.hybrid
    assert idx.shape == (n, k) and cnt.shape == (n,)
    valid = np.arange(k)[None, :] < cnt[:, None]
    assert (idx[valid] >= 0).all() and (idx[valid] < n).all() and (idx[~valid] == -1).all()
    assert np.isnan(h[~valid]).all() and not np.isnan(h[valid]).any()
    assert (h[valid] >= min_similarity).all()
    rows = np.broadcast_to(np.arange(n)[:, None], (n, k))
    assert (idx[valid] != rows[valid]).all()                       # self excluded
    both = valid[:, 1:] & valid[:, :-1]
    d = h[:, :-1] - h[:, 1:]
    assert (d[both] >= 0).all()                                    # descending
    tie = both & (d == 0)
    assert (idx[:, 1:][tie] > idx[:, :-1][tie]).all()              # index ascending on exact ties
    srt = np.sort(np.where(valid, idx, -np.arange(1, k + 1)[None, :]), axis=1)
    assert (np.diff(srt, axis=1) != 0).all()                       # no duplicates
    for name in ("genre", "text", "metadata"):
        v = getattr(top, name)[valid]
        assert (v >= -1e-12).all() and (v <= 1 + 1e-9).all()
    recomb = 0.4 * top.genre[valid] + 0.5 * top.text[valid] + 0.1 * top.metadata[valid]
    assert np.abs(recomb - h[valid]).max() < 1e-12


def test_c2_20k_single_gpu_vs_oracle(engine):
    """BASELINE config 2: 20 k shows, 5 k vocab, top-20 on one B200 vs the CPU oracle."""
    from tvbingefriend_recommendation_service_b200.synthetic import make_config

    cat = make_config("C2")
    top = engine.compute_top_k(cat.features(), (0.4, 0.5, 0.1), 20, 0.1)
    check_table_properties(top, 20_000, 20, 0.1)
    rows = np.unique(np.concatenate([np.arange(0, 20_000, 41), np.arange(19_900, 20_000)]))
    rep = assert_topk_matches(top, cat.features(), rows)
    assert rep.rows == rows.size
    again = engine.compute_top_k(cat.features(), (0.4, 0.5, 0.1), 20, 0.1)
    assert np.array_equal(top.indices, again.indices) and np.array_equal(top.counts, again.counts)


@pytest.mark.parametrize("tuning", [2 << 20, 1 << 20], ids=["symmetric", "one_sided"])
def test_c3_100k_properties_and_sample(engine, tuning):
    """BASELINE config 3 at full size: 100 k shows, 10 k vocab."""
    from tvbingefriend_recommendation_service_b200.engine import stage
    from tvbingefriend_recommendation_service_b200.synthetic import make_config

    cat = make_config("C3")
    w = (0.4, 0.5, 0.1)
    dc = engine.upload(stage(cat.features()), w)
    top = engine.to_host(engine.top_k_device(dc, w, 20, 0.1, tuning=tuning))
    check_table_properties(top, 100_000, 20, 0.1)
    rows = np.linspace(0, 99_999, 96).astype(np.int64)
    assert_topk_matches(top, cat.features(), rows)
    # rows the certificate flagged were repaired by the exact kernel; un-flagged rows must agree
    # with the exact kernel too (spot check)
    ex = engine.exact_rows(dc, rows, w, k=20, min_similarity=0.1)
    assert np.array_equal(ex.indices, top.indices[rows]) and np.array_equal(ex.counts, top.counts[rows])
    m = ex.indices >= 0
    assert np.array_equal(ex.hybrid[m], top.hybrid[rows][m])
    # few listed rows: each row's columns are split over ~29 CTAs that share one survivor list
    few = rows[40:45]
    ex5 = engine.exact_rows(dc, few, w, k=20, min_similarity=0.1)
    assert np.array_equal(ex5.indices, top.indices[few]) and np.array_equal(ex5.hybrid, top.hybrid[few], equal_nan=True)
    assert 0 < top.flagged_rows < 20_000


def test_c4_250k_full_catalogue(engine):
    """BASELINE config 4: 250 k shows (full TVmaze-scale catalogue), N x N never materialised."""
    from tvbingefriend_recommendation_service_b200.engine import stage
    from tvbingefriend_recommendation_service_b200.synthetic import make_config

    cat = make_config("C4")
    w = (0.4, 0.5, 0.1)
    dc = engine.upload(stage(cat.features()), w)
    top = engine.to_host(engine.top_k_device(dc, w, 20, 0.1))
    check_table_properties(top, 250_000, 20, 0.1)
    rows = np.linspace(0, 249_999, 48).astype(np.int64)
    assert_topk_matches(top, cat.features(), rows)


def test_c5_200k_50k_vocab_top100_full_size(engine):
    """BASELINE config 5 at full size (200 k shows, 50 k vocabulary, top-100, weight sweep): 20 GB
    operand, GEMM K = 50 048, 128-candidate lists; the five triples share ONE symmetric sweep
    (one shared list per triple and show) and every table is checked against the oracle."""
    from tvbingefriend_recommendation_service_b200.engine import stage
    from tvbingefriend_recommendation_service_b200.synthetic import WEIGHT_SWEEP, make_config

    cat = make_config("C5")
    dc = engine.upload(stage(cat.features()), WEIGHT_SWEEP[0])
    tabs = engine.top_k_sweep_device(dc, WEIGHT_SWEEP, 100, 0.1)        # shared sweep (auto)
    rows = np.linspace(0, 199_999, 12).astype(np.int64)
    for w, t in zip(WEIGHT_SWEEP, tabs):
        top = engine.to_host(t)
        assert top.indices.shape == (200_000, 100)
        valid = np.arange(100)[None, :] < top.counts[:, None]
        assert (top.hybrid[valid] >= 0.1).all() and (np.diff(top.hybrid, axis=1)[valid[:, 1:]] <= 0).all()
        assert_topk_matches(top, cat.features(), rows, w, 100, 0.1)
    # a separate job for one triple gives the same table
    one = engine.to_host(engine.top_k_device(dc, WEIGHT_SWEEP[1], 100, 0.1))
    got = engine.to_host(tabs[1])
    assert np.array_equal(one.indices, got.indices) and np.array_equal(one.hybrid, got.hybrid, equal_nan=True)
    del dc, tabs
    engine.release()


def test_c5_shape_top100_sweep_on_reduced_rows(engine):
    """BASELINE config 5 shape (50 k vocab, top-100, weight sweep) on a 6 k-show slice: stresses
    GEMM K, the k=100 candidate lists and non-default weights."""
    from tvbingefriend_recommendation_service_b200.engine import stage
    from tvbingefriend_recommendation_service_b200.synthetic import WEIGHT_SWEEP, make_config

    cat = make_config("C5", n_shows=6_000)
    st = stage(cat.features())
    rows = np.arange(0, 6_000, 101)
    for w in WEIGHT_SWEEP:
        dc = engine.upload(st, w)
        top = engine.to_host(engine.top_k_device(dc, w, 100, 0.1))
        assert_topk_matches(top, cat.features(), rows, w, 100, 0.1)


def test_single_process_multi_gpu_matches_single_gpu(engine):
    """device_ids=[0, 1]: rows sharded in 128-row tiles over two GPUs, features replicated."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer
    from tvbingefriend_recommendation_service_b200.synthetic import make_catalogue

    cat = make_catalogue(3000, 2000, nnz=25, seed=77)
    one = SimilarityComputer(engine=engine).compute_top_k(cat.features())
    two = SimilarityComputer(engine=engine).compute_top_k(cat.features(), device_ids=[0, 1])
    assert np.array_equal(one.indices, two.indices) and np.array_equal(one.counts, two.counts)
    m = one.indices >= 0
    assert np.array_equal(one.hybrid[m], two.hybrid[m])


@pytest.mark.parametrize("world,k", [(2, 20), (3, 20), (3, 100), (8, 100)])
def test_tile_sharded_symmetric_protocol_emulated_on_one_gpu(engine, world, k):
    """The three-phase multi-GPU protocol (seed -> MAX-reduce -> sweep -> gather -> rescore) with the
    ranks run one after the other on a single GPU; must give exactly the single-GPU table."""
    import torch

    from tvbingefriend_recommendation_service_b200.engine import TopK, stage
    from tvbingefriend_recommendation_service_b200.sharding import row_shard
    from tvbingefriend_recommendation_service_b200.synthetic import make_catalogue

    cat = make_catalogue(6000, 1024, nnz=20, seed=31)
    w, ms = (0.4, 0.5, 0.1), 0.1
    dc = engine.upload(stage(cat.features()), w)
    assert engine.sym_eligible(dc, w, k, ms)
    from tvbingefriend_recommendation_service_b200.sharding import shard_rows

    thetas = [engine.sym_seed(dc, w, k, ms, r, world) for r in range(world)]
    theta = torch.stack(thetas).amax(dim=0)
    parts = [engine.sym_sweep(dc, w, k, ms, r, world, theta.clone()) for r in range(world)]
    cand_all = torch.stack([p[0] for p in parts])
    cnt_all = torch.stack([p[1] for p in parts])
    bound_all = torch.stack([p[2] for p in parts])
    # the packed form (what the NCCL driver exchanges in one all-to-all): candidate entries + {count, bound}
    rows = shard_rows(6000, world)
    packed_all = torch.stack([engine.sym_sweep(dc, w, k, ms, r, world, theta.clone(), packed_rows=world * rows)
                              for r in range(world)])
    L = cand_all.shape[2]
    assert torch.equal(packed_all[:, :6000, L, 0], cnt_all)
    assert torch.equal(packed_all[:, :6000, L, 1].view(torch.float32), bound_all)
    tabs = []
    for r in range(world):
        b, e = row_shard(6000, world, r)
        if r % 3 == 0:   # tables after an all-gather: every rank's lists for all shows
            t = engine.sym_rescore(dc, w, k, ms, cand_all, cnt_all, bound_all, b, e)
        elif r % 3 == 1:  # tables after the all-to-all: every rank's lists for this rank's rows only
            t = engine.sym_rescore(dc, w, k, ms, cand_all[:, b:e].contiguous(), cnt_all[:, b:e].contiguous(),
                                   bound_all[:, b:e].contiguous(), b, e, table_row0=b)
        else:            # packed rows after the all-to-all (padded shard)
            t = engine.sym_rescore(dc, w, k, ms, packed_all[:, r * rows:(r + 1) * rows].contiguous(), None, None,
                                   b, e, table_row0=b)
        tabs.append(engine.to_host(t))
    got = TopK(*(np.concatenate([getattr(t, f) for t in tabs]) for f in
                 ("indices", "counts", "hybrid", "genre", "text", "metadata")))
    ref = engine.to_host(engine.top_k_device(dc, w, k, ms, force_exact=True))
    assert np.array_equal(got.indices, ref.indices) and np.array_equal(got.counts, ref.counts)
    m = ref.indices >= 0
    assert np.array_equal(got.hybrid[m], ref.hybrid[m])
    assert_topk_matches(got, cat.features(), np.arange(0, 6000, 97), w, k, ms)


@pytest.mark.parametrize("mode,norm,w", [("hstack", True, (0.4, 0.5, 0.1)), ("mean3", False, (0.3, 0.6, 0.1))])
def test_streaming_statistics_match_the_full_matrices(engine, mode, norm, w):
    """get_similarity_statistics for all four matrices from one streaming sweep (no N x N) against
    the reference arithmetic on the materialised float64 matrices (oracle), N = 3000."""
    from oracle.reference_paths import ProductionRows, SimilarityComputerOracle
    from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer
    from tvbingefriend_recommendation_service_b200.synthetic import make_catalogue

    cat = make_catalogue(3000, 2048, nnz=30, seed=17)
    got = SimilarityComputer(*w, engine=engine).compute_similarity_statistics(
        cat.features(), metadata_mode=mode, normalize_weights=norm)
    pr = ProductionRows(cat.features(), *w, metadata_mode=mode, normalize_weights=norm)
    mats = {k: np.zeros((3000, 3000)) for k in ("hybrid_similarity", "genre_similarity", "text_similarity",
                                                "metadata_similarity")}
    for i in range(3000):
        h, g, t, m = pr.row(i)
        mats["hybrid_similarity"][i], mats["genre_similarity"][i] = h, g
        mats["text_similarity"][i], mats["metadata_similarity"][i] = t, m
    for name, mat in mats.items():
        ref = SimilarityComputerOracle.get_similarity_statistics(mat)
        g_ = got[name]
        # SURVEY.md section 8f-1 target: mean / std / min / max to 1e-6.  genre / metadata elements
        # are fp32 values summed in float64; every text-dependent sum (text and hybrid) comes from the
        # exact float64 Gram-type moments (tvbf_text_moments), not from the fp16 sweep
        assert g_["exact_moments"] == {"text_mean": True, "text_std": True}
        assert g_["mean"] == pytest.approx(ref["mean"], rel=1e-6, abs=1e-10), name
        assert g_["std"] == pytest.approx(ref["std"], rel=1e-6, abs=1e-9), name
        assert g_["min"] == pytest.approx(ref["min"], abs=1e-6), name
        assert g_["max"] == pytest.approx(ref["max"], rel=1e-6, abs=1e-6), name
        assert abs(g_["median"] - ref["median"]) <= g_["median_resolution"] * 1.01, (name, g_["median"], ref["median"])


def test_streaming_statistics_at_c3_scale(engine):
    """100 k shows: statistics of 5e9 pairs per matrix in one sweep; sanity against a row sample."""
    from oracle.reference_paths import ProductionRows
    from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer
    from tvbingefriend_recommendation_service_b200.synthetic import make_config

    cat = make_config("C3")
    got = SimilarityComputer(engine=engine).compute_similarity_statistics(cat.features())
    pr = ProductionRows(cat.features(), metadata_mode="hstack", normalize_weights=True)
    n = 100_000
    count = n * (n - 1) / 2

    def exact_mean(xn):
        """mean over i<j of <x_i, x_j> = (|sum_i x_i|^2 - sum_i |x_i|^2) / 2 / count, float64."""
        col = np.asarray(xn.sum(axis=0)).ravel()
        sq = float(xn.multiply(xn).sum()) if hasattr(xn, "multiply") else float((xn * xn).sum())
        return (float(col @ col) - sq) / 2 / count

    mg, mt, mm = exact_mean(pr.G), exact_mean(pr.T), exact_mean(pr.M[0])
    assert got["genre_similarity"]["mean"] == pytest.approx(mg, rel=2e-6)
    assert got["metadata_similarity"]["mean"] == pytest.approx(mm, rel=2e-6)
    assert got["text_similarity"]["mean"] == pytest.approx(mt, rel=1e-9)          # exact float64 moments
    assert got["hybrid_similarity"]["mean"] == pytest.approx(pr.gw * mg + pr.tw * mt + pr.mw * mm, rel=1e-6)
    # sum over i<j of t_ij^2 = (||X^T X||_F^2 - sum_i |x_i|^4) / 2: the vocabulary Gram matrix in float64
    gram = (pr.T.T @ pr.T).toarray()
    sq4 = float(np.asarray(pr.T.multiply(pr.T).sum(axis=1)).ravel() @ np.asarray(pr.T.multiply(pr.T).sum(axis=1)).ravel())
    std_t = np.sqrt((float((gram * gram).sum()) - sq4) / 2 / count - mt * mt)
    assert got["text_similarity"]["std"] == pytest.approx(std_t, rel=1e-8)
    rows = np.linspace(0, 99_999, 40).astype(np.int64)
    sample = np.concatenate([np.delete(pr.row(int(i))[0], int(i)) for i in rows])   # hybrid, self removed
    assert got["hybrid_similarity"]["std"] == pytest.approx(sample.std(), rel=0.1)
    assert abs(got["hybrid_similarity"]["median"] - np.median(sample)) < 0.02
    for name in ("genre_similarity", "text_similarity", "metadata_similarity", "hybrid_similarity"):
        s = got[name]
        assert 0.0 <= s["min"] <= s["median"] <= s["max"] <= 1.0 + 2e-7 and s["std"] >= 0   # genre/metadata extrema are fp32
    assert got["text_similarity"]["max"] == pytest.approx(1.0, abs=1e-9)   # planted duplicate shows


@pytest.mark.parametrize("with_text", [False, True], ids=["no_text", "text"])
def test_tie_plateau_wider_than_the_survivor_lists(engine, with_text):
    """17 000 identical shows: every row's floor is reached by all other columns, more than the exact
    kernel lists per row (16 384), so it must fall back to dense keys -- with and without text."""
    import scipy.sparse as sp

    n = 17_000
    text = sp.csr_matrix(np.tile(np.array([[0.6, 0.8, 0.0]]), (n, 1))) if with_text else sp.csr_matrix((n, 3))
    f = {"genre_features": np.ones((n, 3), dtype=np.int64), "text_features": text,
         "platform_features": np.tile(np.array([[1.0, 0.0]]), (n, 1)),
         "type_features": np.tile(np.array([[True, False]]), (n, 1)),
         "language_features": np.tile(np.array([[1.0, 0.0]]), (n, 1))}
    top = engine.compute_top_k(f, (0.4, 0.5, 0.1), 20, 0.1)
    for i in (0, 7, 9000, n - 1):
        assert top.indices[i].tolist() == [j for j in range(n) if j != i][:20]
    assert np.allclose(top.hybrid, 1.0 if with_text else 0.5)
    assert top.flagged_rows == n
