"""GPU parity tests (``-m gpu``): every call goes through the C ABI of libtvbf.so and is checked
against the CPU oracle (float64 restatement of the reference) or against the committed golden
vectors produced by the unmodified reference.  Tolerances: indices exact (tie-aware comparator,
eps = 1e-9 on float64 scores); scores 1e-5 relative (north_star), in practice ~1e-15."""

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import load_golden
from helpers import assert_topk_matches
from oracle.cosine import normalize_rows
from oracle.reference_paths import SimilarityComputerOracle

pytestmark = pytest.mark.gpu

SYM_OFF, SYM_ON = 1 << 20, 2 << 20   # tvbf_params.tuning bits 20-21


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


@pytest.fixture(scope="module")
def cat2k():
    from tvbingefriend_recommendation_service_b200.synthetic import make_catalogue

    return make_catalogue(2000, 3000, nnz=30, seed=42)


def _cases(z):
    for c in range(int(z["n_cases"])):
        gw, tw, mw, k, ms = z[f"case{c}_params"].tolist()
        yield c, (gw, tw, mw), int(k), ms


# ---- K0 ------------------------------------------------------------------------------------------
def test_prep_operand_and_packing(engine, cat2k):
    from tvbingefriend_recommendation_service_b200.engine import TEXT_SCALE_LOG2, stage

    dc = engine.upload(stage(cat2k.features()))
    indptr, indices, values, operand, col_side, meta_scale = dc.keep[:6]
    tn = normalize_rows(cat2k.text_features)
    assert np.abs(values.cpu().numpy() - tn.data).max() < 1e-15
    dense = (tn.toarray() * 2.0 ** TEXT_SCALE_LOG2).astype(np.float16)
    got = operand.cpu().numpy()
    assert got.shape[0] % 256 == 0 and got.shape[1] % 64 == 0
    assert np.array_equal(got[:2000, :3000], dense)
    assert not got[2000:].any() and not got[:, 3000:].any()
    rec = col_side.cpu().numpy().view(np.uint64)[:2000]
    bits = (cat2k.genre_features.astype(np.uint64) << np.arange(40, dtype=np.uint64)[None, :]).sum(axis=1)
    assert np.array_equal(rec[:, 0], bits)
    low = rec[:, 1]
    rnorm = (low & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.float32)
    pc = cat2k.genre_features.sum(axis=1)
    want = np.where(pc > 0, 1.0 / np.sqrt(np.maximum(pc, 1)), 0.0).astype(np.float32)
    assert np.allclose(rnorm, want, rtol=1e-6)
    mbits = (low >> np.uint64(32)).astype(np.uint32)
    P, T = cat2k.platform_features.shape[1], cat2k.type_features.shape[1]
    want_bits = np.zeros(2000, dtype=np.uint32)
    for off, m in ((0, cat2k.platform_features), (P, cat2k.type_features), (P + T, cat2k.language_features)):
        for c in range(m.shape[1]):
            want_bits |= ((m[:, c] != 0).astype(np.uint32) << np.uint32(off + c))
    assert np.array_equal(mbits, want_bits)
    assert np.allclose(meta_scale.cpu().numpy()[:2000], 1 / np.sqrt(3), rtol=1e-6)


# ---- K1 descriptors / TMEM layout: raw tile against a plain fp32 matmul of the same operand -------
@pytest.mark.parametrize("row0,col0", [(0, 0), (128, 256), (1920, 1792)])
def test_tensor_core_tile_matches_fp32_reference(engine, cat2k, row0, col0):
    from tvbingefriend_recommendation_service_b200.engine import stage

    dc = engine.upload(stage(cat2k.features()))
    operand = dc.keep[3]
    tile = engine.debug_gemm_tile(dc, row0, col0).cpu()
    a = operand[row0:row0 + 128].float().cpu()
    b = operand[col0:col0 + 256].float().cpu()
    ref = a @ b.T
    err = (tile - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item() + 1e-3, err


@pytest.mark.parametrize("row0,col0", [(0, 0), (256, 512), (1792, 1792)])
def test_cta_pair_tile_matches_fp32_reference(engine, cat2k, row0, col0):
    """cta_group::2: M = 256 spans two CTAs' TMEM, each CTA stages half of the B tile."""
    from tvbingefriend_recommendation_service_b200.engine import stage

    dc = engine.upload(stage(cat2k.features()))
    operand = dc.keep[3]
    tile = engine.debug_gemm_tile(dc, row0, col0, pair=True).cpu()
    ref = operand[row0:row0 + 256].float().cpu() @ operand[col0:col0 + 256].float().cpu().T
    err = (tile - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item() + 1e-3, err


def test_fp16_error_bound_holds(engine, cat2k):
    """The certificate relies on |t_fp16 - t_exact| <= rel * t_exact (+ tiny abs)."""
    from tvbingefriend_recommendation_service_b200.engine import TEXT_SCALE_LOG2, stage

    dc = engine.upload(stage(cat2k.features()))
    tn = normalize_rows(cat2k.text_features)
    exact = (tn[:128] @ tn[:256].T).toarray()
    tile = engine.debug_gemm_tile(dc, 0, 0).cpu().numpy().astype(np.float64) * 2.0 ** (-2 * TEXT_SCALE_LOG2)
    rel = (2.0 / 2048 + 2.0 ** -22) * 1.01 + 2.0 ** -18
    assert np.all(np.abs(tile - exact) <= rel * exact + 1e-7)


# ---- K6 exact row kernel ---------------------------------------------------------------------------
@pytest.mark.parametrize("mode,norm", [("mean3", False), ("hstack", True)])
def test_exact_rows_match_oracle(engine, cat2k, mode, norm):
    from tvbingefriend_recommendation_service_b200.engine import stage

    w = (0.4, 0.5, 0.1)
    dc = engine.upload(stage(cat2k.features(), mode), w)
    rows = np.arange(0, 2000, 37)
    top = engine.exact_rows(dc, rows, w, k=20, min_similarity=0.1)
    top.row_begin = 0
    # exact_rows returns one output row per listed row
    from tvbingefriend_recommendation_service_b200.engine import TopK
    full = TopK(np.full((2000, 20), -1, np.int32), np.zeros(2000, np.int32), *(np.full((2000, 20), np.nan) for _ in range(4)))
    for name in ("indices", "counts", "hybrid", "genre", "text", "metadata"):
        getattr(full, name)[rows] = getattr(top, name)
    assert_topk_matches(full, cat2k.features(), rows, w, 20, 0.1, mode, False)


# ---- the full path: K1 -> K5 -> K6 -----------------------------------------------------------------
def test_topk_matches_oracle_2k(engine, cat2k):
    top = engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), 20, 0.1)
    rep = assert_topk_matches(top, cat2k.features())
    assert rep.rows == 2000


def test_forced_exact_equals_certified_path(engine, cat2k):
    a = engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), 20, 0.1)
    b = engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), 20, 0.1, force_exact=True)
    assert np.array_equal(a.indices, b.indices) and np.array_equal(a.counts, b.counts)
    m = a.indices >= 0
    assert np.array_equal(a.hybrid[m], b.hybrid[m])


@pytest.mark.parametrize("tuning", [SYM_OFF, SYM_ON])
@pytest.mark.parametrize("splits", [1, 2, 3])
def test_column_splits_do_not_change_the_table(engine, cat2k, splits, tuning):
    a = engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), 20, 0.1, splits=splits, tuning=tuning)
    b = engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), 20, 0.1, force_exact=True)
    assert np.array_equal(a.indices, b.indices)





@pytest.mark.parametrize("tuning", [0x1, 0x2 | SYM_OFF, 0x1 | (255 << 4), 0x2 | (255 << 4) | SYM_OFF,
                                    0x2 | (1 << 4) | (1 << 12) | SYM_OFF, 0x2 | SYM_ON,
                                    0x2 | SYM_ON | (255 << 4), 0x2 | SYM_ON | (2 << 4) | (1 << 12), 0x2 | SYM_ON | (63 << 22), 0x2 | SYM_ON | (2 << 22)])
def test_kernel_variants_give_the_same_table(engine, cat2k, tuning):
    """cta_group 1 / 2, producer pacing on / off / tight, one-sided / symmetric sweep: same
    certified result."""
    a = engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), 20, 0.1, tuning=tuning, splits=3)
    b = engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), 20, 0.1, force_exact=True)
    assert np.array_equal(a.indices, b.indices) and np.array_equal(a.counts, b.counts)


@pytest.mark.parametrize("name", ["populate_n300", "populate_random_float_n48"])
def test_golden_production_loop(engine, name):
    """Outputs of the UNMODIFIED reference hot loop (scripts/populate_database.py:85-259)."""
    from oracle.compare import compare_topk
    from oracle.reference_paths import ProductionRows

    z, cat = load_golden(name)
    for c, w, k, ms in _cases(z):
        top = engine.compute_top_k(cat.features(), w, k, ms, metadata_mode="mean3")
        pr = ProductionRows(cat.features(), *w)
        rep = compare_topk(z[f"case{c}_idx"].astype(np.int64), z[f"case{c}_cnt"], z[f"case{c}_scores"][0],
                           top.indices, top.counts, top.hybrid, lambda r, js: pr.pair_scores(r, js), k, ms)
        assert rep.ok, rep.summary() + "\n" + "\n".join(rep.failures)
        d = top.to_dict(cat.show_ids.tolist())
        assert sum(len(v) for v in d.values()) == int(z[f"case{c}_total_records"])


@pytest.mark.parametrize("tuning", [SYM_OFF, SYM_ON], ids=["one_sided", "symmetric"])
@pytest.mark.parametrize("mode,norm,w", [("hstack", True, (2.0, 3.0, 1.0)), ("mean3", False, (0.5, 0.5, 0.0)),
                                         ("hstack", False, (0.3, 0.6, 0.1))])
def test_variants_and_weights(engine, cat2k, mode, norm, w, tuning):
    from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer

    top = SimilarityComputer(*w, engine=engine).compute_top_k(cat2k.features(), k=10, min_similarity=0.05,
                                                               metadata_mode=mode, normalize_weights=norm,
                                                               tuning=tuning)
    assert_topk_matches(top, cat2k.features(), None, w, 10, 0.05, mode, norm)


@pytest.mark.parametrize("k", [1, 50, 100])
def test_other_k(engine, cat2k, k):
    top = engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), k, 0.0)
    assert_topk_matches(top, cat2k.features(), np.arange(0, 2000, 7), (0.4, 0.5, 0.1), k, 0.0)


@pytest.mark.parametrize("k", [1, 20, 48, 64, 100])
def test_other_k_symmetric(engine, cat2k, k):
    top = engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), k, 0.02, tuning=SYM_ON)
    assert_topk_matches(top, cat2k.features(), np.arange(0, 2000, 7), (0.4, 0.5, 0.1), k, 0.02)


def test_symmetric_request_on_ineligible_job_is_refused(engine, cat2k):
    from tvbingefriend_recommendation_service_b200._lib import TvbfError

    with pytest.raises(TvbfError):   # k = 150 needs more candidates per show than the shared lists keep
        engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), 150, 0.1, tuning=SYM_ON)
    # a non-positive threshold IS eligible when every score is >= 0 (round 2: thresholds start at +0) ...
    top = engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), 20, 0.0, tuning=SYM_ON)
    assert_topk_matches(top, cat2k.features(), np.arange(0, 2000, 9), (0.4, 0.5, 0.1), 20, 0.0)
    top = engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), 20, -0.25, tuning=SYM_ON)
    assert_topk_matches(top, cat2k.features(), np.arange(0, 2000, 9), (0.4, 0.5, 0.1), 20, -0.25)
    # ... but not with signed text (scores, and so thresholds, may be negative)
    f = dict(cat2k.features())
    f["text_features"] = sp.csr_matrix(np.random.default_rng(1).standard_normal((2000, 32)))
    with pytest.raises(TvbfError):
        engine.compute_top_k(f, (0.4, 0.5, 0.1), 20, 0.0, tuning=SYM_ON)
    assert_topk_matches(engine.compute_top_k(f, (0.4, 0.5, 0.1), 20, 0.0), f, np.arange(0, 2000, 9),
                        (0.4, 0.5, 0.1), 20, 0.0)        # auto: one-sided sweep


@pytest.mark.parametrize("tuning", [SYM_OFF, SYM_ON], ids=["one_sided", "symmetric"])
@pytest.mark.parametrize("n", [1, 2, 127, 129, 300, 513, 1100])
def test_ragged_sizes(engine, n, tuning):
    from tvbingefriend_recommendation_service_b200.synthetic import make_catalogue

    cat = make_catalogue(n, 100, nnz=8, n_genres=12, seed=n)
    top = engine.compute_top_k(cat.features(), (0.4, 0.5, 0.1), 20, 0.1, tuning=tuning)
    assert top.indices.shape == (n, 20)
    assert_topk_matches(top, cat.features())


@pytest.mark.parametrize("tuning", [SYM_OFF, SYM_ON], ids=["one_sided", "symmetric"])
def test_degenerate_rows(engine, tuning):
    """Empty text, all-zero genres, missing type, exact duplicates and a high threshold."""
    from tvbingefriend_recommendation_service_b200.synthetic import make_catalogue

    cat = make_catalogue(600, 200, nnz=6, n_genres=8, seed=5)
    f = cat.features()
    f["genre_features"] = f["genre_features"].copy()
    f["genre_features"][:50] = 0
    t = f["text_features"].tolil()
    for r in range(40, 90):
        t.rows[r], t.data[r] = [], []
    f["text_features"] = t.tocsr()
    for ms in (0.1, 0.75, 0.0):
        top = engine.compute_top_k(f, (0.4, 0.5, 0.1), 20, ms, tuning=tuning)
        assert_topk_matches(top, f, None, (0.4, 0.5, 0.1), 20, ms)
    assert (engine.compute_top_k(f, (0.4, 0.5, 0.1), 20, 5.0, tuning=tuning).counts == 0).all()


@pytest.mark.parametrize("tuning,n", [(SYM_OFF, 300), (SYM_ON, 300), (SYM_ON, 3000)],
                         ids=["one_sided", "symmetric", "symmetric_list_overflow"])
def test_tie_break_is_by_index(engine, tuning, n):
    """Identical shows: all scores tie exactly; the stated order is ascending column index.
    (3000 identical shows overflow the shared candidate lists of the symmetric sweep before a
    threshold refresh can stop the flood: those rows must be repaired by the exact kernel.)"""
    f = {"genre_features": np.ones((n, 3), dtype=np.int64),
         "text_features": sp.csr_matrix(np.tile(np.array([[0.6, 0.8, 0.0]]), (n, 1))),
         "platform_features": np.tile(np.array([[1.0, 0.0]]), (n, 1)),
         "type_features": np.tile(np.array([[True, False]]), (n, 1)),
         "language_features": np.tile(np.array([[1.0, 0.0]]), (n, 1))}
    top = engine.compute_top_k(f, (0.4, 0.5, 0.1), 20, 0.1, tuning=tuning)
    for i in (0, 5, 150, n - 1):
        want = [j for j in range(n) if j != i][:20]
        assert top.indices[i].tolist() == want
    assert np.allclose(top.hybrid, 1.0)
    assert top.flagged_rows == n          # every row is a tie plateau -> repaired by the exact kernel


def test_row_range_shard(engine, cat2k):
    from tvbingefriend_recommendation_service_b200.engine import stage

    dc = engine.upload(stage(cat2k.features()))
    part = engine.to_host(engine.top_k_device(dc, (0.4, 0.5, 0.1), 20, 0.1, row_begin=640, row_end=1500))
    full = engine.compute_top_k(cat2k.features(), (0.4, 0.5, 0.1), 20, 0.1)
    assert part.row_begin == 640 and part.indices.shape[0] == 860
    assert np.array_equal(part.indices, full.indices[640:1500])
    assert np.array_equal(part.counts, full.counts[640:1500])


# ---- N x N variant (SimilarityComputer) against golden output of the reference class ---------------
def test_similarity_computer_matches_reference_golden(engine):
    from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer

    z, cat = load_golden("similarity_computer_n64")
    for w in range(2):
        comp = SimilarityComputer(*z[f"w{w}_weights"].tolist(), engine=engine)
        sims = comp.compute_all_similarities(cat.features())
        for key, mat in sims.items():
            assert mat.dtype == np.float64 and np.abs(mat - z[f"w{w}_{key}"]).max() < 1e-12, key
            st = comp.get_similarity_statistics(z[f"w{w}_{key}"])
            got = [st["mean"], st["std"], st["min"], st["max"], st["median"]]
            assert np.allclose(got, z[f"w{w}_{key}_stats"], rtol=1e-12, atol=1e-14), key


def test_service_matrix_and_feature_modes_match_reference_golden(engine, tmp_path):
    from tvbingefriend_recommendation_service_b200.services.content_based_service import (
        ContentBasedRecommendationService)

    z, cat = load_golden("similarity_computer_n64")
    ids = cat.show_ids.tolist()
    for mode in ("matrix", "features"):
        d = tmp_path / mode
        cat.save(d)
        if mode == "matrix":
            for name in ("genre_similarity", "text_similarity", "metadata_similarity"):
                np.save(d / f"{name}.npy", z[f"w0_{name}"])
        for w in range(2):
            gw, tw, mw = z[f"w{w}_weights"].tolist()
            if mode == "matrix" and w == 1:
                for name in ("genre_similarity", "text_similarity", "metadata_similarity"):
                    np.save(d / f"{name}.npy", z[f"w1_{name}"])
            svc = ContentBasedRecommendationService(d, gw, tw, mw, use_blob=False, engine=engine)
            for tag in "ab":
                n, ms = z[f"w{w}_svc{tag}_params"].tolist()
                for qi, q in enumerate(z["svc_queries"].tolist()):
                    recs = svc.get_recommendations_from_matrix(q, n=int(n), min_similarity=ms)
                    want_n = int(z[f"w{w}_svc{tag}_cnt"][qi])
                    assert len(recs) == want_n
                    if want_n:
                        want_scores = z[f"w{w}_svc{tag}_scores"][0, qi, :want_n]
                        assert np.allclose([r["similarity_score"] for r in recs], want_scores, rtol=1e-9)
                        if np.all(np.diff(want_scores) < -1e-9):
                            assert [ids.index(r["show_id"]) for r in recs] == z[f"w{w}_svc{tag}_idx"][qi, :want_n].tolist()
            stats = svc.compute_and_store_all_similarities(top_n_per_show=5, min_similarity=0.1)
            assert stats["computed_shows"] == stats["unique_shows"] <= 64
            assert stats["top_n_per_show"] == 5 and stats["min_similarity"] == 0.1


def test_populate_driver_end_to_end(engine, tmp_path):
    from tvbingefriend_recommendation_service_b200.scripts.populate_database import compute_and_store_similarities
    from tvbingefriend_recommendation_service_b200.sinks import InMemorySimilaritySink

    z, cat = load_golden("populate_n300")
    cat.save(tmp_path)
    sink = InMemorySimilaritySink()
    stats = compute_and_store_similarities(tmp_path, sink=sink)
    assert stats["total_records"] == int(z["case0_total_records"])
    assert {"total_records", "unique_shows", "avg_similarities_per_show"} <= set(stats)
    ids = cat.show_ids.tolist()
    ref_idx = z["case0_idx"]
    rec = sink.records[ids[0]]
    assert set(rec[0]) == {"similar_show_id", "similarity_score", "genre_score", "text_score", "metadata_score"}
    assert abs(rec[0]["similarity_score"] - z["case0_scores"][0, 0, 0]) < 1e-12
    assert rec[0]["similar_show_id"] == ids[int(ref_idx[0, 0])] or \
        abs(z["case0_scores"][0, 0, 0] - z["case0_scores"][0, 0, 1]) < 1e-9


# ---- the reference's own unit tests for SimilarityComputer, run against the drop-in ---------------
class TestReferenceUnitTests:
    def test_genre(self, engine, sample_genre_features):
        from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer

        s = SimilarityComputer(engine=engine).compute_genre_similarity(sample_genre_features)
        assert s.shape == (3, 3) and np.allclose(np.diag(s), 1.0) and np.allclose(s, s.T)
        assert np.all(s >= -1e-10) and np.all(s <= 1 + 1e-10)

    def test_text_and_metadata(self, engine, sample_text_features, sample_platform_features,
                               sample_type_features, sample_language_features):
        from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer

        c = SimilarityComputer(engine=engine)
        s = c.compute_text_similarity(sample_text_features)
        assert np.allclose(np.diag(s), 1.0) and np.allclose(s, c.compute_text_similarity(sample_text_features.toarray()))
        m = c.compute_metadata_similarity(sample_platform_features, sample_type_features, sample_language_features)
        o = SimilarityComputerOracle().compute_metadata_similarity(sample_platform_features, sample_type_features,
                                                                   sample_language_features)
        assert np.allclose(m, o, atol=1e-14)

    def test_hybrid_known_values(self, engine, sample_similarity_matrix):
        from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer

        h = SimilarityComputer(0.4, 0.5, 0.1, engine=engine).compute_hybrid_similarity(
            sample_similarity_matrix, sample_similarity_matrix * 0.9, sample_similarity_matrix * 0.8)
        assert np.allclose(np.diag(h), 0.93, atol=1e-6)
        h = SimilarityComputer(2.0, 3.0, 1.0, engine=engine).compute_hybrid_similarity(
            np.array([[1, .6], [.6, 1]]), np.array([[1, .8], [.8, 1]]), np.array([[1, .4], [.4, 1]]))
        assert h[0, 1] == pytest.approx((2 / 6) * 0.6 + (3 / 6) * 0.8 + (1 / 6) * 0.4, abs=1e-6)

    def test_statistics(self, engine):
        from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer

        c = SimilarityComputer(engine=engine)
        m = np.array([[1.0, 0.5, 0.5], [0.5, 1.0, 0.5], [0.5, 0.5, 1.0]])
        st = c.get_similarity_statistics(m)
        assert st["mean"] == pytest.approx(0.5) and st["max"] == pytest.approx(0.5)
        m = np.full((4, 4), 0.7)
        np.fill_diagonal(m, 1.0)
        st = c.get_similarity_statistics(m)
        assert st["std"] == pytest.approx(0.0, abs=1e-6) and st["median"] == pytest.approx(0.7)
        assert all(isinstance(v, float) for v in st.values())


# ---- weight sweep over one device-resident catalogue ------------------------------------------------
def test_weight_sweep_equals_separate_runs(engine, cat2k):
    from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer
    from tvbingefriend_recommendation_service_b200.synthetic import WEIGHT_SWEEP

    comp = SimilarityComputer(engine=engine)
    tabs = comp.compute_top_k_sweep(cat2k.features(), WEIGHT_SWEEP, k=10)
    assert len(tabs) == len(WEIGHT_SWEEP)
    for w, got in zip(WEIGHT_SWEEP, tabs):
        ref = SimilarityComputer(*w, engine=engine).compute_top_k(cat2k.features(), k=10)
        assert np.array_equal(got.indices, ref.indices) and np.array_equal(got.hybrid, ref.hybrid, equal_nan=True)
    assert_topk_matches(tabs[3], cat2k.features(), np.arange(0, 2000, 41), WEIGHT_SWEEP[3], 10, 0.1)


def test_weight_sweep_with_folded_features_uploads_per_triple(engine):
    z, cat = load_golden("populate_random_float_n48")
    triples = [(2.0, 3.0, 1.0), (0.4, 0.5, 0.1)]
    tabs = engine.compute_top_k_sweep(cat.features(), triples, 7, 0.5)
    for w, got in zip(triples, tabs):
        ref = engine.compute_top_k(cat.features(), w, 7, 0.5)
        assert np.array_equal(got.indices, ref.indices) and np.array_equal(got.hybrid, ref.hybrid, equal_nan=True)


@pytest.mark.parametrize("k", [20, 100])
def test_shared_tensor_core_sweep_equals_separate_runs(engine, cat2k, k):
    """One symmetric sweep for five weight triples (one candidate list per triple and show)."""
    from tvbingefriend_recommendation_service_b200.engine import stage
    from tvbingefriend_recommendation_service_b200.synthetic import WEIGHT_SWEEP

    dc = engine.upload(stage(cat2k.features()))
    triples = list(WEIGHT_SWEEP) + [(0.0, 1.0, 0.0)]       # 6 triples: two launches (5 + 1)
    tabs = [engine.to_host(t) for t in engine.top_k_sweep_device(dc, triples, k, 0.1, shared=True)]
    assert len(tabs) == len(triples)
    for w, got in zip(triples, tabs):
        ref = engine.to_host(engine.top_k_device(dc, w, k, 0.1))
        assert np.array_equal(got.indices, ref.indices), w
        assert np.array_equal(got.counts, ref.counts), w
        m = ref.indices >= 0
        for name in ("hybrid", "genre", "text", "metadata"):
            assert np.array_equal(getattr(got, name)[m], getattr(ref, name)[m]), (w, name)
    assert_topk_matches(tabs[3], cat2k.features(), np.arange(0, 2000, 41), triples[3], k, 0.1)


def test_shared_sweep_refuses_ineligible_triples(engine, cat2k):
    from tvbingefriend_recommendation_service_b200._lib import TvbfError
    from tvbingefriend_recommendation_service_b200.engine import stage

    dc = engine.upload(stage(cat2k.features()))
    with pytest.raises(TvbfError):   # negative weight: not eligible for the symmetric sweep
        engine.top_k_sweep_device(dc, [(0.4, 0.5, 0.1), (-0.1, 0.5, 0.1)], 20, 0.1, shared=True)
