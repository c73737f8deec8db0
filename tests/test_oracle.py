"""Pin the CPU oracle: against the reference's own known answers and the committed golden
vectors produced by the unmodified reference (oracle/make_golden.py).  CPU only."""

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from conftest import load_golden
from oracle.compare import compare_topk
from oracle.cosine import cosine_similarity
from oracle.reference_paths import (ProductionRows, SimilarityComputerOracle, dict_to_arrays,
                                    production_loop, service_hybrid, service_recommendations)


# ---- known answers of the reference's tests/test_ml/test_similarity_computer.py ----------------
class TestReferenceKnownAnswers:
    def test_defaults(self):  # :12-30
        c = SimilarityComputerOracle()
        assert (c.genre_weight, c.text_weight, c.metadata_weight) == (0.4, 0.5, 0.1)

    def test_genre_basic(self, sample_genre_features):  # :36-52
        s = SimilarityComputerOracle().compute_genre_similarity(sample_genre_features)
        assert s.shape == (3, 3)
        assert np.allclose(np.diag(s), 1.0)
        assert np.allclose(s, s.T)
        assert np.all(s >= -1e-10) and np.all(s <= 1 + 1e-10)

    def test_genre_identical_and_disjoint(self):  # :54-79
        c = SimilarityComputerOracle()
        s = c.compute_genre_similarity(np.array([[1, 1, 0], [1, 1, 0]], dtype=float))
        assert s[0, 1] == pytest.approx(1.0, abs=1e-6)
        s = c.compute_genre_similarity(np.array([[1, 0, 0], [0, 1, 0]], dtype=float))
        assert s[0, 1] == pytest.approx(0.0, abs=1e-6)

    def test_text_dense_and_sparse(self, sample_text_features):  # :85-111
        c = SimilarityComputerOracle()
        s = c.compute_text_similarity(sample_text_features)
        d = c.compute_text_similarity(sample_text_features.toarray())
        assert s.shape == (3, 3) and np.allclose(np.diag(s), 1.0) and np.allclose(s, d)

    def test_hybrid_known_values(self, sample_similarity_matrix):  # :158-211
        c = SimilarityComputerOracle(0.4, 0.5, 0.1)
        h = c.compute_hybrid_similarity(sample_similarity_matrix, sample_similarity_matrix * 0.9,
                                        sample_similarity_matrix * 0.8)
        assert np.allclose(np.diag(h), 0.93, atol=1e-6)
        c = SimilarityComputerOracle(2.0, 3.0, 1.0)
        h = c.compute_hybrid_similarity(np.array([[1, .6], [.6, 1]]), np.array([[1, .8], [.8, 1]]),
                                        np.array([[1, .4], [.4, 1]]))
        assert h[0, 1] == pytest.approx((2 / 6) * 0.6 + (3 / 6) * 0.8 + (1 / 6) * 0.4, abs=1e-6)
        c = SimilarityComputerOracle(1.0, 1.0, 1.0)
        h = c.compute_hybrid_similarity(np.array([[1, .3], [.3, 1]]), np.array([[1, .6], [.6, 1]]),
                                        np.array([[1, .9], [.9, 1]]))
        assert h[0, 1] == pytest.approx((0.3 + 0.6 + 0.9) / 3, abs=1e-6)

    def test_statistics(self):  # :280-312
        c = SimilarityComputerOracle()
        m = np.array([[1.0, 0.5, 0.5], [0.5, 1.0, 0.5], [0.5, 0.5, 1.0]])
        st = c.get_similarity_statistics(m)
        assert st["mean"] == pytest.approx(0.5) and st["max"] == pytest.approx(0.5)
        m = np.full((4, 4), 0.7)
        np.fill_diagonal(m, 1.0)
        st = c.get_similarity_statistics(m)
        assert st["std"] == pytest.approx(0.0, abs=1e-6) and st["median"] == pytest.approx(0.7)

    def test_zero_rows_give_zero_not_nan(self):
        x = np.array([[0, 0, 0], [1, 0, 1]], dtype=np.int64)
        s = cosine_similarity(x)
        assert s[0, 0] == 0.0 and s[0, 1] == 0.0 and not np.isnan(s).any()
        s = cosine_similarity(csr_matrix(x.astype(float)))
        assert s[0, 0] == 0.0 and not np.isnan(s).any()

    def test_matches_sklearn_when_installed(self):
        sk = pytest.importorskip("sklearn.metrics.pairwise")
        rng = np.random.default_rng(0)
        x = rng.random((17, 9))
        assert np.abs(sk.cosine_similarity(x) - cosine_similarity(x)).max() < 1e-15
        xs = csr_matrix(np.where(rng.random((17, 40)) < 0.2, rng.random((17, 40)), 0.0))
        assert np.abs(sk.cosine_similarity(xs[3:4], xs) - cosine_similarity(xs[3:4], xs)).max() < 1e-15
        f32 = sk.cosine_similarity(x.astype(np.float32))
        assert f32.dtype == np.float32 == cosine_similarity(x.astype(np.float32)).dtype


# ---- golden vectors from the unmodified reference ---------------------------------------------
def _cases(z):
    for c in range(int(z["n_cases"])):
        gw, tw, mw, k, ms = z[f"case{c}_params"].tolist()
        yield c, dict(genre_weight=gw, text_weight=tw, metadata_weight=mw,
                      top_n_per_show=int(k), min_similarity=ms)


@pytest.mark.parametrize("name", ["populate_n300", "populate_random_float_n48", "populate_v500_n1500"])
def test_production_loop_matches_real_reference(name):
    z, cat = load_golden(name)
    for c, kw in _cases(z):
        got = production_loop(cat.features(), cat.show_ids.tolist(), **kw)
        idx, cnt, sc = dict_to_arrays(got, cat.show_ids.tolist(), kw["top_n_per_show"])
        assert np.array_equal(cnt, z[f"case{c}_cnt"])
        assert np.array_equal(idx, z[f"case{c}_idx"])
        assert np.nanmax(np.abs(sc - z[f"case{c}_scores"])) < 1e-14
        assert sum(len(v) for v in got.values()) == int(z[f"case{c}_total_records"])


def test_hoisted_rows_identical_to_verbatim_loop():
    z, cat = load_golden("populate_n300")
    for c, kw in _cases(z):
        pr = ProductionRows(cat.features(), kw["genre_weight"], kw["text_weight"], kw["metadata_weight"])
        idx, cnt, sc = pr.topk_arrays(range(cat.n_shows), kw["top_n_per_show"], kw["min_similarity"])
        assert np.array_equal(cnt, z[f"case{c}_cnt"]) and np.array_equal(idx, z[f"case{c}_idx"])
        assert np.nanmax(np.abs(sc - z[f"case{c}_scores"])) < 1e-14


def test_similarity_computer_and_service_match_real_reference():
    z, cat = load_golden("similarity_computer_n64")
    for w in range(2):
        comp = SimilarityComputerOracle(*z[f"w{w}_weights"].tolist())
        sims = comp.compute_all_similarities(cat.features())
        for key, mat in sims.items():
            assert np.abs(mat - z[f"w{w}_{key}"]).max() < 1e-14
            st = comp.get_similarity_statistics(mat)
            assert np.allclose([st["mean"], st["std"], st["min"], st["max"], st["median"]],
                               z[f"w{w}_{key}_stats"], rtol=0, atol=1e-14)
        gw, tw, mw = z[f"w{w}_weights"].tolist()
        hyb = service_hybrid(sims["genre_similarity"], sims["text_similarity"],
                             sims["metadata_similarity"], gw, tw, mw)
        ids = cat.show_ids.tolist()
        for tag in "ab":
            n, ms = z[f"w{w}_svc{tag}_params"].tolist()
            for qi, q in enumerate(z["svc_queries"].tolist()):
                recs = service_recommendations(hyb, sims["genre_similarity"], sims["text_similarity"],
                                               sims["metadata_similarity"], ids, q, n=int(n),
                                               min_similarity=ms)
                assert len(recs) == int(z[f"w{w}_svc{tag}_cnt"][qi])
                got = [ids.index(r["show_id"]) for r in recs]
                assert got == z[f"w{w}_svc{tag}_idx"][qi, :len(recs)].tolist()
                if recs:
                    assert np.allclose([r["similarity_score"] for r in recs],
                                       z[f"w{w}_svc{tag}_scores"][0, qi, :len(recs)], atol=1e-14)
    assert service_recommendations(hyb, hyb, hyb, hyb, ids, 10 ** 9) == []


# ---- the comparator itself --------------------------------------------------------------------
def _mk(idx, sc, k):
    i = np.full((1, k), -1, dtype=np.int64)
    s = np.full((1, k), np.nan)
    i[0, :len(idx)] = idx
    s[0, :len(sc)] = sc
    return i, np.array([len(idx)]), s


class TestComparator:
    full = np.array([0.0, 0.9, 0.8, 0.8, 0.8, 0.3, 0.05])  # reference row of source 0

    def ps(self, r, js):
        return self.full[js]

    def test_identical(self):
        a = _mk([1, 2, 3], [0.9, 0.8, 0.8], 3)
        rep = compare_topk(*a, *a, self.ps, 3, 0.1)
        assert rep.ok and rep.rows_identical_ordered == 1

    def test_tie_at_cut_may_swap(self):
        ref = _mk([1, 2, 3], [0.9, 0.8, 0.8], 3)
        got = _mk([1, 2, 4], [0.9, 0.8, 0.8], 3)  # 4 ties with the cut value 0.8
        rep = compare_topk(*ref, *got, self.ps, 3, 0.1)
        assert rep.ok and rep.rows_tie_permuted == 1

    def test_wrong_member_fails(self):
        ref = _mk([1, 2, 3], [0.9, 0.8, 0.8], 3)
        got = _mk([1, 2, 5], [0.9, 0.8, 0.3], 3)
        assert not compare_topk(*ref, *got, self.ps, 3, 0.1).ok

    def test_order_defined_must_match(self):
        ref = _mk([1, 2], [0.9, 0.8], 2)
        got = _mk([2, 1], [0.9, 0.8], 2)
        assert not compare_topk(*ref, *got, self.ps, 2, 0.1).ok

    def test_score_tolerance(self):
        ref = _mk([1, 2, 3], [0.9, 0.8, 0.8], 3)
        got = _mk([1, 2, 3], [0.9 * (1 + 5e-5), 0.8, 0.8], 3)
        assert not compare_topk(*ref, *got, self.ps, 3, 0.1).ok

    def test_short_list_and_threshold(self):
        ref = _mk([1, 2, 3, 4, 5], [0.9, 0.8, 0.8, 0.8, 0.3], 6)
        got = _mk([1, 2, 3, 4], [0.9, 0.8, 0.8, 0.8], 6)
        assert not compare_topk(*ref, *got, self.ps, 6, 0.1).ok  # dropped a clear member
        got = _mk([1, 2, 3, 4, 5, 6], [0.9, 0.8, 0.8, 0.8, 0.3, 0.05], 6)
        assert not compare_topk(*ref, *got, self.ps, 6, 0.1).ok  # below threshold

    def test_unsorted_own_list_fails(self):
        ref = _mk([1, 2, 3], [0.9, 0.8, 0.8], 3)
        got = _mk([1, 3, 2], [0.9, 0.8, 0.8], 3)  # equal scores must be index-ascending
        assert not compare_topk(*ref, *got, self.ps, 3, 0.1).ok


# ---- property test: hoisted oracle == verbatim loop on random (tie-heavy) inputs --------------------
def test_hoisted_rows_equal_the_verbatim_loop_on_random_inputs():
    """ProductionRows (normalisations hoisted; what the GPU parity tests use at scale) against the
    verbatim loop on small random catalogues with many exact ties, empty rows and float features.
    Both sort with the same unstable argsort, so only scores and tie-insensitive facts are compared
    bit-for-bit: counts, score vectors, and the index sets above the cut."""
    import scipy.sparse as sp

    rng = np.random.default_rng(5)
    for trial in range(30):
        n = int(rng.integers(2, 40))
        g_dim, v = int(rng.integers(1, 6)), int(rng.integers(1, 12))
        genre = (rng.random((n, g_dim)) < 0.5).astype(np.int64)
        text = sp.csr_matrix(np.where(rng.random((n, v)) < 0.3, rng.integers(1, 3, (n, v)), 0).astype(np.float64))
        onehot = lambda c: np.eye(c)[rng.integers(0, c, n)]   # noqa: E731
        f = {"genre_features": genre if trial % 3 else rng.random((n, g_dim)),
             "text_features": text, "platform_features": onehot(3),
             "type_features": onehot(2).astype(bool), "language_features": onehot(2)}
        if trial % 4 == 0:
            f["genre_features"][0] = 0
        w = tuple(float(x) for x in rng.choice([0.0, 0.1, 0.4, 0.5, 1.0, 2.0], 3))
        if sum(w) == 0:
            w = (0.4, 0.5, 0.1)
        k, ms = int(rng.integers(1, 8)), float(rng.choice([0.0, 0.1, 0.5]))
        ids = list(range(100, 100 + n))
        ref = production_loop(f, ids, *w, top_n_per_show=k, min_similarity=ms)
        ridx, rcnt, rsc = dict_to_arrays(ref, ids, k)
        pr = ProductionRows(f, *w)
        idx, cnt, sc = pr.topk_arrays(range(n), k, ms)
        assert np.array_equal(cnt, rcnt), trial
        assert np.allclose(np.nan_to_num(sc), np.nan_to_num(rsc), rtol=0, atol=1e-15), trial
        for i in range(n):   # members strictly above the row's last kept score are the same set
            c = int(cnt[i])
            if c:
                cut = rsc[0][i, c - 1]
                assert set(idx[i, :c][sc[0][i, :c] > cut]) == set(ridx[i, :c][rsc[0][i, :c] > cut]), (trial, i)
