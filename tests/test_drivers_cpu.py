"""Host-side contract of the drop-in drivers that needs no GPU (mirrors
tests/test_scripts/test_compute_similarities.py:46-89, :304-524 and
tests/test_services/test_content_based_service.py:290-303 of the reference): required files and
the FileNotFoundError text, exit codes of the ``main``s, save/skip of the N x N matrices, unknown
show ids, and the loud failure when the job reaches the GPU path without a GPU."""

import numpy as np
import pytest
import scipy.sparse as sp

from tvbingefriend_recommendation_service_b200.scripts import compute_similarities as cs
from tvbingefriend_recommendation_service_b200.synthetic import make_catalogue


def _write_features(d, n=12):
    cat = make_catalogue(n, 40, nnz=5, seed=3)
    cat.save(d)
    return cat


def test_load_features_requires_the_five_reference_files(tmp_path):
    _write_features(tmp_path)
    f = cs.load_features(tmp_path)
    assert set(f) == {"genre_features", "text_features", "platform_features", "type_features",
                      "language_features"}
    assert sp.issparse(f["text_features"]) and f["genre_features"].shape[0] == 12
    (tmp_path / "type_features.npy").unlink()
    with pytest.raises(FileNotFoundError) as e:
        cs.load_features(tmp_path)
    assert "Feature file not found" in str(e.value) and "Run compute_features.py first." in str(e.value)
    assert "type_features.npy" in str(e.value)


def test_save_similarities_is_skipped_unless_asked(tmp_path):
    sims = {"genre_similarity": np.eye(3), "hybrid_similarity": np.full((3, 3), 0.5)}
    cs.save_similarities(sims, tmp_path / "out", save_to_disk=False)
    assert not (tmp_path / "out").exists()
    cs.save_similarities(sims, tmp_path / "out", save_to_disk=True)
    assert np.array_equal(np.load(tmp_path / "out" / "hybrid_similarity.npy"), sims["hybrid_similarity"])


def test_main_exits_1_on_bad_weights_and_on_any_error(tmp_path):
    with pytest.raises(SystemExit) as e:
        cs.main(["--genre-weight", "0", "--text-weight", "0", "--metadata-weight", "0"])
    assert e.value.code == 1
    with pytest.raises(SystemExit) as e:      # missing feature files
        cs.main(["--input-dir", str(tmp_path / "nowhere")])
    assert e.value.code == 1


def test_gpu_path_fails_loudly_without_a_gpu(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    from tvbingefriend_recommendation_service_b200._lib import TvbfError
    from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer
    from tvbingefriend_recommendation_service_b200.scripts import populate_database as pd_

    cat = _write_features(tmp_path)
    with pytest.raises(TvbfError, match="no CPU fallback"):
        SimilarityComputer().compute_top_k(cat.features())
    with pytest.raises(TvbfError, match="no CPU fallback"):
        pd_.compute_and_store_similarities(tmp_path)
    with pytest.raises(SystemExit) as e:      # the driver's main logs and exits 1, like the reference
        pd_.main(["--input-dir", str(tmp_path)])
    assert e.value.code == 1


def test_similarity_computer_keeps_the_reference_constructor():
    from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer

    c = SimilarityComputer()
    assert (c.genre_weight, c.text_weight, c.metadata_weight) == (0.4, 0.5, 0.1)      # reference :15-28
    c = SimilarityComputer(genre_weight=2.0, text_weight=3.0, metadata_weight=1.0)
    assert (c.genre_weight, c.text_weight, c.metadata_weight) == (2.0, 3.0, 1.0)


# ---- storage half of the populate driver (no GPU: the table is made by hand) ---------------------
class _FakeRepository:
    """Shaped like SimilarityRepository (repos/similarity_repository.py:72-124): no ``records``
    attribute, only the two methods the drivers call."""

    def __init__(self):
        self.calls = []
        self.rows = {}

    def bulk_store_all_similarities(self, all_similarities, batch_size=1000, clear_existing=True):
        self.calls.append((len(all_similarities), clear_existing))
        if clear_existing:
            self.rows = {}
        n = 0
        for sid, recs in all_similarities.items():
            self.rows[sid] = recs
            n += len(recs)
        return n

    def get_similarity_stats(self):
        total = sum(len(v) for v in self.rows.values())
        return {"total_records": total, "unique_shows": len(self.rows),
                "avg_similarities_per_show": total / max(len(self.rows), 1), "last_computed": None}


def _hand_table(n=12001, k=3):
    from tvbingefriend_recommendation_service_b200.engine import TopK

    rng = np.random.default_rng(0)
    idx = rng.integers(0, n, size=(n, k)).astype(np.int32)
    cnt = rng.integers(0, k + 1, size=n).astype(np.int32)
    cnt[5] = 0
    idx[np.arange(k)[None, :] >= cnt[:, None]] = -1
    sc = [rng.random((n, k)) for _ in range(4)]
    return TopK(idx, cnt, *sc), np.arange(100, 100 + 3 * n, 3)


def test_store_batches_clears_first_even_for_a_real_repository_shape():
    """reference populate_database.py:156-162 deletes everything up front, then appends per 5000
    shows (:223-234) -- also when the sink is a repository without a ``records`` attribute."""
    from tvbingefriend_recommendation_service_b200.scripts.populate_database import store_similarity_batches

    top, ids = _hand_table()
    repo = _FakeRepository()
    repo.rows = {-1: [{"stale": True}]}
    total = store_similarity_batches(top, ids, repo)
    assert repo.calls[0] == (0, True) and all(c[1] is False for c in repo.calls[1:])
    assert len(repo.calls) == 1 + 3                     # 12001 shows -> batches of 5000, 5000, 2001
    assert -1 not in repo.rows and total == int(top.counts.sum())
    assert ids[5] not in repo.rows                      # shows without a neighbour are omitted (:220-221)
    r = 7 if top.counts[7] else int(np.nonzero(top.counts)[0][0])
    rec = repo.rows[ids[r]][0]
    assert set(rec) == {"similar_show_id", "similarity_score", "genre_score", "text_score", "metadata_score"}
    assert rec["similar_show_id"] == ids[top.indices[r, 0]] and rec["similarity_score"] == top.hybrid[r, 0]


def test_columnar_sink_receives_the_same_record_stream():
    from tvbingefriend_recommendation_service_b200.scripts.populate_database import store_similarity_batches
    from tvbingefriend_recommendation_service_b200.sinks import ColumnarSimilaritySink, InMemorySimilaritySink

    top, ids = _hand_table()
    col, mem = ColumnarSimilaritySink(), InMemorySimilaritySink()
    col.bulk_store_records({c: np.zeros(1, dtype=np.int64) for c in
                            ("show_id", "similar_show_id", "similarity_score", "genre_score", "text_score",
                             "metadata_score")})       # stale content that must be cleared
    assert store_similarity_batches(top, ids, col) == store_similarity_batches(top, ids, mem)
    c = col.columns()
    flat = [(sid, r["similar_show_id"], r["similarity_score"], r["genre_score"], r["text_score"], r["metadata_score"])
            for sid in ids.tolist() for r in mem.records.get(sid, [])]
    assert len(flat) == len(c["show_id"]) == int(top.counts.sum())
    got = list(zip(c["show_id"].tolist(), c["similar_show_id"].tolist(), c["similarity_score"].tolist(),
                   c["genre_score"].tolist(), c["text_score"].tolist(), c["metadata_score"].tolist()))
    assert got == flat
    assert col.get_similarity_stats()["unique_shows"] == mem.get_similarity_stats()["unique_shows"]
    assert set(col.get_similarity_stats()) >= {"total_records", "unique_shows", "avg_similarities_per_show",
                                               "last_computed"}          # repos/similarity_repository.py:239-263


def test_populate_cli_accepts_the_reference_flags(tmp_path):
    """--skip-metadata / --skip-test (reference populate_database.py:318-323) parse; without a GPU the
    similarity step then fails loudly and main exits 1 like the reference (:399-401)."""
    import torch

    from tvbingefriend_recommendation_service_b200.scripts import populate_database as pd_

    _write_features(tmp_path)
    if torch.cuda.is_available():
        stats = pd_.main(["--input-dir", str(tmp_path), "--skip-metadata", "--skip-test", "--top-n", "3"])
        assert stats["top_n_per_show"] == 3
    else:
        with pytest.raises(SystemExit) as e:
            pd_.main(["--input-dir", str(tmp_path), "--skip-metadata", "--skip-test", "--top-n", "3"])
        assert e.value.code == 1
    with pytest.raises(SystemExit) as e:      # unknown flags are still argparse errors (exit 2)
        pd_.main(["--no-such-flag"])
    assert e.value.code == 2
