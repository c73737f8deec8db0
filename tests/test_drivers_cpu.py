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
