"""Run the UNMODIFIED reference functions from /root/reference.  Test infrastructure only.

Only usable in the build container (``/root/reference`` does not exist on the GPU box); the GPU
tests read the fixtures this produced (``tests/golden``) instead.  The reference imports
``sqlalchemy`` and the private ``tvbingefriend_azure_storage_service`` package at module import;
neither is installed, and neither touches the arithmetic, so both are replaced by ``MagicMock``
modules and the repository / session are replaced by capturing fakes -- the same patch points the
reference's own test uses (tests/test_scripts/test_populate_database.py:148-154).
"""

from __future__ import annotations

import importlib
import os
import sys
import tempfile
from pathlib import Path
from unittest.mock import MagicMock, patch

REFERENCE_ROOT = Path(os.environ.get("TVBF_REFERENCE_ROOT", "/root/reference"))

_STUBBED = ["sqlalchemy", "sqlalchemy.orm", "sqlalchemy.ext", "sqlalchemy.ext.declarative",
            "sqlalchemy.sql", "sqlalchemy.dialects", "sqlalchemy.dialects.mysql",
            "tvbingefriend_azure_storage_service"]


def available() -> bool:
    return (REFERENCE_ROOT / "scripts" / "populate_database.py").exists()


def _prepare():
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in _STUBBED:
        sys.modules.setdefault(name, MagicMock())
    os.environ.setdefault("DATABASE_URL", "sqlite://")
    root = str(REFERENCE_ROOT)
    if root not in sys.path:
        sys.path.insert(0, root)


def similarity_computer_class():
    """The reference's own ``SimilarityComputer`` (imports cleanly: numpy + sklearn only)."""
    _prepare()
    mod = importlib.import_module("tvbingefriend_recommendation_service.ml.similarity_computer")
    return mod.SimilarityComputer


class _CapturingRepo:
    """Stands in for SimilarityRepository: records what the hot loop hands to the sink
    (repos/similarity_repository.py:72-124 contract: dict[show_id] -> list[dict])."""

    captured: dict = {}

    def __init__(self, _db):
        pass

    def bulk_store_all_similarities(self, all_similarities, batch_size=1000, clear_existing=True):
        n = 0
        for sid, recs in all_similarities.items():
            type(self).captured[sid] = [dict(r) for r in recs]
            n += len(recs)
        return n

    def get_similarity_stats(self):
        cap = type(self).captured
        total = sum(len(v) for v in cap.values())
        return {"total_similarities": total, "unique_shows": len(cap),
                "avg_similarities_per_show": (total / len(cap)) if cap else 0.0,
                "last_computed": None}


def run_populate(catalogue, genre_weight=0.4, text_weight=0.5, metadata_weight=0.1,
                 top_n_per_show=20, min_similarity=0.1):
    """``scripts/populate_database.compute_and_store_similarities`` (:85-259), unmodified, on the
    files of ``catalogue``; returns (captured dict, stats dict)."""
    _prepare()
    pop = importlib.import_module("scripts.populate_database")
    _CapturingRepo.captured = {}
    with tempfile.TemporaryDirectory() as tmp:
        catalogue.save(tmp)
        with patch("tvbingefriend_recommendation_service.repos.SimilarityRepository", _CapturingRepo), \
                patch("tvbingefriend_recommendation_service.models.database.SessionLocal", MagicMock()), \
                patch("tvbingefriend_recommendation_service.models.ShowSimilarity", MagicMock()):
            stats = pop.compute_and_store_similarities(
                input_dir=Path(tmp), genre_weight=genre_weight, text_weight=text_weight,
                metadata_weight=metadata_weight, top_n_per_show=top_n_per_show,
                min_similarity=min_similarity)
    return dict(_CapturingRepo.captured), stats


def run_service_matrix(similarities: dict, show_ids, show_id_queries, n=10, min_similarity=0.0,
                       genre_weight=0.4, text_weight=0.5, metadata_weight=0.1):
    """``ContentBasedRecommendationService.get_recommendations_from_matrix``
    (services/content_based_service.py:161-236), unmodified, on saved N x N ``.npy`` files."""
    import numpy as np

    _prepare()
    svc_mod = importlib.import_module(
        "tvbingefriend_recommendation_service.services.content_based_service")
    with tempfile.TemporaryDirectory() as tmp:
        d = Path(tmp)
        for name in ("genre_similarity", "text_similarity", "metadata_similarity"):
            np.save(d / f"{name}.npy", similarities[name])
        with open(d / "shows_metadata.csv", "w") as fh:
            fh.write("id,name\n")
            for sid in list(show_ids):
                fh.write(f"{int(sid)},Show {int(sid)}\n")
        svc = svc_mod.ContentBasedRecommendationService(
            processed_data_dir=d, genre_weight=genre_weight, text_weight=text_weight,
            metadata_weight=metadata_weight, use_blob=False)
        return {int(q): svc.get_recommendations_from_matrix(int(q), n=n, min_similarity=min_similarity)
                for q in show_id_queries}
