"""Drop-in ``ContentBasedRecommendationService`` (reference:
services/content_based_service.py:23-403) for the similarity path, on B200.

Same constructor and method signatures.  What differs is only where the arithmetic runs:

* matrix mode  -- if ``genre/text/metadata_similarity.npy`` exist in the data directory (what the
  reference requires, :123-126) they are uploaded once, the hybrid is combined on the GPU with
  the normalised weights (:132-138) and rows are ranked by the exact row kernel;
* feature mode -- otherwise the five feature files of compute_features.py are used directly
  (the production pipeline no longer writes N x N files, scripts/compute_similarities.py:152-157)
  with the conventions those matrices would have had: hstack metadata, normalised weights.

Storage (MySQL repositories, Azure blobs) is out of scope: pass any object with
``bulk_store_all_similarities`` / ``get_similarity_stats`` as ``sink`` (default: in memory).
"""

from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import torch

from ..engine import HybridTopKEngine, default_engine
from ..sinks import InMemorySimilaritySink

logger = logging.getLogger(__name__)

_FEATURE_FILES = ("genre_features.npy", "text_features.npz", "platform_features.npy",
                  "type_features.npy", "language_features.npy")
_MATRIX_FILES = ("genre_similarity.npy", "text_similarity.npy", "metadata_similarity.npy")


def load_feature_files(data_dir: Path) -> dict:
    """The on-disk contract of compute_features.py:115-129 / populate_database.py:125-131."""
    from scipy.sparse import load_npz

    data_dir = Path(data_dir)
    return {"genre_features": np.load(data_dir / "genre_features.npy"),
            "text_features": load_npz(data_dir / "text_features.npz"),
            "platform_features": np.load(data_dir / "platform_features.npy"),
            "type_features": np.load(data_dir / "type_features.npy"),
            "language_features": np.load(data_dir / "language_features.npy")}


def load_show_ids(data_dir: Path) -> list:
    """``id`` column of shows_metadata.csv; row order = matrix row order (populate_database.py:136-137)."""
    import pandas as pd

    return pd.read_csv(Path(data_dir) / "shows_metadata.csv")["id"].tolist()


# noinspection PyMethodMayBeStatic
class ContentBasedRecommendationService:
    """Content-based TV show recommendations served from GPU-computed similarities."""

    def __init__(self, processed_data_dir: Path | None = None, genre_weight: float = 0.4,
                 text_weight: float = 0.5, metadata_weight: float = 0.1,
                 use_blob: bool | None = None, blob_prefix: str = "processed",
                 engine: HybridTopKEngine | None = None, sink=None):
        if processed_data_dir is None:
            project_root = Path(__file__).resolve().parent.parent.parent
            processed_data_dir = project_root / "data" / "processed"
        self.processed_data_dir = Path(processed_data_dir)
        self.genre_weight = genre_weight
        self.text_weight = text_weight
        self.metadata_weight = metadata_weight
        self.blob_prefix = blob_prefix
        self.use_blob = bool(use_blob)
        if self.use_blob:
            raise RuntimeError("Azure blob storage is outside the scope of this path: download the "
                               "processed files first and pass processed_data_dir")
        self._engine = engine
        self.sink = sink if sink is not None else InMemorySimilaritySink()
        self._mode: str | None = None
        self._matrices: dict | None = None          # matrix mode: device tensors
        self._catalogue = None                      # feature mode: DeviceCatalogue
        self._show_id_to_index: dict | None = None
        self._index_to_show_id: dict | None = None
        logger.info("Initialized ContentBasedRecommendationService")
        logger.info(f"Weights - Genre: {genre_weight}, Text: {text_weight}, Metadata: {metadata_weight}")

    @property
    def engine(self) -> HybridTopKEngine:
        if self._engine is None:
            self._engine = default_engine()
        return self._engine

    def _get_data_dir(self) -> Path:
        return self.processed_data_dir

    def _weights(self) -> tuple[float, float, float]:
        total_weight = self.genre_weight + self.text_weight + self.metadata_weight   # reference :132
        return (self.genre_weight / total_weight, self.text_weight / total_weight,
                self.metadata_weight / total_weight)

    def _load_similarity_matrices(self):
        """reference :113-140"""
        if self._mode is not None:
            return
        data_dir = self._get_data_dir()
        if all((data_dir / f).exists() for f in _MATRIX_FILES):
            dev = self.engine.device
            g, t, m = (torch.from_numpy(np.ascontiguousarray(np.load(data_dir / f), dtype=np.float64)).to(dev)
                       for f in _MATRIX_FILES)
            gw, tw, mw = self._weights()
            self._matrices = {"genre": g, "text": t, "metadata": m,
                              "hybrid": self.engine.hybrid_combine(g, t, m, gw, tw, mw)}
            self._mode = "matrix"
            logger.info(f"✓ Loaded similarity matrices: {tuple(g.shape)}")
        elif all((data_dir / f).exists() for f in _FEATURE_FILES):
            feats = load_feature_files(data_dir)
            self._catalogue = self.engine.ingest(feats, "hstack", self._weights())
            self._mode = "features"
            logger.info(f"✓ Prepared features for {self._catalogue.n_shows} shows on {self.engine.device}")
        else:
            raise FileNotFoundError(
                f"Neither similarity matrices {_MATRIX_FILES} nor feature files {_FEATURE_FILES} "
                f"found in {data_dir}")

    def _load_show_mappings(self):
        """reference :142-159"""
        if self._show_id_to_index is not None:
            return
        ids = load_show_ids(self._get_data_dir())
        self._show_id_to_index = {show_id: idx for idx, show_id in enumerate(ids)}
        self._index_to_show_id = {idx: show_id for show_id, idx in self._show_id_to_index.items()}
        logger.info(f"✓ Loaded mappings for {len(self._show_id_to_index)} shows")

    def _rows_topk(self, rows, n: int, min_similarity: float):
        if self._mode == "matrix":
            mats = self._matrices
            return self.engine.matrix_rows_topk(mats["hybrid"], mats["genre"], mats["text"], mats["metadata"],
                                                rows, k=n, min_similarity=min_similarity)
        return self.engine.exact_rows(self._catalogue, rows, self._weights(), k=n, min_similarity=min_similarity)

    def get_recommendations_from_matrix(self, show_id: int, n: int = 10, min_similarity: float = 0.0) -> list[dict]:
        """reference :161-236 -- descending score, self skipped, ``score < min_similarity`` cut,
        at most ``n``; unknown id -> []."""
        self._load_similarity_matrices()
        self._load_show_mappings()
        if show_id not in self._show_id_to_index:
            logger.warning(f"Show ID {show_id} not found in similarity matrix")
            return []
        if n <= 0:
            return []
        top = self._rows_topk([self._show_id_to_index[show_id]], n, min_similarity)
        return [{"show_id": self._index_to_show_id[int(top.indices[0, e])],
                 "similarity_score": float(top.hybrid[0, e]), "genre_score": float(top.genre[0, e]),
                 "text_score": float(top.text[0, e]), "metadata_score": float(top.metadata[0, e])}
                for e in range(int(top.counts[0]))]

    def compute_and_store_all_similarities(self, top_n_per_show: int = 20, min_similarity: float = 0.1) -> dict:
        """reference :262-338"""
        logger.info("=" * 60)
        logger.info("COMPUTING AND STORING ALL SIMILARITIES")
        logger.info("=" * 60)
        self._load_similarity_matrices()
        self._load_show_mappings()
        ids = [self._index_to_show_id[i] for i in range(len(self._index_to_show_id))]
        logger.info(f"Computing similarities for {len(ids)} shows...")
        if self._mode == "matrix":
            top = self._rows_topk(np.arange(len(ids), dtype=np.int32), top_n_per_show, min_similarity)
        else:
            top = self.engine.to_host(self.engine.top_k_device(self._catalogue, self._weights(), top_n_per_show,
                                                               min_similarity, True))
        all_similarities = top.to_dict(ids, id_key="similar_show_id")
        logger.info(f"✓ Computed similarities for {len(all_similarities)} shows")
        logger.info("Storing similarities...")
        total_records = self.sink.bulk_store_all_similarities(all_similarities)
        stats = dict(self.sink.get_similarity_stats())
        stats["computed_shows"] = len(all_similarities)
        stats["top_n_per_show"] = top_n_per_show
        stats["min_similarity"] = min_similarity
        logger.info(f"Total records stored: {total_records}")
        return stats

    # ---- storage-side methods of the reference: delegated, never computed here -------------------
    def get_recommendations_from_db(self, show_id: int, n: int = 10, min_similarity: float = 0.0) -> list[dict]:
        """reference :238-260 reads MySQL; here: whatever the sink stored."""
        getter = getattr(self.sink, "get_similar_shows_with_metadata", None)
        if getter is not None:
            return getter(show_id=show_id, n=n, min_similarity=min_similarity)
        recs = getattr(self.sink, "records", {}).get(show_id, [])
        return [r for r in recs if r["similarity_score"] >= min_similarity][:n]

    def sync_metadata_to_db(self, shows_data: list[dict]) -> int:
        """reference :340-381 (MetadataRepository) -- storage, out of scope unless the sink has it."""
        store = getattr(self.sink, "bulk_store_shows", None)
        if store is None:
            raise NotImplementedError("metadata persistence is outside the similarity path; "
                                      "pass a sink with bulk_store_shows")
        return store(shows_data)

    def get_stats(self) -> dict:
        """reference :383-403"""
        return {"similarity_stats": self.sink.get_similarity_stats(),
                "cached_shows": len(self._show_id_to_index or {}),
                "weights": {"genre": self.genre_weight, "text": self.text_weight,
                            "metadata": self.metadata_weight}}
