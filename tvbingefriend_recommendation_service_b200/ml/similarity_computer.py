"""Drop-in ``SimilarityComputer`` (reference: ml/similarity_computer.py:12-190) on B200.

Same constructor, methods, defaults and return types as the reference class; every method runs
CUDA kernels of ``libtvbf.so`` and raises if the library or an sm_100 GPU is missing.  The
additive ``compute_top_k`` is the production path: features -> hybrid score -> per-show top-K
without materialising any N x N matrix.
"""

from __future__ import annotations

import logging

import numpy as np
import torch

from ..engine import HybridTopKEngine, TopK, default_engine

logger = logging.getLogger(__name__)


# noinspection PyMethodMayBeStatic
class SimilarityComputer:
    """Compute similarity matrices (and top-K tables) from feature arrays."""

    def __init__(self, genre_weight: float = 0.4, text_weight: float = 0.5,
                 metadata_weight: float = 0.1, engine: HybridTopKEngine | None = None):
        # reference :15-28
        self.genre_weight = genre_weight
        self.text_weight = text_weight
        self.metadata_weight = metadata_weight
        self._engine = engine

    @property
    def engine(self) -> HybridTopKEngine:
        if self._engine is None:
            self._engine = default_engine()
        return self._engine

    # ---- N x N variant (reference :30-130) ------------------------------------------------------
    def _cosine(self, x, label: str) -> torch.Tensor:
        logger.info(f"Computing {label} similarity...")
        sim = self.engine.cosine_matrix(x)
        logger.info(f" {label.capitalize()} similarity: {tuple(sim.shape)}")
        return sim

    @staticmethod
    def _out_dtype(*arrays):
        """sklearn's dtype rule (check_pairwise_arrays / _return_float_dtype):
        the result is float32 only when EVERY input is float32, else float64.  The arithmetic here is
        float64 either way (more precise than the reference's float32 run, within 1e-6 of it)."""
        return np.float32 if all(getattr(a, "dtype", None) == np.float32 for a in arrays) else np.float64

    def compute_genre_similarity(self, genre_features: np.ndarray) -> np.ndarray:
        """cosine_similarity(genre_features) -> (n_shows, n_shows) float64 (reference :30-45)."""
        return self._cosine(genre_features, "genre").cpu().numpy().astype(self._out_dtype(genre_features), copy=False)

    def compute_text_similarity(self, text_features) -> np.ndarray:
        """cosine_similarity on TF-IDF vectors, dense or scipy sparse (reference :47-62)."""
        return self._cosine(text_features, "text").cpu().numpy().astype(self._out_dtype(text_features), copy=False)

    def compute_metadata_similarity(self, platform_features: np.ndarray, type_features: np.ndarray,
                                    language_features: np.ndarray) -> np.ndarray:
        """cosine of the hstack of platform/type/language (reference :64-90)."""
        metadata_features = np.hstack([platform_features, type_features, language_features])
        return self._cosine(metadata_features, "metadata").cpu().numpy().astype(
            self._out_dtype(metadata_features), copy=False)

    def _normalized_weights(self) -> tuple[float, float, float]:
        total_weight = self.genre_weight + self.text_weight + self.metadata_weight  # reference :112-115
        return (self.genre_weight / total_weight, self.text_weight / total_weight,
                self.metadata_weight / total_weight)

    def compute_hybrid_similarity(self, genre_similarity: np.ndarray, text_similarity: np.ndarray,
                                  metadata_similarity: np.ndarray) -> np.ndarray:
        """Weighted combination with weights normalised by their sum (reference :92-130)."""
        logger.info("Computing hybrid similarity...")
        gw, tw, mw = self._normalized_weights()
        logger.info(f"  Weights - Genre: {gw:.2f}, Text: {tw:.2f}, Metadata: {mw:.2f}")
        dev = self.engine.device
        g, t, m = (torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
                   for a in (genre_similarity, text_similarity, metadata_similarity))
        return self.engine.hybrid_combine(g, t, m, gw, tw, mw).cpu().numpy()

    def compute_all_similarities(self, features: dict[str, np.ndarray]) -> dict[str, np.ndarray]:
        """All four matrices (reference :132-169); intermediates stay on the GPU."""
        logger.info("=" * 60)
        logger.info("COMPUTING ALL SIMILARITIES")
        logger.info("=" * 60)
        g = self._cosine(features["genre_features"], "genre")
        t = self._cosine(features["text_features"], "text")
        m = self._cosine(np.hstack([features["platform_features"], features["type_features"],
                                    features["language_features"]]), "metadata")
        gw, tw, mw = self._normalized_weights()
        h = self.engine.hybrid_combine(g, t, m, gw, tw, mw)
        logger.info("=" * 60)
        logger.info("SIMILARITY COMPUTATION COMPLETE")
        logger.info("=" * 60)
        return {"genre_similarity": g.cpu().numpy(), "text_similarity": t.cpu().numpy(),
                "metadata_similarity": m.cpu().numpy(), "hybrid_similarity": h.cpu().numpy()}

    def get_similarity_statistics(self, similarity_matrix: np.ndarray) -> dict[str, float]:
        """mean/std/min/max/median of the strict upper triangle (reference :171-190)."""
        mat = torch.from_numpy(np.ascontiguousarray(similarity_matrix, dtype=np.float64)).to(self.engine.device)
        return self.engine.matrix_stats(mat)

    def compute_similarity_statistics(self, features: dict, metadata_mode: str = "hstack",
                                      normalize_weights: bool = True) -> dict[str, dict[str, float]]:
        """The four ``get_similarity_statistics`` results of scripts/compute_similarities.py:119-131
        for catalogues whose N x N matrices cannot exist: one streaming tensor-core sweep (defaults =
        this class' conventions: hstack metadata, normalised weights).  mean / std / min / max as
        the reference's, median to ``median_resolution``.  Needs binary genre and one-hot metadata
        features."""
        weights = self._normalized_weights() if normalize_weights else \
            (self.genre_weight, self.text_weight, self.metadata_weight)
        cat = self.engine.ingest(features, metadata_mode, weights)
        return self.engine.similarity_stats(cat, weights)

    # ---- production variant: no N x N ------------------------------------------------------------
    def compute_top_k(self, features: dict, k: int = 20, min_similarity: float = 0.1,
                      exclude_self: bool = True, metadata_mode: str = "mean3",
                      normalize_weights: bool = False, device_ids=None, **kw) -> TopK:
        """features -> per-show top-K table.

        Defaults reproduce the production loop (scripts/populate_database.py:170-218): mean of the
        three metadata cosines and RAW weights.  ``metadata_mode="hstack", normalize_weights=True``
        reproduces what this class + the service produce (reference :84-86, :112-115).
        Ties are broken by (score descending, column index ascending)."""
        weights = self._normalized_weights() if normalize_weights else \
            (self.genre_weight, self.text_weight, self.metadata_weight)
        if device_ids is not None and len(device_ids) > 1:
            from ..multi_gpu import compute_top_k_multi_gpu
            return compute_top_k_multi_gpu(features, weights, k, min_similarity, metadata_mode,
                                           exclude_self, list(device_ids), **kw)
        eng = self.engine if device_ids is None else HybridTopKEngine(device_ids[0])
        return eng.compute_top_k(features, weights, k, min_similarity, metadata_mode, exclude_self, **kw)

    def compute_top_k_sweep(self, features: dict, weight_list, k: int = 20, min_similarity: float = 0.1,
                            exclude_self: bool = True, metadata_mode: str = "mean3",
                            normalize_weights: bool = False, **kw) -> list[TopK]:
        """One top-K table per (genre, text, metadata) weight triple -- the comparison of weighting
        schemes the reference does by re-running everything per scheme (notebooks/03 cell 6) -- with
        staging, upload and feature prep shared by all triples and, on large catalogues of binary /
        one-hot features, ONE tensor-core sweep for up to five triples.  Each table is identical to
        what ``SimilarityComputer(*triple).compute_top_k(...)`` returns."""
        triples = []
        for gw, tw, mw in weight_list:
            tot = float(gw) + float(tw) + float(mw)
            triples.append((gw / tot, tw / tot, mw / tot) if normalize_weights else (float(gw), float(tw), float(mw)))
        return self.engine.compute_top_k_sweep(features, triples, k, min_similarity, metadata_mode, exclude_self, **kw)
