"""Deterministic synthetic show catalogues of the BASELINE.json shapes.

The reference has no data generator: its inputs are the five feature matrices that
``scripts/compute_features.py:115-138`` writes (``genre_features.npy`` int64 multi-hot,
``text_features.npz`` CSR float64 L2-normalised TF-IDF, ``platform/type/language_features.npy``
one-hot) plus the ``id`` column of ``shows_metadata.csv`` (populate_database.py:125-137).
This module produces catalogues with exactly those types and row properties
 from ``numpy.random.default_rng(seed)``:

* text: Zipf-like column draw (p ~ rank^-0.8), values U(0.1, 1.1), duplicates summed, rows
  L2-normalised, 1 % empty rows, 0.5 % planted exact duplicates of earlier shows (tie stress);
* genre: int64 Bernoulli(0.06) per bit (about 8 % all-zero rows at G=40);
* platform / language: float64 one-hot from a Zipf category draw; type: bool one-hot with 2 %
  all-False rows (``pd.get_dummies`` on a missing type, feature_extractor.py:157-158);
* show ids: a random permutation of non-contiguous positive ints (TVmaze ids are sparse).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

# name -> (N, V, mean nnz/row, G, (P, T, L), K, seed)   
CONFIGS: dict[str, dict] = {
    "C1": dict(n_shows=1_000, vocab=5_000, nnz=40, n_genres=40, meta=(5, 3, 2), k=20, seed=20261),
    "C2": dict(n_shows=20_000, vocab=5_000, nnz=40, n_genres=40, meta=(5, 3, 2), k=20, seed=20262),
    "C3": dict(n_shows=100_000, vocab=10_000, nnz=50, n_genres=40, meta=(5, 3, 2), k=20, seed=20263),
    "C4": dict(n_shows=250_000, vocab=10_000, nnz=50, n_genres=40, meta=(21, 5, 6), k=20, seed=20264),
    "C5": dict(n_shows=200_000, vocab=50_000, nnz=60, n_genres=40, meta=(21, 5, 6), k=100, seed=20265),
    # the reference's real production run: ~80 k shows, max_text_features default 500
    # (scripts/compute_features.py:174; README.md:254-258 "60-90 min"), ~96 % sparse text (notebook 02),
    # reference default metadata widths 21 / 5 / 6
    "P80k": dict(n_shows=80_000, vocab=500, nnz=20, n_genres=40, meta=(21, 5, 6), k=20, seed=20266),
}

# Weight schemes of the reference's notebook (notebooks/03_content_similarity cell 6), used by C5.
WEIGHT_SWEEP = [(0.4, 0.5, 0.1), (0.3, 0.6, 0.1), (0.6, 0.3, 0.1), (1.0, 1.0, 1.0), (0.5, 0.5, 0.0)]


@dataclass
class Catalogue:
    """The five feature matrices + show ids, typed as the reference's files are."""

    genre_features: np.ndarray  # int64 [N, G]
    text_features: sp.csr_matrix  # float64 [N, V]
    platform_features: np.ndarray  # float64 [N, P]
    type_features: np.ndarray  # bool [N, T]
    language_features: np.ndarray  # float64 [N, L]
    show_ids: np.ndarray  # int64 [N]

    def features(self) -> dict:
        """The dict ``SimilarityComputer.compute_all_similarities`` takes (similarity_computer.py:132)."""
        return {
            "genre_features": self.genre_features,
            "text_features": self.text_features,
            "platform_features": self.platform_features,
            "type_features": self.type_features,
            "language_features": self.language_features,
        }

    @property
    def n_shows(self) -> int:
        return int(self.genre_features.shape[0])

    def save(self, directory) -> None:
        """Write the on-disk layout populate_database.py:125-137 reads."""
        from pathlib import Path

        d = Path(directory)
        d.mkdir(parents=True, exist_ok=True)
        np.save(d / "genre_features.npy", self.genre_features)
        sp.save_npz(d / "text_features.npz", self.text_features)
        np.save(d / "platform_features.npy", self.platform_features)
        np.save(d / "type_features.npy", self.type_features)
        np.save(d / "language_features.npy", self.language_features)
        with open(d / "shows_metadata.csv", "w") as fh:
            fh.write("id,name\n")
            for sid in self.show_ids.tolist():
                fh.write(f"{sid},Show {sid}\n")


def _zipf_probs(n: int, exponent: float) -> np.ndarray:
    p = np.arange(1, n + 1, dtype=np.float64) ** (-exponent)
    return p / p.sum()


def _one_hot(codes: np.ndarray, width: int, dtype) -> np.ndarray:
    out = np.zeros((codes.shape[0], width), dtype=dtype)
    valid = codes >= 0
    out[np.nonzero(valid)[0], codes[valid]] = 1
    return out


def make_text(rng: np.random.Generator, n: int, vocab: int, mean_nnz: int,
              empty_frac: float = 0.01, dup_frac: float = 0.005) -> sp.csr_matrix:
    """CSR float64 TF-IDF-like matrix, rows L2-normalised (TfidfVectorizer norm='l2')."""
    counts = rng.poisson(mean_nnz, size=n).astype(np.int64)
    counts = np.clip(counts, 1, vocab)
    counts[rng.random(n) < empty_frac] = 0
    total = int(counts.sum())
    cdf = np.cumsum(_zipf_probs(vocab, 0.8))
    cols = np.searchsorted(cdf, rng.random(total), side="right").astype(np.int64)
    np.clip(cols, 0, vocab - 1, out=cols)
    vals = rng.uniform(0.1, 1.1, size=total)
    rows = np.repeat(np.arange(n, dtype=np.int64), counts)
    m = sp.coo_matrix((vals, (rows, cols)), shape=(n, vocab)).tocsr()  # duplicates summed
    m.sum_duplicates()
    m.sort_indices()
    norms = np.sqrt(np.asarray(m.multiply(m).sum(axis=1)).ravel())
    norms[norms == 0] = 1.0
    m = sp.diags(1.0 / norms) @ m
    m = sp.csr_matrix(m, dtype=np.float64)
    m.sort_indices()
    n_dup = int(round(dup_frac * n))
    if n_dup and n > 2 * n_dup:
        # later rows copy earlier rows exactly -> exact score ties and 1.0 text neighbours
        dst = rng.choice(np.arange(n // 2, n), size=n_dup, replace=False)
        src = rng.integers(0, n // 2, size=n_dup)
        take = np.arange(n, dtype=np.int64)
        take[dst] = src
        m = sp.csr_matrix(m[take], dtype=np.float64)
        m.sort_indices()
    return m


def make_catalogue(n_shows: int, vocab: int, nnz: int = 40, n_genres: int = 40,
                   meta: tuple[int, int, int] = (5, 3, 2), seed: int = 0, **_ignored) -> Catalogue:
    """Build one synthetic catalogue; all randomness from ``default_rng(seed)``."""
    rng = np.random.default_rng(seed)
    text = make_text(rng, n_shows, vocab, nnz)
    genre = (rng.random((n_shows, n_genres)) < 0.06).astype(np.int64)
    n_p, n_t, n_l = meta
    p_codes = rng.choice(n_p, size=n_shows, p=_zipf_probs(n_p, 1.0))
    t_codes = rng.choice(n_t, size=n_shows, p=_zipf_probs(n_t, 1.0))
    t_codes[rng.random(n_shows) < 0.02] = -1  # missing type -> all-False row
    l_codes = rng.choice(n_l, size=n_shows, p=_zipf_probs(n_l, 1.5))
    ids = np.sort(rng.choice(np.arange(1, 8 * n_shows + 64), size=n_shows, replace=False))
    ids = ids[rng.permutation(n_shows)].astype(np.int64)
    return Catalogue(
        genre_features=genre,
        text_features=text,
        platform_features=_one_hot(p_codes, n_p, np.float64),
        type_features=_one_hot(t_codes, n_t, np.bool_),
        language_features=_one_hot(l_codes, n_l, np.float64),
        show_ids=ids,
    )


def make_config(name: str, n_shows: int | None = None) -> Catalogue:
    """One of C1..C5; ``n_shows`` overrides N (same seed, same column statistics)."""
    cfg = dict(CONFIGS[name])
    if n_shows is not None:
        cfg["n_shows"] = int(n_shows)
    return make_catalogue(**cfg)
