"""CPU restatement of the reference's three hot-path variants.  Test infrastructure only.

Variant A: ``SimilarityComputer``            (ml/similarity_computer.py:30-190)
Variant B: production per-show loop           (scripts/populate_database.py:170-218)
Variant C: service matrix path                (services/content_based_service.py:113-140,161-236,
                                               262-308)

All arithmetic is float64 (the reference promotes int64/bool/float64 inputs to float64 inside
``cosine_similarity``); nothing here is vectorised differently from the reference where that
could change a result: variant B is the same five cosines, the same ``(p + t + l) / 3``, the
same raw-weight sum, the same ``np.argsort(...)[::-1]`` walk with self-skip / threshold / top-N.
"""

from __future__ import annotations

from typing import Callable, Iterable

import numpy as np
import scipy.sparse as sp

from .cosine import cosine_similarity as _np_cosine
from .cosine import normalize_rows


# --------------------------------------------------------------------------------------------
# Variant A -- ml/similarity_computer.py
# --------------------------------------------------------------------------------------------
class SimilarityComputerOracle:
    """Restates ``SimilarityComputer`` (ml/similarity_computer.py:12-190) without logging."""

    def __init__(self, genre_weight: float = 0.4, text_weight: float = 0.5,
                 metadata_weight: float = 0.1, cosine: Callable = _np_cosine):
        # similarity_computer.py:15-28
        self.genre_weight = genre_weight
        self.text_weight = text_weight
        self.metadata_weight = metadata_weight
        self._cos = cosine

    def compute_genre_similarity(self, genre_features):  # :30-45
        return self._cos(genre_features)

    def compute_text_similarity(self, text_features):  # :47-62
        return self._cos(text_features)

    def compute_metadata_similarity(self, platform_features, type_features, language_features):
        # :64-90 -- hstack first, one cosine over the concatenation
        return self._cos(np.hstack([platform_features, type_features, language_features]))

    def compute_hybrid_similarity(self, genre_similarity, text_similarity, metadata_similarity):
        # :92-130 -- weights normalised by their sum
        total = self.genre_weight + self.text_weight + self.metadata_weight
        gw = self.genre_weight / total
        tw = self.text_weight / total
        mw = self.metadata_weight / total
        return gw * genre_similarity + tw * text_similarity + mw * metadata_similarity

    def compute_all_similarities(self, features: dict) -> dict:  # :132-169
        g = self.compute_genre_similarity(features["genre_features"])
        t = self.compute_text_similarity(features["text_features"])
        m = self.compute_metadata_similarity(features["platform_features"],
                                             features["type_features"],
                                             features["language_features"])
        h = self.compute_hybrid_similarity(g, t, m)
        return {"genre_similarity": g, "text_similarity": t,
                "metadata_similarity": m, "hybrid_similarity": h}

    @staticmethod
    def get_similarity_statistics(similarity_matrix) -> dict:  # :171-190
        ut = similarity_matrix[np.triu_indices_from(similarity_matrix, k=1)]
        return {"mean": float(ut.mean()), "std": float(ut.std()), "min": float(ut.min()),
                "max": float(ut.max()), "median": float(np.median(ut))}


# --------------------------------------------------------------------------------------------
# Variant B -- scripts/populate_database.py:170-218
# --------------------------------------------------------------------------------------------
def _select(hybrid_sim, idx, top_n, min_similarity, genre_sim, text_sim, metadata_sim, show_ids):
    """populate_database.py:195-218: full descending argsort, skip self, stop below threshold."""
    top_indices = np.argsort(hybrid_sim)[::-1]
    recs = []
    for similar_idx in top_indices:
        if similar_idx == idx:
            continue
        score = float(hybrid_sim[similar_idx])
        if score < min_similarity:
            break
        if len(recs) >= top_n:
            break
        recs.append({
            "similar_show_id": show_ids[similar_idx],
            "similarity_score": score,
            "genre_score": float(genre_sim[similar_idx]),
            "text_score": float(text_sim[similar_idx]),
            "metadata_score": float(metadata_sim[similar_idx]),
        })
    return recs


def production_loop(features: dict, show_ids, genre_weight=0.4, text_weight=0.5,
                    metadata_weight=0.1, top_n_per_show=20, min_similarity=0.1,
                    rows: Iterable[int] | None = None, cosine: Callable = _np_cosine) -> dict:
    """The hot loop verbatim: every iteration re-normalises all N rows five times, exactly as the
    reference does (this is the arm that ``bench.py`` times as the CPU baseline).  ``rows`` limits
    the iteration to a sample of source shows (rows are independent)."""
    genre = features["genre_features"]
    text = features["text_features"]
    plat = features["platform_features"]
    typ = features["type_features"]
    lang = features["language_features"]
    show_ids = list(show_ids)
    out = {}
    it = range(len(show_ids)) if rows is None else rows
    for idx in it:
        idx = int(idx)
        genre_sim = cosine(genre[idx:idx + 1], genre)[0]          # :180
        text_sim = cosine(text[idx:idx + 1], text)[0]             # :181
        platform_sim = cosine(plat[idx:idx + 1], plat)[0]         # :184
        type_sim = cosine(typ[idx:idx + 1], typ)[0]               # :185
        language_sim = cosine(lang[idx:idx + 1], lang)[0]         # :186
        metadata_sim = (platform_sim + type_sim + language_sim) / 3   # :187
        hybrid_sim = (genre_weight * genre_sim + text_weight * text_sim
                      + metadata_weight * metadata_sim)           # :190-192 raw weights
        recs = _select(hybrid_sim, idx, top_n_per_show, min_similarity,
                       genre_sim, text_sim, metadata_sim, show_ids)
        if recs:                                                   # :220-221
            out[show_ids[idx]] = recs
    return out


class ProductionRows:
    """Variant B with the loop-invariant normalisations hoisted (identical values: sklearn's
    ``normalize`` is a pure per-row function, so normalising ``Y`` once or N times gives the same
    matrix).  Used for parity checks at sizes where the verbatim loop would take hours."""

    def __init__(self, features: dict, genre_weight=0.4, text_weight=0.5, metadata_weight=0.1,
                 metadata_mode: str = "mean3", normalize_weights: bool = False):
        self.gw, self.tw, self.mw = float(genre_weight), float(text_weight), float(metadata_weight)
        if normalize_weights:  # variants A / C: similarity_computer.py:112-115
            tot = self.gw + self.tw + self.mw
            self.gw, self.tw, self.mw = self.gw / tot, self.tw / tot, self.mw / tot
        self.G = normalize_rows(features["genre_features"])
        T = features["text_features"]
        self.T = normalize_rows(T if sp.issparse(T) else sp.csr_matrix(np.asarray(T, dtype=np.float64)))
        self.Tt = sp.csr_matrix(self.T.T)
        self.mode = metadata_mode
        if metadata_mode == "mean3":      # populate_database.py:184-187
            self.M = [normalize_rows(features[k]) for k in
                      ("platform_features", "type_features", "language_features")]
        elif metadata_mode == "hstack":   # similarity_computer.py:84-86
            self.M = [normalize_rows(np.hstack([features["platform_features"],
                                                features["type_features"],
                                                features["language_features"]]))]
        else:
            raise ValueError(metadata_mode)
        self.n = self.G.shape[0]

    def row(self, idx: int):
        """(hybrid, genre, text, metadata) float64 rows of show ``idx`` against all shows."""
        g = (self.G[idx:idx + 1] @ self.G.T)[0]
        t = np.asarray((self.T[idx:idx + 1] @ self.Tt).todense()).ravel()
        if self.mode == "mean3":
            p, ty, la = [(m[idx:idx + 1] @ m.T)[0] for m in self.M]
            md = (p + ty + la) / 3
        else:
            md = (self.M[0][idx:idx + 1] @ self.M[0].T)[0]
        h = self.gw * g + self.tw * t + self.mw * md
        return h, g, t, md

    def pair_scores(self, idx: int, js) -> np.ndarray:
        """Reference hybrid score of (idx, j) for arbitrary j -- used by the comparator."""
        return self.row(idx)[0][np.asarray(js, dtype=np.int64)]

    def rows_block(self, rows):
        """``row`` for a block of source shows at once: four float64 [B, N] arrays.  Same
        expressions; the dense products go through one BLAS call per block instead of one per row
        (last-ulp differences at most, far below the comparator's tolerance)."""
        rows = np.asarray(rows, dtype=np.int64)
        g = np.asarray(self.G[rows] @ self.G.T)
        t = np.asarray((self.T[rows] @ self.Tt).todense())
        if self.mode == "mean3":
            p, ty, la = [np.asarray(m[rows] @ m.T) for m in self.M]
            md = (p + ty + la) / 3
        else:
            md = np.asarray(self.M[0][rows] @ self.M[0].T)
        h = self.gw * g + self.tw * t + self.mw * md
        return h, g, t, md

    def topk_arrays(self, rows, k=20, min_similarity=0.1, block: int = 0):
        """Top-K of the given source rows as arrays: indices [R,k] (-1 padded), counts [R],
        and the four score arrays [4,R,k] (NaN padded).  ``block`` > 0 scores that many source
        rows per BLAS / CSR product (full-catalogue checks); the selection is the reference's
        per-row ``argsort()[::-1]`` walk either way."""
        rows = np.asarray(list(rows), dtype=np.int64)
        R = rows.shape[0]
        idx = np.full((R, k), -1, dtype=np.int64)
        cnt = np.zeros(R, dtype=np.int64)
        sc = np.full((4, R, k), np.nan, dtype=np.float64)
        step = block if block > 0 else 1
        for r0 in range(0, R, step):
            chunk = rows[r0:r0 + step]
            if block > 0:
                hb, gb, tb, mb = self.rows_block(chunk)
            for q, i in enumerate(chunk.tolist()):
                r = r0 + q
                h, g, t, md = (hb[q], gb[q], tb[q], mb[q]) if block > 0 else self.row(i)
                order = np.argsort(h)[::-1]
                c = 0
                for j in order:
                    if j == i:
                        continue
                    if h[j] < min_similarity or c >= k:
                        break
                    idx[r, c] = j
                    sc[0, r, c], sc[1, r, c], sc[2, r, c], sc[3, r, c] = h[j], g[j], t[j], md[j]
                    c += 1
                cnt[r] = c
        return idx, cnt, sc


# --------------------------------------------------------------------------------------------
# Variant C -- services/content_based_service.py
# --------------------------------------------------------------------------------------------
def service_hybrid(genre_sim, text_sim, metadata_sim, genre_weight=0.4, text_weight=0.5,
                   metadata_weight=0.1):
    """content_based_service.py:132-138 (weights normalised by their sum)."""
    total = genre_weight + text_weight + metadata_weight
    return ((genre_weight / total) * genre_sim + (text_weight / total) * text_sim
            + (metadata_weight / total) * metadata_sim)


def service_recommendations(hybrid, genre_sim, text_sim, metadata_sim, show_ids, show_id,
                            n=10, min_similarity=0.0) -> list[dict]:
    """``get_recommendations_from_matrix`` (content_based_service.py:161-236)."""
    show_ids = list(show_ids)
    id_to_index = {sid: i for i, sid in enumerate(show_ids)}
    if show_id not in id_to_index:      # :199-201
        return []
    show_idx = id_to_index[show_id]
    scores = hybrid[show_idx]
    recs = []
    for idx in np.argsort(scores)[::-1]:   # :209
        if idx == show_idx:
            continue
        s = float(scores[idx])
        if s < min_similarity:
            break
        recs.append({"show_id": show_ids[idx], "similarity_score": s,
                     "genre_score": float(genre_sim[show_idx, idx]),
                     "text_score": float(text_sim[show_idx, idx]),
                     "metadata_score": float(metadata_sim[show_idx, idx])})
        if len(recs) >= n:
            break
    return recs


def service_all(hybrid, genre_sim, text_sim, metadata_sim, show_ids, top_n_per_show=20,
                min_similarity=0.1) -> dict:
    """The loop of ``compute_and_store_all_similarities`` (content_based_service.py:293-308)."""
    out = {}
    for sid in list(show_ids):
        recs = service_recommendations(hybrid, genre_sim, text_sim, metadata_sim, show_ids, sid,
                                       n=top_n_per_show, min_similarity=min_similarity)
        if recs:
            out[sid] = [{"similar_show_id": r["show_id"], "similarity_score": r["similarity_score"],
                         "genre_score": r["genre_score"], "text_score": r["text_score"],
                         "metadata_score": r["metadata_score"]} for r in recs]
    return out


# --------------------------------------------------------------------------------------------
# dict <-> array helpers
# --------------------------------------------------------------------------------------------
def dict_to_arrays(all_similarities: dict, show_ids, k: int):
    """Turn the a9/a12 dict into (indices [N,k] -1 padded, counts [N], scores [4,N,k])."""
    show_ids = list(show_ids)
    pos = {sid: i for i, sid in enumerate(show_ids)}
    n = len(show_ids)
    idx = np.full((n, k), -1, dtype=np.int64)
    cnt = np.zeros(n, dtype=np.int64)
    sc = np.full((4, n, k), np.nan, dtype=np.float64)
    for sid, recs in all_similarities.items():
        i = pos[sid]
        cnt[i] = len(recs)
        for c, r in enumerate(recs):
            idx[i, c] = pos[r["similar_show_id"]]
            sc[0, i, c] = r["similarity_score"]
            sc[1, i, c] = r["genre_score"]
            sc[2, i, c] = r["text_score"]
            sc[3, i, c] = r["metadata_score"]
    return idx, cnt, sc
