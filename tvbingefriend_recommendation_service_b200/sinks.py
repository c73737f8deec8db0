"""Sinks for the computed similarity records.

The reference persists through ``SimilarityRepository.bulk_store_all_similarities``
(repos/similarity_repository.py:72-124, SQLAlchemy/MySQL) -- storage is outside this path, so
the drivers take any object with that method (a real ``SimilarityRepository(db)`` works as is).
Two in-process sinks are provided; both report the statistics keys of
``SimilarityRepository.get_similarity_stats`` (:239-263: ``total_records``, ``unique_shows``,
``avg_similarities_per_show``, ``last_computed``):

* ``InMemorySimilaritySink``  -- dict of record lists, exactly what the reference hands over;
* ``ColumnarSimilaritySink``  -- flat numpy columns (one entry per kept pair).  A sink that offers
  ``bulk_store_records(records, clear_existing)`` is fed the columnar batches of ``TopK.records``
  directly, so a 250 k-show catalogue never builds 5 M Python dicts.
"""

from __future__ import annotations

import numpy as np

RECORD_KEYS = ("similar_show_id", "similarity_score", "genre_score", "text_score", "metadata_score")
COLUMNS = ("show_id",) + RECORD_KEYS


class InMemorySimilaritySink:
    def __init__(self):
        self.records: dict = {}

    def bulk_store_all_similarities(self, all_similarities: dict, batch_size: int = 1000,
                                    clear_existing: bool = True) -> int:
        if clear_existing:
            self.records = {}
        total = 0
        for show_id, recs in all_similarities.items():
            self.records[show_id] = list(recs)
            total += len(recs)
        return total

    def get_similarity_stats(self) -> dict:
        total = sum(len(v) for v in self.records.values())
        shows = len(self.records)
        return {"total_records": total, "total_similarities": total, "unique_shows": shows,
                "avg_similarities_per_show": (total / shows) if shows else 0,
                "last_computed": None}


class ColumnarSimilaritySink:
    """Keeps the record stream as columns: ``show_id``, ``similar_show_id`` (int64) and the four
    float64 scores -- the row shape of models/show_similarity.py:10-35 without the ORM."""

    def __init__(self):
        self._parts: list[dict] = []

    def bulk_store_records(self, records: dict, clear_existing: bool = False) -> int:
        if clear_existing:
            self._parts = []
        n = int(len(records["show_id"]))
        if n:
            self._parts.append({c: np.asarray(records[c]) for c in COLUMNS})
        return n

    def bulk_store_all_similarities(self, all_similarities: dict, batch_size: int = 1000,
                                    clear_existing: bool = True) -> int:
        """Same signature as the repository's method (dict of record lists)."""
        rows = [(sid, *(r[k] for k in RECORD_KEYS)) for sid, recs in all_similarities.items() for r in recs]
        if clear_existing:
            self._parts = []
        if not rows:
            return 0
        cols = list(zip(*rows))
        return self.bulk_store_records({c: np.asarray(v) for c, v in zip(COLUMNS, cols)})

    def columns(self) -> dict:
        if not self._parts:
            return {c: np.zeros(0, dtype=np.int64 if c.endswith("show_id") else np.float64) for c in COLUMNS}
        return {c: np.concatenate([p[c] for p in self._parts]) for c in COLUMNS}

    def get_similarity_stats(self) -> dict:
        total = sum(len(p["show_id"]) for p in self._parts)
        shows = len(np.unique(np.concatenate([p["show_id"] for p in self._parts]))) if self._parts else 0
        return {"total_records": total, "total_similarities": total, "unique_shows": shows,
                "avg_similarities_per_show": (total / shows) if shows else 0, "last_computed": None}
