"""Sinks for the computed similarity records.

The reference persists through ``SimilarityRepository.bulk_store_all_similarities``
(repos/similarity_repository.py:72-124, SQLAlchemy/MySQL) -- storage is outside this path, so
the drivers take any object with that method; ``InMemorySimilaritySink`` is the default and
mirrors the statistics ``get_similarity_stats`` reports (:214-263 keys used by the callers:
``unique_shows``, ``avg_similarities_per_show``).
"""

from __future__ import annotations


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
        return {"total_similarities": total, "unique_shows": shows,
                "avg_similarities_per_show": (total / shows) if shows else 0.0,
                "last_computed": None}
