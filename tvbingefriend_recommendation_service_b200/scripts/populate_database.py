"""Drop-in for the hot function of scripts/populate_database.py:
``compute_and_store_similarities`` (reference :85-259).

Same arguments and defaults, same record stream (``similar_show_id, similarity_score, genre_score,
text_score, metadata_score``; shows without a qualifying neighbour omitted; flushed to the sink
every 5000 shows with ``clear_existing=False``), same returned statistics keys -- but the N
iterations of five cosine_similarity calls + argsort become one GPU pass."""

from __future__ import annotations

import argparse
import logging
import sys
from pathlib import Path

from ..ml.similarity_computer import SimilarityComputer
from ..services.content_based_service import load_feature_files, load_show_ids
from ..sinks import InMemorySimilaritySink

logger = logging.getLogger(__name__)

BATCH_SIZE = 5000  # reference :166


def compute_and_store_similarities(input_dir: Path, genre_weight: float = 0.4, text_weight: float = 0.5,
                                   metadata_weight: float = 0.1, top_n_per_show: int = 20,
                                   min_similarity: float = 0.1, sink=None, device_ids=None) -> dict:
    input_dir = Path(input_dir)
    logger.info("COMPUTING AND STORING SIMILARITIES")
    logger.info(f"Genre weight: {genre_weight}  Text weight: {text_weight}  Metadata weight: {metadata_weight}")
    logger.info(f"Top N per show: {top_n_per_show}  Min similarity: {min_similarity}")
    features = load_feature_files(input_dir)
    show_ids = load_show_ids(input_dir)
    logger.info(f"✓ Loaded features for {len(show_ids)} shows")
    computer = SimilarityComputer(genre_weight=genre_weight, text_weight=text_weight,
                                  metadata_weight=metadata_weight)
    # production conventions (reference :184-192): mean of three metadata cosines, raw weights
    top = computer.compute_top_k(features, k=top_n_per_show, min_similarity=min_similarity,
                                 exclude_self=True, metadata_mode="mean3", normalize_weights=False,
                                 device_ids=device_ids)
    sink = sink if sink is not None else InMemorySimilaritySink()
    all_similarities = top.to_dict(show_ids)
    total_records_stored = 0
    batch = {}
    clear = getattr(sink, "records", None) is not None
    if clear:
        sink.bulk_store_all_similarities({}, clear_existing=True)   # reference :156-162 delete-all up front
    for n_done, show_id in enumerate(show_ids, start=1):
        recs = all_similarities.get(show_id)
        if recs:
            batch[show_id] = recs
        if n_done % BATCH_SIZE == 0 or n_done == len(show_ids):    # reference :223-234
            if batch:
                total_records_stored += sink.bulk_store_all_similarities(batch, clear_existing=False)
                batch = {}
            logger.info(f"  Processed {n_done}/{len(show_ids)} shows... (stored {total_records_stored} records)")
    stats = dict(sink.get_similarity_stats())
    stats["total_records"] = total_records_stored
    stats["top_n_per_show"] = top_n_per_show
    stats["min_similarity"] = min_similarity
    stats["flagged_rows"] = top.flagged_rows
    return stats


def main(argv=None):
    """The similarity-related flags of reference :298-401."""
    parser = argparse.ArgumentParser(description="Compute top-N similarities on B200 and store them")
    parser.add_argument("--input-dir", type=str, default="data/processed")
    parser.add_argument("--top-n", type=int, default=20)
    parser.add_argument("--min-similarity", type=float, default=0.1)
    parser.add_argument("--genre-weight", type=float, default=0.4)
    parser.add_argument("--text-weight", type=float, default=0.5)
    parser.add_argument("--metadata-weight", type=float, default=0.1)
    args = parser.parse_args(argv)
    try:
        stats = compute_and_store_similarities(Path(args.input_dir), args.genre_weight, args.text_weight,
                                               args.metadata_weight, args.top_n, args.min_similarity)
        logger.info(f"Total records stored: {stats['total_records']}")
        return stats
    except Exception as e:  # reference :399-401
        logger.error(f"Error: {e}", exc_info=True)
        sys.exit(1)


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    main()
