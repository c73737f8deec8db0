"""Drop-in for scripts/populate_recommendations.py (reference :45-101): the older driver that
syncs the show metadata and then calls ``service.compute_and_store_all_similarities()``.

Metadata persistence is storage (out of scope): it is forwarded to the sink when that offers
``bulk_store_shows`` and skipped otherwise.  The similarity step is the service's GPU path."""

from __future__ import annotations

import logging
import sys
from pathlib import Path

from ..services.content_based_service import ContentBasedRecommendationService

logger = logging.getLogger(__name__)


def main(processed_data_dir: Path | None = None, sink=None, service: ContentBasedRecommendationService | None = None):
    logger.info("=" * 70)
    logger.info("POPULATING RECOMMENDATION DATABASE")
    logger.info("=" * 70)
    try:
        if service is None:
            service = ContentBasedRecommendationService(processed_data_dir=processed_data_dir, sink=sink)
        processed_dir = service.processed_data_dir
        import pandas as pd

        shows_df = pd.read_csv(processed_dir / "shows_metadata.csv")          # reference :56
        logger.info(f"Loaded {len(shows_df)} shows from CSV")
        if hasattr(service.sink, "bulk_store_shows"):                         # reference :58-66
            cleaned = shows_df.astype(object).where(pd.notnull(shows_df), None)
            metadata_count = service.sync_metadata_to_db(cleaned.to_dict("records"))
            logger.info(f"✓ Synced {metadata_count} shows")
        else:
            logger.info("⊘ No metadata store configured (sink has no bulk_store_shows): skipping metadata sync")
        stats = service.compute_and_store_all_similarities()                  # reference :70
        logger.info(f"Total similarity records: {stats['total_records']}")
        logger.info(f"Unique shows with recommendations: {stats['unique_shows']}")
        logger.info(f"Average similarities per show: {stats['avg_similarities_per_show']:.1f}")
        for show_id in shows_df["id"].head(3).tolist():                       # reference :84-98
            recs = service.get_recommendations_from_db(show_id=show_id, n=5)
            logger.info(f"Recommendations for show {show_id}: "
                        f"{[(r.get('similar_show_id', r.get('show_id')), round(r['similarity_score'], 3)) for r in recs]}")
        logger.info("\n✓ Database population complete!")
        return stats
    except Exception as e:
        logger.error(f"Error during population: {e}", exc_info=True)
        sys.exit(1)


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    main()
