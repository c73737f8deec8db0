"""B200-native (sm_100a) implementation of the hybrid all-pairs similarity -> top-K path of
tomboone/tvbingefriend-recommendation-service, behind the reference's own Python API:

    from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer
    from tvbingefriend_recommendation_service_b200.services.content_based_service import (
        ContentBasedRecommendationService)

All arithmetic runs in ``libtvbf.so`` (hand-written CUDA: TMA-fed tcgen05 GEMM with a fused
scoring / candidate-selection epilogue, fp64 rescoring, exact repair) through the C ABI declared
in ``include/tvbf.h``.  There is no CPU fallback.
"""

__version__ = "0.1.0"
