"""recsys_lite_b200 -- B200-native (sm_100a) drop-in for the recsys-lite hot path.

Re-exports the names of /root/reference/src/recsys_lite/__init__.py:21-52 that lie on the path named in
BASELINE.json: ``RecommenderSystem`` (algorithm "als" | "svd"), ``svd_reconstruct``, ``top_n``,
``compute_rmse``, ``compute_mae``.  The arithmetic runs in librecsys_b200.so (hand-written CUDA, C ABI in
include/recsys_b200.h) -- there is no CPU fallback.
"""
from .api import RecommenderSystem, compute_mae, compute_rmse, svd_reconstruct, top_n
from .benchmark import benchmark_algorithm, quick_benchmark
from .metrics import ndcg_at_k, precision_at_k, ranking_metrics, recall_at_k
from .pipeline import RecsysPipeline

__version__ = "0.1.0"

__all__ = ["RecommenderSystem", "svd_reconstruct", "top_n", "compute_rmse", "compute_mae", "benchmark_algorithm",
           "quick_benchmark", "RecsysPipeline", "precision_at_k", "recall_at_k", "ndcg_at_k", "ranking_metrics", "__version__"]
