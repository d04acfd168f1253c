"""smartcore_b200 -- B200-native (sm_100a) k-means hot path of smartcore behind its own API.

Layout
  csrc/   hand-written CUDA kernels + the C ABI (include/smartcore_kmeans_cuda.h)
  host/   C++ mirror of smartcore's KMeans / KMeansParameters / DenseMatrix / Failed / rand_custom
  lib/    built shared objects (git-ignored; built by __graft_entry__.build())
  cabi.py     ctypes binding of the C ABI (Context, Dataset)
  cluster.py  Python face of the host mirror: KMeans.fit / predict, KMeansParameters, DenseMatrix
  metrics.py  cluster-quality scores of the reference (contingency on the GPU): HCVScore, entropy, mutual_info_score
  neighbour.py  LinearKNNSearch (Euclidian rows) with the batched search on the GPU
  dist.py     one-process-per-GPU plumbing over torch.distributed (row sharding, NCCL id exchange)

There is no CPU fallback: importing cabi without the built CUDA library raises.
"""
from .cabi import Context, Dataset, SckmError, F32, F64  # noqa: F401
from .cluster import KMeans, KMeansParameters, KMeansSearchParameters, DenseMatrix, Failed, set_std_rand, get_std_rand  # noqa: F401
from .metrics import HCVScore, contingency_matrix, entropy, mutual_info_score  # noqa: F401
from .neighbour import LinearKNNSearch  # noqa: F401
