"""Throughput of sckm_knn (batched LinearKNNSearch::find, Euclidian, exact reference arithmetic):
query-row pairs per second against the FP64-ALU bound of 3*d non-fused operations per pair."""
import sys, time, numpy as np
sys.path.insert(0, ".")
import smartcore_b200 as sc
from smartcore_b200 import cabi
ctx = sc.Context(0)
for n, d, nq, k, dtype in ((1_000_000, 64, 1024, 8, np.float64), (1_000_000, 16, 4096, 5, np.float32), (10_000_000, 64, 64, 8, np.float64)):
    ds = ctx.generate_blobs(n, d, 16, 7, dtype=dtype)
    q = cabi.blobs_host(123, nq, d, 16, 7, dtype=dtype) + 0.25
    ds.knn(q[:8], k)
    t = time.perf_counter(); idx, dist = ds.knn(q, k); dt = time.perf_counter() - t
    pairs = n * nq
    print("knn n=%d d=%d %s nq=%d k=%d: %.1f ms, %.3g pairs/s, %.2f TFLOP-equivalent/s (3d ops per pair; FP64 non-fused peak ~18 T op/s)"
          % (n, d, np.dtype(dtype).name, nq, k, dt * 1e3, pairs / dt, 3 * d * pairs / dt / 1e12))
    ds.close()
