"""kmeans++ alone on a device-generated C3-shaped dataset (10M x 64, k = 256): wall time of the k passes."""
import sys, time, numpy as np
sys.path.insert(0, ".")
import smartcore_b200 as sc
from smartcore_b200 import cluster
n, d, k = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (10_000_000, 64, 256)))
ctx = sc.Context(0)
ds = ctx.generate_blobs(n, d, k, 20260101)
first, u = cluster.kmeanspp_draws(42, n, k)
for rep in range(2):
    t = time.perf_counter(); seeds = ds.kmeanspp(k, first, u); dt = time.perf_counter() - t
    print("kmeans++ %dx%d k=%d: %.1f ms (%.3f ms/pass) seeds[:4]=%s" % (n, d, k, dt * 1e3, dt * 1e3 / k, seeds[:4].tolist()))
