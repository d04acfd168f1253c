import os, sys, time, numpy as np
os.environ["SCKM_TRACE"] = "1"
sys.path.insert(0, ".")
import smartcore_b200 as sc
from smartcore_b200 import cluster
n, d, k = 10_000_000, 64, 256
ctx = sc.Context(0)
ds = ctx.generate_blobs(n, d, k, 20260101)
first, u = cluster.kmeanspp_draws(42, n, k)
for i in range(40):
    t = time.perf_counter(); ds.kmeanspp(k, first, u); dt = (time.perf_counter() - t) * 1e3
    sys.stderr.write("  -> call %d wall %.1f ms\n" % (i, dt))
