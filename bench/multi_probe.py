"""ONE process, every visible GPU behind one context (sckm_ctx_create_multi): phase times of sckm_kmeans_fit on a
pageable host matrix (env N, D, K, ITERS), with the library's kmeans++ trace (SCKM_TRACE=1) on stderr."""
import os, sys, time, numpy as np
sys.path.insert(0, ".")
import smartcore_b200 as sc
from smartcore_b200 import cabi, cluster
n, d, k = int(os.environ.get("N", 10_000_000)), int(os.environ.get("D", 64)), int(os.environ.get("K", 256))
iters = int(os.environ.get("ITERS", 10))
one = sc.Context(0)
ds = one.generate_blobs(n, d, k, 20260101)
x = np.empty((n, d))
for q in range(0, n, 1 << 20):
    m = min(1 << 20, n - q); x[q:q + m] = ds.download_rows(q, m)
ds.close()
first, u = cluster.kmeanspp_draws(42, n, k)
for label, ctx in (("1 device", one), ("all devices", sc.Context(devices="all"))):
    ctx.kmeans_fit(x[: 1 << 21], k, 2, first % (1 << 21), u)
    os.environ["SCKM_TRACE"] = "1"
    t = time.perf_counter(); r = ctx.kmeans_fit(x, k, iters, first, u); dt = time.perf_counter() - t
    os.environ.pop("SCKM_TRACE")
    ph = ctx.last_fit_times()
    print("%-12s devices %d: total %.3f s | upload %.3f kmeans++ + means %.3f lloyd %.3f (%d iters) download %.3f" % (
        label, ph["devices"], dt, ph["upload_s"], ph["kmeanspp_init_s"], ph["lloyd_s"], r["iters"], ph["download_s"]), flush=True)
