"""Config C2 timeline probe (diagnostic build: tools/build_variant.sh trace "-DSCKM_STREAM_TRACE", SCKM_LIB_VARIANT=trace):
where the ~40 us of one assign_stream_kernel launch go -- per-CTA globaltimer stamps at entry, after the dependency
wait, at the start and end of the batch loop, after the CTA barrier and at exit, relative to the earliest entry."""
import os, sys, ctypes, numpy as np
sys.path.insert(0, ".")
os.environ.setdefault("SCKM_LIB_VARIANT", "trace")
import smartcore_b200 as sc
from smartcore_b200 import cabi, cluster
n, d, k = int(os.environ.get("N", 1_000_000)), 16, 8
ctx = sc.Context(0)
ds = ctx.generate_blobs(n, d, k, 20260101)
first, u = cluster.kmeanspp_draws(42, n, k)
ds.kmeanspp(k, first, u)
cent0, _ = ds.init_centroids(k)
lib = ctypes.CDLL(cabi.LIB_PATH)
for rep in range(3):
    out = ds.lloyd_iterate(cent0, 6)
    ctas = 296
    buf = (ctypes.c_ulonglong * (ctas * 8))()
    rc = lib.sckm_debug_stream_trace(buf, ctas)
    t = np.frombuffer(buf, dtype=np.uint64).reshape(ctas, 8).astype(np.int64)
    t0 = t[:, 0].min()
    names = ["entry", "after pdl wait", "loop start", "loop end", "exit", "after barrier"]
    print("rep %d rc=%d  assign_ms (events) last steps: %s" % (rep, rc, " ".join("%.1f" % (v * 1e3) for v in out["assign_ms"][-3:])))
    for i in (0, 1, 2, 3, 5, 4):
        c = (t[:, i] - t0) / 1e3
        print("  %-15s min %6.2f  median %6.2f  p90 %6.2f  max %6.2f us" % (names[i], c.min(), np.median(c), np.percentile(c, 90), c.max()))
    dur = (t[:, 3] - t[:, 2]) / 1e3
    print("  loop duration   min %6.2f  median %6.2f  max %6.2f us" % (dur.min(), np.median(dur), dur.max()))
ds.close(); ctx.close()
