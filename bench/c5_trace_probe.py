"""Config C5 timeline probe (diagnostic build: tools/build_variant.sh h5trace "-DSCKM_TC5H_TRACE", SCKM_LIB_VARIANT=h5trace):
where the MMA warp and one epilogue warp of every CTA of assign_tc5h_kernel spend their clocks."""
import os, sys, ctypes, numpy as np
sys.path.insert(0, ".")
os.environ.setdefault("SCKM_LIB_VARIANT", "h5trace")
import smartcore_b200 as sc
from smartcore_b200 import cabi, cluster
n, d, k = int(os.environ.get("N", 10_000_000)), 32, 4096
ctx = sc.Context(0)
ds = ctx.generate_blobs(n, d, k, 20260101, dtype=np.float32)
first, u = cluster.kmeanspp_draws(42, n, k)
ds.kmeanspp(k, first, u)
cent0, _ = ds.init_centroids(k)
lib = ctypes.CDLL(cabi.LIB_PATH)
ctas = 148
buf = (ctypes.c_longlong * (ctas * 8))()
ds.lloyd_iterate(cent0, 3)
lib.sckm_debug_tc5h_trace(buf, ctas, 1)
steps = 6
out = ds.lloyd_iterate(cent0, steps)
lib.sckm_debug_tc5h_trace(buf, ctas, 1)
t = np.frombuffer(buf, dtype=np.int64).reshape(ctas, 8).astype(np.float64) / steps
nblk = (n / 256 / 148) * (k / 128)
print("assign ms per step:", " ".join("%.2f" % v for v in out["assign_ms"]), " block iterations per CTA per step: %.0f" % nblk)
names = ["MMA wait x_ready", "MMA wait c_full", "MMA wait t_empty", "MMA warp total", "epi wait t_full", "epi split", "epi exact part", "epi warp total"]
for i, nm in enumerate(names):
    c = t[:, i]
    print("  %-18s mean %12.0f clk per step  = %7.0f clk per block iteration   (min %.0f max %.0f)" % (nm, c.mean(), c.mean() / nblk, c.min() / nblk, c.max() / nblk))
ds.close(); ctx.close()
