"""Times sckm_predict (host buffers in, labels out) on a large shape: DMMA ranking vs the direct form."""
import time, sys, numpy as np
sys.path.insert(0, ".")
import smartcore_b200 as sc
from smartcore_b200 import cabi
ctx = sc.Context(0)
n, d, k = 2_000_000, 64, 256
x = cabi.blobs_host(0, n, d, k, 3)
cent = x[:: n // k][:k].copy() + 0.01
for name, mode in (("auto(dmma)", cabi.ASSIGN_AUTO), ("direct", cabi.ASSIGN_DIRECT)):
    ctx.set_assign_kernel(mode)
    ctx.predict(x[:1000], cent)
    t = time.perf_counter(); y = ctx.predict(x, cent); dt = time.perf_counter() - t
    print(name, "predict %dx%d k=%d: %.1f ms (%.3g rows/s) checksum %d" % (n, d, k, dt * 1e3, n / dt, int(y.sum())))
