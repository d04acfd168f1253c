"""Config C2 (1M x 16, k = 8, f64) tuning probe: per-step and assignment-kernel times of the streaming path under the
run-time toggles (one-launch reduce + finalize, bulk-copy ring) and library variants (SCKM_LIB_VARIANT), and the wall
time of a whole lloyd_fit with the device-side stop rule at different enqueue batch sizes."""
import os, sys, time, numpy as np
sys.path.insert(0, ".")
import smartcore_b200 as sc
from smartcore_b200 import cabi, cluster

n, d, k = int(os.environ.get("N", 1_000_000)), 16, 8
ctx = sc.Context(0)
ds = ctx.generate_blobs(n, d, k, 20260101)
first, u = cluster.kmeanspp_draws(42, n, k)
ds.kmeanspp(k, first, u)
cent0, _ = ds.init_centroids(k)
hbm = n * (d * 8 + 4)

def steps(label, env):
    for key in ("SCKM_NO_STEP_SMALL", "SCKM_STREAM_TMA"):
        os.environ.pop(key, None)
    os.environ.update(env)
    ds.lloyd_iterate(cent0, 5)
    best = None
    for _ in range(3):
        out = ds.lloyd_iterate(cent0, 40)
        ms, ams = float(np.mean(out["ms"][2:])), float(np.mean(out["assign_ms"][2:]))
        if best is None or ms < best[0]:
            best = (ms, ams, out)
    ms, ams, out = best
    print("[%s] %-40s step %7.2f us  assign %7.2f us  -> %5.3f of 6551 GB/s per step, %5.3f kernel" %
          (os.environ.get("SCKM_LIB_VARIANT", "default"), label, ms * 1e3, ams * 1e3, hbm / (ms * 1e-3) / 1e9 / 6551.4, hbm / (ams * 1e-3) / 1e9 / 6551.4), flush=True)
    return out

a = steps("default", {})
f = steps("TMA ring", {"SCKM_STREAM_TMA": "1"})

for key in ("SCKM_NO_STEP_SMALL", "SCKM_STREAM_TMA"):
    os.environ.pop(key, None)
for batch in ("1", "4", "8", ""):
    if batch: os.environ["SCKM_LLOYD_BATCH"] = batch
    else: os.environ.pop("SCKM_LLOYD_BATCH", None)
    ds.lloyd_fit(cent0, 100)
    t = time.perf_counter(); r = ds.lloyd_fit(cent0, 100); dt = time.perf_counter() - t
    print("lloyd_fit batch=%-7s iters %3d  wall %7.2f ms  %6.1f us/iter" % (batch or "default", r["iters"], dt * 1e3, dt * 1e6 / r["iters"]), flush=True)
ds.close(); ctx.close()
