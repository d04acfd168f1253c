"""Config C5 shape (f32, d = 32, k = 4096) tuning probe: step / assignment times of the tcgen05 kernel with and without
the primed chunk skipping (SCKM_TC5_NOPRIME), over the first steps of a fit (priming pays more as fewer rows move)."""
import os, sys, numpy as np
sys.path.insert(0, ".")
import smartcore_b200 as sc
from smartcore_b200 import cluster
n, d, k = int(os.environ.get("N", 10_000_000)), int(os.environ.get("D", 32)), int(os.environ.get("K", 4096))
ctx = sc.Context(0)
ds = ctx.generate_blobs(n, d, k, 20260101, dtype=np.float32)
first, u = cluster.kmeanspp_draws(42, n, k)
ds.kmeanspp(k, first, u)
cent0, _ = ds.init_centroids(k)
roof = 1366.4 / 2 / 3
for label, env in (("primed", {}), ("unprimed", {"SCKM_TC5_NOPRIME": "1"})):
    os.environ.pop("SCKM_TC5_NOPRIME", None); os.environ.update(env)
    ds.lloyd_iterate(cent0, 2)
    out = ds.lloyd_iterate(cent0, 12, want_inertia=True)
    a = out["assign_ms"]
    tf = lambda ms: 2.0 * n * k * d / (ms * 1e-3) / 1e12
    print("%-9s n=%d: assign ms per step %s" % (label, n, " ".join("%.2f" % v for v in a)), flush=True)
    print("%-9s mean step %.2f ms, assign %.2f ms = %.1f TFLOP/s = %.3f of the 3xTF32 roof (%.1f); last 4 steps: %.2f ms = %.3f" % (
        label, float(np.mean(out["ms"])), float(np.mean(a)), tf(float(np.mean(a))), tf(float(np.mean(a))) / roof, roof,
        float(np.mean(a[-4:])), tf(float(np.mean(a[-4:]))) / roof), flush=True)
    ref = out if label == "primed" else ref
    if label == "unprimed":
        print("same sizes:", np.array_equal(out["size"], ref["size"]), "centroids bit-equal:", np.array_equal(out["centroids"], ref["centroids"]),
              "inertia equal:", np.array_equal(out["inertia"], ref["inertia"]))
ds.close(); ctx.close()
