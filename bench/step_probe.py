"""A handful of Lloyd steps on synthetic blobs of any shape (env N, D, K, DTYPE=f64|f32, STEPS): step time and the
assignment kernel's time against the roofline of its shape.  Defaults: one GPU's shard of config C4 (12.5M x 128,
k = 1024, f64).  Small enough to sit under `ncu` (tools/profile.sh)."""
import os, sys, numpy as np
sys.path.insert(0, ".")
import smartcore_b200 as sc
n, d, k = int(os.environ.get("N", 12_500_000)), int(os.environ.get("D", 128)), int(os.environ.get("K", 1024))
f32 = os.environ.get("DTYPE", "f64") == "f32"
steps = int(os.environ.get("STEPS", 6))
ctx = sc.Context(0)
pk = ctx.device_peaks()
ds = ctx.generate_blobs(n, d, k, 20260101, dtype=np.float32 if f32 else np.float64)
cent0 = np.vstack([ds.download_rows(i * (n // k), 1) for i in range(k)]).astype(np.float64)
ds.lloyd_iterate(cent0, 2)
out = ds.lloyd_iterate(cent0, steps)
ms, ams = float(np.mean(out["ms"])), float(np.mean(out["assign_ms"]))
tf = 2.0 * n * k * d / (ams * 1e-3) / 1e12
gbs = n * (d * (4 if f32 else 8) + 4) / (ams * 1e-3) / 1e9
fp_peak = 1366.4 / 6 if f32 else max(pk["fp64_dmma_tflops"], pk["fp64_dfma_tflops"])
print("variant=%s sl=%s  %d x %d k=%d %s: step %.3f ms, assign %.3f ms = %.1f TFLOP/s = %.3f of %.1f | %.0f GB/s = %.3f of 6551" % (
    os.environ.get("SCKM_LIB_VARIANT", "default"), os.environ.get("SCKM_DMMA_SL", "12"), n, d, k, "f32" if f32 else "f64", ms, ams, tf,
    tf / fp_peak, fp_peak, gbs, gbs / 6551.4), flush=True)
