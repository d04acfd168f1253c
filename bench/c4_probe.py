"""One GPU's shard of config C4 (12.5M x 128, k = 1024, f64): step time of the streamed-centroid tile kernel."""
import os, sys, numpy as np
sys.path.insert(0, ".")
import smartcore_b200 as sc
n, d, k = int(os.environ.get("N", 12_500_000)), int(os.environ.get("D", 128)), int(os.environ.get("K", 1024))
ctx = sc.Context(0)
pk = ctx.device_peaks()
ds = ctx.generate_blobs(n, d, k, 20260101)
cent0 = np.vstack([ds.download_rows(i * (n // k), 1) for i in range(k)]).astype(np.float64)
ds.lloyd_iterate(cent0, 2)
out = ds.lloyd_iterate(cent0, 6)
ms, ams = float(np.mean(out["ms"])), float(np.mean(out["assign_ms"]))
print("variant=%s sl=%s  %d x %d k=%d: step %.2f ms, assign %.2f ms = %.1f TFLOP/s = %.3f of %.1f" % (
    os.environ.get("SCKM_LIB_VARIANT", "default"), os.environ.get("SCKM_DMMA_SL", "12"), n, d, k, ms, ams, 2.0 * n * k * d / (ams * 1e-3) / 1e12,
    2.0 * n * k * d / (ams * 1e-3) / 1e12 / max(pk["fp64_dmma_tflops"], pk["fp64_dfma_tflops"]), max(pk["fp64_dmma_tflops"], pk["fp64_dfma_tflops"])), flush=True)
