"""Host -> device ingest rates of sckm_dataset_upload: pageable numpy memory (threaded pinned ring vs the plain
single-copy path, SCKM_INGEST_DIRECT=1), pinned memory, column-major; and sckm_predict end to end."""
import os, sys, time, numpy as np
sys.path.insert(0, ".")
import torch
import smartcore_b200 as sc
from smartcore_b200 import cabi

ctx = sc.Context(0)
n, d, k = 4_000_000, 64, 256            # 2.05 GB of f64
x = cabi.blobs_host(0, 1_000_000, d, k, 3)
x = np.ascontiguousarray(np.tile(x, (n // x.shape[0], 1)))
gb = x.nbytes / 1e9
print("host cores:", os.cpu_count())

def timed(label, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); r = fn(); dt = time.perf_counter() - t
        best = min(best, dt)
        if hasattr(r, "close"): r.close()
    print("%-46s %7.1f ms  %6.1f GB/s" % (label, best * 1e3, gb / best))
    return best

timed("upload pageable, staged ring", lambda: ctx.upload(x))
for t in (2, 4, 8):
    os.environ["SCKM_INGEST_THREADS"] = str(t)
    c2 = sc.Context(0)
    timed("upload pageable, staged ring, %d threads" % t, lambda: c2.upload(x))
    c2.close()
del os.environ["SCKM_INGEST_THREADS"]
os.environ["SCKM_INGEST_DIRECT"] = "1"
timed("upload pageable, plain cudaMemcpy", lambda: ctx.upload(x))
del os.environ["SCKM_INGEST_DIRECT"]
xp = torch.empty((n, d), dtype=torch.float64).pin_memory()
xp.numpy()[:] = x
timed("upload pinned", lambda: ctx.upload(xp.numpy()))
xt = np.ascontiguousarray(x.T)                       # the column-major image, built outside the timed call
timed("upload pageable column-major (+transpose)", lambda: ctx.upload_colmajor_image(xt, n, d))
cent = x[: 1_000_000 : 1_000_000 // k][:k].copy() + 0.01   # distinct rows (x repeats every 1M rows)
timed("predict pageable (chunked, overlapped)", lambda: ctx.predict(x, cent))
os.environ["SCKM_INGEST_DIRECT"] = "1"
timed("predict pageable, plain cudaMemcpy", lambda: ctx.predict(x, cent))
del os.environ["SCKM_INGEST_DIRECT"]
timed("predict pinned", lambda: ctx.predict(xp.numpy(), cent))
