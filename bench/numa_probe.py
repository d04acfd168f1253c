"""Host topology seen by the process + H2D upload rate of a 5.12 GB pageable matrix with the staging threads bound to
the GPU's NUMA node and unbound (SCKM_INGEST_NO_NUMA=1), lane counts 4 / 8."""
import os, sys, time, subprocess, numpy as np
sys.path.insert(0, ".")
import smartcore_b200 as sc
from smartcore_b200 import cabi
print("cpu_count", os.cpu_count(), "allowed", len(os.sched_getaffinity(0)))
for cmd in ("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)'", "nvidia-smi topo -m | head -12", "free -g | head -2"):
    print(subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout)
n, d = 10_000_000, 64
x = np.empty((n, d)); x[:] = 1.5
gb = x.nbytes / 1e9
os.environ["SCKM_TRACE"] = "1"
for numa in (True, False):
    for threads in (8, 4, 16):
        if numa: os.environ.pop("SCKM_INGEST_NO_NUMA", None)
        else: os.environ["SCKM_INGEST_NO_NUMA"] = "1"
        os.environ["SCKM_INGEST_THREADS"] = str(threads)
        c = sc.Context(0)
        ts = []
        for _ in range(3):
            t = time.perf_counter(); ds = c.upload(x); ts.append(time.perf_counter() - t); ds.close()
        print("numa=%s threads=%d: first %.1f ms (%.1f GB/s), best %.1f ms (%.1f GB/s)" % (numa, threads, ts[0] * 1e3, gb / ts[0], min(ts) * 1e3, gb / min(ts)), flush=True)
        c.close()
