"""Does a background `nvidia-smi -lms 100` (the bench's clock sampler) perturb launch-bound phases?  Wall time of
kmeans++ (1500 launches, ~75 ms) repeated 40 times with and without the sampler, and with an in-process NVML poll."""
import subprocess, sys, threading, time, numpy as np
sys.path.insert(0, ".")
import smartcore_b200 as sc
from smartcore_b200 import cluster
n, d, k = 10_000_000, 64, 256
ctx = sc.Context(0)
ds = ctx.generate_blobs(n, d, k, 20260101)
first, u = cluster.kmeanspp_draws(42, n, k)
ds.kmeanspp(k, first, u)

def series(label, reps=40):
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); ds.kmeanspp(k, first, u); ts.append((time.perf_counter() - t) * 1e3)
    ts = np.array(ts)
    print("%-28s median %.1f ms  p90 %.1f  max %.1f  (>150 ms: %d of %d)" % (label, np.median(ts), np.percentile(ts, 90), ts.max(), int((ts > 150).sum()), reps))

series("no sampler")
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                      "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.DEVNULL)
time.sleep(0.5)
series("nvidia-smi -lms 100")
p.terminate(); p.wait()
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
stop = False
def poll():
    while not stop:
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        time.sleep(0.1)
th = threading.Thread(target=poll, daemon=True); th.start()
series("in-process NVML @100 ms")
stop = True
series("no sampler (again)")
