import os, sys, time
sys.path.insert(0, os.getcwd())
os.environ.setdefault("NCCL_DEBUG", "WARN")
import numpy as np, torch, torch.distributed as tdist
import smartcore_b200 as sc
from smartcore_b200 import cluster, dist as scd
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = sc.Context(local); scd.join_comm(ctx)
n, d, k = 10_000_000, 64, 64
ds = ctx.generate_blobs(n, d, 256, 20260101, row_offset=rank * n, n_global=n * world)
first, u = cluster.kmeanspp_draws(42, n * world, k)
for rep in range(3):
    tdist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    ds.kmeanspp(k, first, u)
    tdist.barrier(); torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0: print("world", world, "kmeans++ k=%d: %.1f ms total, %.3f ms/pass" % (k, 1e3 * (t1 - t0), 1e3 * (t1 - t0) / k), os.environ.get("SCKM_DBG", ""), flush=True)
# raw torch NCCL small collectives for comparison
x = torch.zeros(65, dtype=torch.float64, device="cuda")
tdist.all_reduce(x); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): tdist.all_reduce(x)
torch.cuda.synchronize(); t1 = time.perf_counter()
if rank == 0: print("torch all_reduce(65 f64): %.1f us each" % (1e6 * (t1 - t0) / 200), flush=True)
ctx.close(); tdist.destroy_process_group()
