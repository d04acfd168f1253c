"""Worker for tests/test_gpu_multirank.py (launched by torchrun, one rank per GPU): sharded kmeans++ + Lloyd
through the C ABI with the library's NCCL communicator; rank 0 compares with the single-rank result."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as tdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smartcore_b200 as sc  # noqa: E402
from smartcore_b200 import cabi, cluster, dist as scd  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sc.Context(local)
    scd.join_comm(ctx)
    n, d, k = 50_000, 32, 24
    x = cabi.blobs_host(0, n, d, k, 99)
    lo, hi = scd.shard_range(n, world, rank)
    first, u = cluster.kmeanspp_draws(7, n, k)
    ds = ctx.upload(x[lo:hi], row_offset=lo, n_global=n)
    seeds = ds.kmeanspp(k, first, u)
    cent, size = ds.init_centroids(k)
    fit = ds.lloyd_fit(cent, 50)
    truth = (np.arange(n) % k).astype(np.uint32)           # blobs_host: row i belongs to centre i % k
    table = ds.contingency(truth[lo:hi], k, k)             # summed over ranks inside the library (NCCL, u64)
    labels = torch.zeros(n, dtype=torch.int64, device="cuda")
    labels[lo:hi] = torch.from_numpy(ds.labels().astype(np.int64)).cuda()
    tdist.all_reduce(labels)
    if rank == 0:
        ctx1 = sc.Context(local)
        ds1 = ctx1.upload(x)
        seeds1 = ds1.kmeanspp(k, first, u)
        cent1, size1 = ds1.init_centroids(k)
        fit1 = ds1.lloyd_fit(cent1, 50)
        ok = (seeds.tolist() == seeds1.tolist() and size.tolist() == size1.tolist()
              and np.allclose(cent, cent1, rtol=1e-12) and fit["iters"] == fit1["iters"]
              and np.allclose(fit["centroids"], fit1["centroids"], rtol=1e-9)
              and abs(fit["distortion"] - fit1["distortion"]) <= 1e-9 * fit1["distortion"]
              and fit["size"].tolist() == fit1["size"].tolist()
              and np.array_equal(labels.cpu().numpy(), ds1.labels().astype(np.int64))
              and np.array_equal(table, ds1.contingency(truth, k, k)) and int(table.sum()) == n)
        print("MULTIRANK_RESULT " + json.dumps({"ok": bool(ok), "iters": int(fit["iters"]), "world": world}))
    ds.close(); ctx.close()
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
