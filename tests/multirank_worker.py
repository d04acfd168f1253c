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
    path = ctx.allreduce_path()                            # 'peer': the sum over the ranks ran inside the finalize kernel
    # the same loop with the all-reduce through NCCL; every rank's result of the peer path gathered for a bitwise comparison
    os.environ["SCKM_PEER_ALLREDUCE"] = "0"
    fit_nccl = ds.lloyd_fit(cent, 50)
    path_nccl = ctx.allreduce_path()
    del os.environ["SCKM_PEER_ALLREDUCE"]
    mine = torch.from_numpy(np.concatenate([fit["centroids"].ravel(), [fit["distortion"]]])).cuda()
    everyone = [torch.empty_like(mine) for _ in range(world)]
    tdist.all_gather(everyone, mine)
    ranks_agree = all(torch.equal(e.view(torch.int64), mine.view(torch.int64)) for e in everyone)
    fit = ds.lloyd_fit(cent, 50)                           # leaves the peer path's labels on the dataset
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
        # NCCL vs peer: same iteration count; two ranks add the same two numbers (bit-equal), more ranks may differ in order
        ok_paths = (path_nccl == "nccl" and path in ("peer", "nccl") and ranks_agree and fit_nccl["iters"] == fit["iters"]
                    and fit_nccl["size"].tolist() == fit["size"].tolist()
                    and np.allclose(fit_nccl["centroids"], fit["centroids"], rtol=1e-12, atol=0)
                    and (world != 2 or np.array_equal(fit_nccl["centroids"], fit["centroids"])))
        if os.environ.get("SCKM_EXPECT_PEER", "1") == "1" and torch.cuda.can_device_access_peer(0, 1):
            ok_paths = ok_paths and path == "peer"
        print("MULTIRANK_RESULT " + json.dumps({"ok": bool(ok and ok_paths), "fit_ok": bool(ok), "paths_ok": bool(ok_paths), "path": path,
                                                "ranks_agree": bool(ranks_agree), "iters": int(fit["iters"]), "world": world}))
    ds.close(); ctx.close()
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
