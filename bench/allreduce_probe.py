"""Per-step cost of the sum over the ranks: peer exchange inside the finalize kernel (csrc/sckm_peer.cu) vs ncclAllReduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 bench/allreduce_probe.py

Strong-scaling shapes (the rows of a BASELINE.json config split over the N ranks), where the fixed per-step cost shows:
every shape runs the same K timed steps twice per path, alternating, device-timed as a whole loop (max over ranks)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as tdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smartcore_b200 as sc  # noqa: E402
from smartcore_b200 import dist as scd  # noqa: E402

SHAPES = [("C3 10M x 64 k=256 f64", 10_000_000, 64, 256, np.float64, 20),
          ("C2 1M x 16 k=8 f64", 1_000_000, 16, 8, np.float64, 100),
          ("C5 50M x 32 k=4096 f32", 50_000_000, 32, 4096, np.float32, 6),
          ("8M x 16 k=8 f64", 8_000_000, 16, 8, np.float64, 100)]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sc.Context(local)
    scd.join_comm(ctx)
    only = os.environ.get("SHAPES")
    rows = []
    for name, n, d, k, dt, steps in SHAPES:
        if only and name.split()[0] not in only.split(","):
            continue
        lo, hi = scd.shard_range(n, world, rank)
        ds = ctx.generate_blobs(hi - lo, d, k, 20240607, dtype=dt, row_offset=lo, n_global=n)
        step = max(1, (hi - lo) // k)
        mine = np.vstack([ds.download_rows(i * step, 1) for i in range(k)]).astype(np.float64)
        cent0 = torch.from_numpy(mine).cuda()
        tdist.broadcast(cent0, 0)                               # every rank starts from rank 0's rows
        cent0 = cent0.cpu().numpy()
        res = {}
        for rep in range(2):
            for mode in ("peer", "nccl"):
                if mode == "nccl":
                    os.environ["SCKM_PEER_ALLREDUCE"] = "0"
                else:
                    os.environ.pop("SCKM_PEER_ALLREDUCE", None)
                ds.lloyd_iterate(cent0, 3, per_step_events=False)
                tdist.barrier(); torch.cuda.synchronize()
                out = ds.lloyd_iterate(cent0, steps, want_inertia=True, per_step_events=False)
                ms = scd.max_over_ranks(float(out["ms"].sum())) / steps
                res.setdefault(mode, []).append(ms)
                res[mode + "_path"] = ctx.allreduce_path()
                res[mode + "_out"] = out
        os.environ.pop("SCKM_PEER_ALLREDUCE", None)
        same = bool(np.allclose(res["peer_out"]["centroids"], res["nccl_out"]["centroids"], rtol=1e-12, atol=0))
        if rank == 0:
            row = {"shape": name, "n_gpus": world, "rows_per_gpu": hi - lo, "payload_bytes": 8 * (k * d + k + 1),
                   "ms_per_step_peer": min(res["peer"]), "ms_per_step_nccl": min(res["nccl"]), "paths": [res["peer_path"], res["nccl_path"]],
                   "us_saved_per_step": 1e3 * (min(res["nccl"]) - min(res["peer"])), "centroids_agree_1e-12": same,
                   "all_runs_ms": {m: res[m] for m in ("peer", "nccl")}}
            rows.append(row)
            print("ALLREDUCE_PROBE " + json.dumps(row), flush=True)
        ds.close()
    ctx.close()
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
