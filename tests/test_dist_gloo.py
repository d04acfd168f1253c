"""CPU, world_size=2 over gloo: the multi-rank plumbing (row sharding, id broadcast, max-over-ranks)
and the sharded algorithm itself -- partial [sums|counts|inertia] + one all-reduce per Lloyd step, and
the kmeans++ owner-selection protocol -- emulated with the oracle per shard and compared with the
single-process oracle.  (The CUDA library runs the same protocol with NCCL)"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py as O
    from smartcore_b200 import dist as scd
    try:
        # plumbing
        payload = bytes(range(128)) if rank == 0 else None
        got = scd.broadcast_bytes(payload, 0, 128)
        assert got == bytes(range(128))
        assert scd.max_over_ranks(1.0 + rank) == float(world)

        # sharded Lloyd step == single-process step
        rng = np.random.default_rng(5)
        n, d, k = 5000, 6, 5
        x = rng.normal(size=(n, d)) + 4.0 * rng.integers(0, k, size=(n, 1))
        cent = x[rng.choice(n, k, replace=False)]
        lo, hi = scd.shard_range(n, world, rank)
        dist_l, sums_l, counts_l, mem_l = O.brute_clustering(x[lo:hi], cent)
        packed = torch.from_numpy(np.concatenate([sums_l.reshape(-1), counts_l.astype(np.float64), [dist_l]]))
        tdist.all_reduce(packed)
        packed = packed.numpy()
        dist_g, sums_g, counts_g, mem_g = O.brute_clustering(x, cent)
        np.testing.assert_allclose(packed[:k * d].reshape(k, d), sums_g, rtol=1e-12)
        assert packed[k * d:k * d + k].astype(np.int64).tolist() == counts_g.tolist()
        assert abs(packed[-1] - dist_g) <= 1e-12 * dist_g
        assert np.array_equal(mem_l, mem_g[lo:hi])

        # kmeans++ owner selection: all-gather rank totals -> cutoff -> owner -> residual scan
        y_g, idx_g, dd_g = O.kmeanspp(x, 2, seed=11)          # one D^2 pass against seed row idx_g[0]
        first = int(idx_g[0])
        dd_l = ((x[lo:hi] - x[first]) ** 2)
        dd_l = np.array([O.squared_distance(r, x[first]) for r in x[lo:hi]])
        tot = torch.zeros(world, dtype=torch.float64); tot[rank] = float(np.add.reduce(dd_l))
        tdist.all_reduce(tot)
        r = O.Rng(11); r.gen_range(n); u = r.gen_f64()
        cutoff = u * float(tot.sum())
        run, owner = 0.0, world - 1
        for q in range(world):
            if run + float(tot[q]) >= cutoff:
                owner = q; break
            run += float(tot[q])
        pick = torch.zeros(1, dtype=torch.int64)
        if owner == rank:
            cost, index = run, 0
            while index < hi - lo:
                cost += dd_l[index]
                if cost >= cutoff:
                    break
                index += 1
            pick[0] = lo + min(index, hi - lo - 1)
        tdist.all_reduce(pick)                                   # zeros elsewhere: broadcast from unknown root
        assert int(pick[0]) == int(idx_g[1]), (int(pick[0]), int(idx_g[1]))
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, "FAIL %s\n%s" % (e, traceback.format_exc())))
    finally:
        tdist.destroy_process_group()


def test_two_rank_gloo_protocol():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
