"""Small end-to-end exercise of every CUDA kernel, meant to run under compute-sanitizer on a B200:

    compute-sanitizer --tool memcheck  python tests/sanitizer_smoke.py
    compute-sanitizer --tool racecheck python tests/sanitizer_smoke.py

Shapes are tiny (the tools slow kernels down 10-100x); results are still checked against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smartcore_b200 as sc  # noqa: E402
from smartcore_b200 import cabi, cluster  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def main():
    ctx = sc.Context(0)
    rng = np.random.default_rng(0)
    cases = [(1500, 16, 8, np.float64, cabi.ASSIGN_STREAM), (1111, 7, 5, np.float32, cabi.ASSIGN_STREAM),
             (2000, 64, 48, np.float64, cabi.ASSIGN_DMMA), (1777, 20, 33, np.float32, cabi.ASSIGN_DMMA),
             (1200, 128, 300, np.float64, cabi.ASSIGN_DMMA),          # streamed centroid blocks
             (4096, 32, 64, np.float32, cabi.ASSIGN_TC5),              # tcgen05 / TMA / TMEM kernel (3xFP16 form, d <= 32)
             (3000, 12, 150, np.float32, cabi.ASSIGN_TC5),             # ... one K-step per product, padded centroid block
             (3000, 32, 130, np.float32, "tc5_ew16"),                  # ... two epilogue threads per row
             (3000, 32, 130, np.float32, "tc5_nofold"),                # ... norm term in the epilogue
             (3000, 32, 130, np.float32, "tc5_tf32"),                  # 3xTF32 form (the only one for 32 < d <= 64)
             (2500, 48, 70, np.float32, cabi.ASSIGN_TC5),
             (900, 9, 4, np.float64, cabi.ASSIGN_DIRECT), (700, 40, 6, np.float64, cabi.ASSIGN_DIRECT),
             (1300, 64, 5, np.float64, cabi.ASSIGN_AUTO),              # k < 16, d > 32: tile kernel, partly padded sub-block
             (1500, 64, 40, np.float64, "center"), (1100, 128, 200, np.float64, "center"),   # centred instantiations
             (1400, 12, 9, np.float64, "center")]                      # streaming kernel on offset data
    for n, d, k, dt, kern in cases:
        x = (rng.normal(size=(n, d)) + 3.0 * rng.integers(0, k, size=(n, 1))).astype(dt)
        if kern == "center":                                           # data far from the origin: launch_cnorm centres
            x = x + 1e4
            kern = cabi.ASSIGN_AUTO
        for key in ("SCKM_TC5H_EW16", "SCKM_TC5H_NOFOLD", "SCKM_TC5_TF32"):
            os.environ.pop(key, None)
        if isinstance(kern, str) and kern.startswith("tc5_"):
            os.environ[{"tc5_ew16": "SCKM_TC5H_EW16", "tc5_nofold": "SCKM_TC5H_NOFOLD", "tc5_tf32": "SCKM_TC5_TF32"}[kern]] = "1"
            kern = cabi.ASSIGN_TC5
        first, u = cluster.kmeanspp_draws(3, n, k)
        ds = ctx.upload(x)
        seeds = ds.kmeanspp(k, first, u)
        y_o, idx_o, dd_o = O.kmeanspp(x, k, seed=3)
        assert seeds.tolist() == idx_o.tolist() and np.array_equal(ds.mindist(), dd_o)
        cent, size = ds.init_centroids(k)
        ctx.set_assign_kernel(kern)
        inertia, sums, counts = ds.lloyd_step(cent)
        ctx.set_assign_kernel(cabi.ASSIGN_AUTO)
        d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent, want_gap=True)
        bad = np.nonzero(ds.labels().astype(np.int64) != m_o)[0]
        assert np.all(gap[bad] < (1e-5 if dt == np.float32 else 1e-12))
        assert len(bad) or counts.tolist() == c_o.tolist()
        assert abs(inertia - d_o) <= (1e-4 if dt == np.float32 else 1e-9) * d_o
        table = ds.contingency((np.arange(n) % 5).astype(np.uint32), 5, k)     # counted against the resident labels
        want = np.zeros((5, k), dtype=np.int64); np.add.at(want, (np.arange(n) % 5, ds.labels().astype(np.int64)), 1)
        assert np.array_equal(table, want)
        fit = ds.lloyd_fit(cent, 5)
        assert fit["size"].sum() == n
        assert np.array_equal(ctx.predict(x, fit["centroids"]).astype(np.int64), O.predict(x, fit["centroids"]))
        assert np.array_equal(ctx.predict(x, fit["centroids"], column_major=True).astype(np.int64), O.predict(x, fit["centroids"]))
        ds.close()
        print("ok", n, d, k, np.dtype(dt).name, kern, flush=True)
    # initial means of a larger fit go through the tile-layout update kernel (n >= 65536)
    g = ctx.generate_blobs(70000, 16, 16, 9)
    xs = g.download_rows(0, 70000)
    first, u = cluster.kmeanspp_draws(4, 70000, 16)
    g.kmeanspp(16, first, u)
    cent, size = g.init_centroids(16)
    lab = g.labels().astype(np.int64)
    for j in range(16):
        assert size[j] == (lab == j).sum() and np.allclose(cent[j], xs[lab == j].mean(axis=0), rtol=1e-9, atol=1e-9)
    g.close()
    print("ok init means 70000x16 k=16", flush=True)
    # batched k-NN: warp-private sorted lists in shared memory, merge across row chunks
    xk = rng.normal(size=(3000, 16)); qk = rng.normal(size=(11, 16))
    dk = ctx.upload(xk)
    idx, dist = dk.knn(qk, 33)
    for qi in range(11):
        d2 = ((xk - qk[qi]) ** 2).sum(axis=1)
        assert idx[qi].tolist() == np.lexsort((np.arange(3000), d2))[:33].tolist()
    res = dk.radius(qk[:5], 4.0)                            # find_radius: count / scan / fill
    for qi in range(5):
        full = np.sqrt(((xk - qk[qi]) ** 2).sum(axis=1))
        assert set(np.nonzero(full <= 4.0 * (1 - 1e-9))[0].tolist()) <= set(res[qi][0].tolist()) <= set(np.nonzero(full <= 4.0 * (1 + 1e-9))[0].tolist())
    dk.close()
    print("ok knn / radius 3000x16 k=33", flush=True)
    # rows that are not 16-byte multiples (element-wise staging) and k > 64 (passes of 64)
    xo = rng.normal(size=(900, 3)).astype(np.float32); qo = xo[:4] + np.float32(0.1)
    do = ctx.upload(xo)
    idx, dist = do.knn(qo, 100)
    for qi in range(4):
        d2 = ((xo.astype(np.float64) - qo[qi].astype(np.float64)) ** 2).sum(axis=1)
        assert set(idx[qi].tolist()) == set(np.argsort(d2, kind="stable")[:100].tolist()) and np.all(np.diff(dist[qi]) >= 0)
    res = do.radius(qo, 0.8)
    assert all(np.all(np.diff(i) > 0) and np.all(dv <= 0.8) for i, dv in res)
    do.close()
    print("ok knn / radius 900x3 f32 k=100", flush=True)
    g = ctx.generate_blobs(3000, 16, 8, 5)
    assert np.array_equal(g.download_rows(10, 20), cabi.blobs_host(10, 20, 16, 8, 5))
    g.close()
    ctx.close()
    # >= 2 GPUs: the sharded fit of a multi-GPU context -- its per-step sum over the devices runs inside the reduce /
    # finalize kernels over peer memory (csrc/sckm_peer.cu)
    os.environ["SCKM_MULTI_MIN_ROWS"] = "1"
    mc = sc.Context(devices="all")
    if mc.device_count() >= 2:
        xm = (rng.normal(size=(6000, 16)) + 4.0 * rng.integers(0, 6, size=(6000, 1))).astype(np.float64)
        first, u = cluster.kmeanspp_draws(11, 6000, 6)
        one = sc.Context(0)
        a = one.kmeans_fit(xm, 6, 30, first, u)
        b = mc.kmeans_fit(xm, 6, 30, first, u)
        assert mc.allreduce_path() in ("peer", "nccl") and b["iters"] == a["iters"] and np.array_equal(b["labels"], a["labels"])
        assert np.allclose(b["centroids"], a["centroids"], rtol=1e-9, atol=1e-12)
        one.close()
        print("ok multi-GPU fit over %d devices, all-reduce path: %s" % (mc.device_count(), mc.allreduce_path()), flush=True)
    mc.close()
    print("sanitizer smoke done")


if __name__ == "__main__":
    main()
