"""GPU (-m gpu): the CUDA path against the CPU oracle, through the C ABI.

Bars (BASELINE.json north_star): kmeans++ D^2 values, predict labels and direct-form Lloyd labels are
BIT-EXACT; GEMM-form Lloyd labels are identical except where the oracle's (second-best)/best squared
distance gap is < 1e-12 relative; centroids and inertia agree within 1e-9 relative (f64)."""
import numpy as np
import pytest

import smartcore_b200 as sc
from smartcore_b200 import cabi, cluster

pytestmark = pytest.mark.gpu

RTOL = 1e-9       # centroids / inertia, f64 (north_star)
GAP_TOL = 1e-12   # label tolerance, relative best/second-best gap (north_star)


def blobs(n, d, k, seed, dtype=np.float64, spread=4.0):
    rng = np.random.default_rng(seed)
    centers = rng.uniform(-10, 10, size=(k, d))
    x = centers[np.arange(n) % k] * (spread / 4.0) + rng.normal(size=(n, d))
    return x.astype(dtype)


def assert_labels_match(got, want, gap):
    bad = np.nonzero(np.asarray(got, dtype=np.int64) != np.asarray(want, dtype=np.int64))[0]
    assert np.all(gap[bad] < GAP_TOL), "labels differ at %d rows with gap >= %g" % (len(bad), GAP_TOL)


# ---- kmeans++ ------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,k,dtype", [(150, 4, 3, np.float64), (150, 4, 3, np.float32), (1000, 16, 8, np.float64),
                                         (5000, 7, 5, np.float64), (4097, 3, 4, np.float32), (33, 64, 6, np.float64)])
def test_kmeanspp_bit_exact(ctx, O, n, d, k, dtype):
    x = blobs(n, d, k, 100 + n + d, dtype)
    first, u = cluster.kmeanspp_draws(42, n, k)
    ds = ctx.upload(x)
    seeds = ds.kmeanspp(k, first, u)
    y_o, idx_o, dd_o = O.kmeanspp(x, k, seed=42)
    assert seeds.tolist() == idx_o.tolist()
    assert np.array_equal(ds.labels(), y_o.astype(np.uint64))
    assert np.array_equal(ds.mindist(), dd_o)          # bit-identical D^2 (euclidian.rs:56-63 rounding)
    ds.close()


@pytest.mark.parametrize("n,d,k,dtype,offset,scale", [
    (20000, 16, 24, np.float64, 0.0, 1.0), (20000, 64, 40, np.float64, 0.0, 1.0), (30000, 32, 33, np.float32, 0.0, 1.0),
    (12000, 8, 16, np.float64, 1e4, 1.0),       # far from the origin: bf16 keeps ~3 digits of (x - seed0), not of x
    (12000, 24, 16, np.float32, -300.0, 0.01),  # tight clusters, f32 element arithmetic
    (9000, 128, 12, np.float64, 5.0, 1e-3), (5000, 16, 64, np.float32, 0.0, 1e3), (4100, 40, 9, np.float64, 1e8, 1.0),
    (3000, 8, 10, np.float64, 0.0, 1e24),       # squares overflow f32 (not f64): the screening bound must stand down
    (3000, 16, 9, np.float64, 0.0, 1e-30)])     # squares underflow f32
def test_kmeanspp_pruned_passes_bit_exact(ctx, O, n, d, k, dtype, offset, scale, monkeypatch):
    """The pruned passes (triangle test, bf16-shadow screening, compacted exact pass) must leave every D^2, label and
    seed exactly as the reference's full passes do -- also when the screening bound is weak (large offsets), tight
    (tiny spreads), or meets duplicate rows (zero distances)."""
    x = (blobs(n, d, k, 11 * n + d, np.float64, spread=3.0) * scale + offset).astype(dtype)
    x[100:120] = x[7]                                   # duplicates of a row
    first, u = cluster.kmeanspp_draws(77, n, k)
    y_o, idx_o, dd_o = O.kmeanspp(x, k, seed=77)
    for env in ({}, {"SCKM_KPP_NOSHADOW": "1"}, {"SCKM_KPP_GEN1": "1"}, {"SCKM_KPP_NOPRUNE": "1"}):
        for key in ("SCKM_KPP_NOSHADOW", "SCKM_KPP_GEN1", "SCKM_KPP_NOPRUNE"):
            monkeypatch.delenv(key, raising=False)
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        ds = ctx.upload(x)
        seeds = ds.kmeanspp(k, first, u)
        assert seeds.tolist() == idx_o.tolist(), env
        assert np.array_equal(ds.labels(), y_o.astype(np.uint64)), env
        assert np.array_equal(ds.mindist(), dd_o), env
        ds.close()


def test_kmeanspp_injected_rows_and_iris(ctx, O, iris_f32):
    x = iris_f32[0]
    inj = np.array([7, 77, 140, 3])
    for data in (x, x.astype(np.float64)):
        ds = ctx.upload(data)
        seeds = ds.kmeanspp(4, inject_rows=inj)
        y_o, idx_o, dd_o = O.kmeanspp(data, 4, inject=inj)
        assert seeds.tolist() == inj.tolist() == idx_o.tolist()
        assert np.array_equal(ds.labels(), y_o.astype(np.uint64)) and np.array_equal(ds.mindist(), dd_o)
        ds.close()


# ---- initial centroids / Lloyd step ------------------------------------------------------------
def test_bbdtree_iris_golden_on_gpu(ctx, kat, iris20):  # the reference's only numeric golden (bbd_tree.rs:349-363)
    g = kat["bbdtree_iris"]
    for cm in (False, True):
        ds = ctx.upload(iris20, column_major=cm)
        inertia, sums, counts = ds.lloyd_step(g["centroids"])
        assert abs(inertia - g["cost"]) < g["cost_tol"]
        assert abs(sums[0][0] - g["sums_0_0"]) < g["sums_tol"] and abs(sums[1][3] - g["sums_1_3"]) < g["sums_tol"]
        assert ds.labels()[17] == g["membership_17"] and counts.tolist() == [10, 10]
        ds.close()


@pytest.mark.parametrize("n,d,k,dtype", [(150, 4, 3, np.float64), (2000, 16, 8, np.float64), (3000, 64, 32, np.float64),
                                         (1537, 5, 7, np.float64), (2048, 32, 16, np.float32), (999, 128, 10, np.float64)])
@pytest.mark.parametrize("kernel", [cabi.ASSIGN_DIRECT, cabi.ASSIGN_AUTO])
def test_lloyd_step_matches_oracle(ctx, O, n, d, k, dtype, kernel):
    x = blobs(n, d, k, 7 * n + d, dtype)
    cent = x[np.random.default_rng(3).choice(n, k, replace=False)].astype(np.float64) + 0.05
    ctx.set_assign_kernel(kernel)
    ds = ctx.upload(x)
    inertia, sums, counts = ds.lloyd_step(cent)
    ctx.set_assign_kernel(cabi.ASSIGN_AUTO)
    d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent, want_gap=True)
    labels = ds.labels()
    if kernel == cabi.ASSIGN_DIRECT:
        assert np.array_equal(labels, m_o.astype(np.uint64))       # direct form: bit-exact argmin
        md = ds.mindist()                                          # and bit-exact min distances
        for i in range(0, n, 97):
            assert md[i] == O.squared_distance(x[i].astype(np.float64), cent[m_o[i]])
    assert_labels_match(labels, m_o, gap)
    if np.array_equal(labels.astype(np.int64), m_o):
        assert counts.tolist() == c_o.tolist()
        np.testing.assert_allclose(sums, s_o, rtol=RTOL, atol=1e-9)
    assert abs(inertia - d_o) <= RTOL * d_o
    # and against the reference's own tree-filter path
    d_t, s_t, c_t, m_t = O.BBDTree(x).clustering(cent)
    assert_labels_match(labels, m_t, gap)
    assert abs(inertia - d_t) <= RTOL * d_t
    ds.close()


def test_dmma_ties_duplicates_and_exact_refine(ctx, O):
    """Degenerate data: duplicated rows and duplicated centroids make every row an exact tie in GEMM form;
    the refine pass must reproduce the reference's strict-< / lowest-index rule bit for bit."""
    rng = np.random.default_rng(0)
    base = rng.normal(size=(16, 32))
    x = base[rng.integers(0, 16, size=4000)]
    cent = np.vstack([base, base, base[:8] + 1e-13])            # k = 40: exact and 1e-13 near-duplicates
    ctx.set_assign_kernel(cabi.ASSIGN_DMMA)
    ds = ctx.upload(x)
    inertia, sums, counts = ds.lloyd_step(cent)
    ctx.set_assign_kernel(cabi.ASSIGN_AUTO)
    d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent, want_gap=True)
    assert np.array_equal(ds.labels().astype(np.int64), m_o)     # includes gap == 0 rows
    assert counts.tolist() == c_o.tolist()
    np.testing.assert_allclose(sums, s_o, rtol=RTOL, atol=1e-9)
    assert abs(inertia - d_o) <= 1e-9 * max(d_o, 1e-30) + 1e-18
    ds.close()


@pytest.mark.parametrize("tma_ring", [False, True, "ldg256"])
def test_stream_kernel_ties_and_shapes(ctx, O, tma_ring, monkeypatch):
    """Small-k streaming kernel (k < 16): exact ties, odd d (scalar staging path), f32, ragged n; with the rows
    loaded straight into registers (default), through the cp.async.bulk ring (SCKM_STREAM_TMA), and with 32-byte loads
    (SCKM_STREAM_256: the opt-in LDG.256 instantiation)."""
    monkeypatch.delenv("SCKM_STREAM_TMA", raising=False)
    monkeypatch.delenv("SCKM_STREAM_256", raising=False)
    if tma_ring == "ldg256":
        monkeypatch.setenv("SCKM_STREAM_256", "1")
    elif tma_ring:
        monkeypatch.setenv("SCKM_STREAM_TMA", "1")
    rng = np.random.default_rng(2)
    base = rng.normal(size=(6, 16))
    x = base[rng.integers(0, 6, size=3001)]
    cent = np.vstack([base, base])                                # k = 12, every row an exact tie
    for data, c in ((x, cent), (blobs(5003, 7, 5, 3), None), (blobs(4097, 32, 15, 4, np.float32), None),
                    (blobs(31, 2, 3, 5), None), (blobs(6000, 12, 9, 6, np.float32), None)):
        if c is None:
            kk = {7: 5, 32: 15, 2: 3, 12: 9}[data.shape[1]]
            c = data[np.random.default_rng(9).choice(len(data), kk, replace=False)].astype(np.float64) + 0.01
        ctx.set_assign_kernel(cabi.ASSIGN_STREAM)
        ds = ctx.upload(data)
        inertia, sums, counts = ds.lloyd_step(c)
        again = ds.lloyd_step(c)
        ctx.set_assign_kernel(cabi.ASSIGN_AUTO)
        d_o, s_o, c_o, m_o, gap = O.brute_clustering(data, c, want_gap=True)
        assert np.array_equal(ds.labels().astype(np.int64), m_o)
        assert counts.tolist() == c_o.tolist()
        np.testing.assert_allclose(sums, s_o, rtol=RTOL, atol=1e-9)
        assert abs(inertia - d_o) <= RTOL * max(d_o, 1e-30) + 1e-18
        assert again[0] == inertia and np.array_equal(again[1], sums)        # bit-reproducible
        ds.close()


def test_dmma_step_is_bit_reproducible(ctx):
    x = blobs(50000, 64, 64, 5, spread=1.0)
    cent = x[:64] + 0.01
    ctx.set_assign_kernel(cabi.ASSIGN_DMMA)
    ds = ctx.upload(x)
    a = ds.lloyd_step(cent); la = ds.labels()
    b = ds.lloyd_step(cent); lb = ds.labels()
    ctx.set_assign_kernel(cabi.ASSIGN_AUTO)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(la, lb)
    ds.close()


@pytest.mark.parametrize("n,d,k,dtype", [(20000, 64, 256, np.float64), (9000, 128, 100, np.float64), (7777, 20, 33, np.float64),
                                         (30000, 32, 512, np.float32), (5000, 12, 17, np.float32),
                                         # centroid sets larger than shared memory: streamed in blocks, row state parked
                                         (6001, 128, 300, np.float64), (20000, 64, 700, np.float64), (40000, 32, 1500, np.float32)])
def test_dmma_step_shapes(ctx, O, n, d, k, dtype):
    x = blobs(n, d, k, n + d, dtype, spread=1.5)
    cent = x[np.random.default_rng(1).choice(n, k, replace=False)].astype(np.float64) * 1.001
    ctx.set_assign_kernel(cabi.ASSIGN_DMMA)
    ds = ctx.upload(x)
    inertia, sums, counts = ds.lloyd_step(cent)
    ctx.set_assign_kernel(cabi.ASSIGN_AUTO)
    d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent, want_gap=True)
    assert np.array_equal(ds.labels().astype(np.int64), m_o)
    assert counts.tolist() == c_o.tolist()
    np.testing.assert_allclose(sums, s_o, rtol=RTOL, atol=1e-9)
    assert abs(inertia - d_o) <= RTOL * d_o
    ds.close()


@pytest.mark.parametrize("n,d,k,dtype", [(5000, 32, 16, np.float32), (33333, 32, 300, np.float32), (20000, 16, 128, np.float32),
                                         (7001, 8, 100, np.float32), (9000, 28, 257, np.float32), (50000, 32, 1500, np.float32),
                                         (255, 32, 64, np.float32),
                                         # two swizzle atoms along K (d <= 64), rows split over two column parts
                                         (30000, 64, 256, np.float32), (12345, 48, 100, np.float32), (9999, 36, 17, np.float32),
                                         # f64 data: tensor cores rank an f32 shadow copy, everything else stays f64
                                         (30000, 64, 256, np.float64), (20000, 32, 500, np.float64), (5001, 20, 40, np.float64)])
def test_tc5_tile_kernel(ctx, O, n, d, k, dtype):
    """tcgen05 3xTF32 kernel: the ranking runs in reduced precision, the decision must still be exact (labels equal
    the f64 direct-form argmin), sums/inertia in f64, bit-reproducible."""
    x = blobs(n, d, k, 3 * n + d, dtype, spread=1.5)
    cent = x[np.random.default_rng(2).choice(n, k, replace=False)].astype(np.float64) * 1.001
    ctx.set_assign_kernel(cabi.ASSIGN_TC5)
    ds = ctx.upload(x)
    inertia, sums, counts = ds.lloyd_step(cent)
    again = ds.lloyd_step(cent)
    ctx.set_assign_kernel(cabi.ASSIGN_AUTO)
    d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent, want_gap=True)
    assert np.array_equal(ds.labels().astype(np.int64), m_o)
    assert counts.tolist() == c_o.tolist()
    np.testing.assert_allclose(sums, s_o, rtol=RTOL, atol=1e-9)
    assert abs(inertia - d_o) <= RTOL * d_o
    assert again[0] == inertia and np.array_equal(again[1], sums) and np.array_equal(again[2], counts)
    ds.close()


@pytest.mark.parametrize("n,d,k", [(40000, 32, 512), (30000, 16, 128), (20000, 64, 300)])
def test_tc5_primed_steps_match_oracle(ctx, O, n, d, k, monkeypatch):
    """From the second step on, the tcgen05 kernel primes every row with the score of its previous centroid and skips
    32-column chunks that cannot matter.  The labels it starts from may be anything: stale ones from very different
    centroids (most rows change cluster), or exactly right ones (nothing changes) -- labels, sums and inertia must equal
    the oracle's either way, and equal the unprimed kernel's (SCKM_TC5_NOPRIME)."""
    x = blobs(n, d, k, 3 * n + d, np.float32, spread=1.5)
    rng = np.random.default_rng(4)
    cent_a = x[rng.choice(n, k, replace=False)].astype(np.float64) + 0.01
    cent_b = x[rng.choice(n, k, replace=False)].astype(np.float64) - 0.02          # unrelated centroids
    ds = ctx.upload(x)
    ds.lloyd_step(cent_a)                                                          # unprimed (no labels yet); leaves labels for cent_a
    for cent in (cent_b, cent_b, cent_a):                                          # stale labels, exact labels, stale again
        inertia, sums, counts = ds.lloyd_step(cent)
        lab = ds.labels().astype(np.int64)
        d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent, want_gap=True)
        bad = np.nonzero(lab != m_o)[0]
        assert np.all(gap[bad] < 1e-5), "%d labels differ beyond the f32 tolerance" % len(bad)
        assert abs(inertia - d_o) <= 1e-4 * d_o
        if len(bad) == 0:
            assert counts.tolist() == c_o.tolist()
            np.testing.assert_allclose(sums, s_o, rtol=1e-9, atol=1e-6)
    # a labelling that points every row at a far-away centroid (worst case for the priming bound)
    far = O.predict(x, -cent_b)
    ds2 = ctx.upload(x)
    ds2.lloyd_step(-cent_b)
    inertia, sums, counts = ds2.lloyd_step(cent_b)
    d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent_b, want_gap=True)
    bad = np.nonzero(ds2.labels().astype(np.int64) != m_o)[0]
    assert np.all(gap[bad] < 1e-5) and abs(inertia - d_o) <= 1e-4 * d_o
    monkeypatch.setenv("SCKM_TC5_NOPRIME", "1")
    i2, s2, c2 = ds2.lloyd_step(cent_b)
    assert i2 == inertia and np.array_equal(s2, sums) and np.array_equal(c2, counts)
    ds.close(); ds2.close()


def test_tc5_ties_and_duplicates(ctx, O):
    rng = np.random.default_rng(0)
    base = rng.normal(size=(16, 32)).astype(np.float32)
    x = base[rng.integers(0, 16, size=3000)]
    cent = np.vstack([base, base, base[:8] * (1 + 1e-7)]).astype(np.float64)     # exact and near duplicates (k = 40)
    ctx.set_assign_kernel(cabi.ASSIGN_TC5)
    ds = ctx.upload(x)
    inertia, sums, counts = ds.lloyd_step(cent)
    ctx.set_assign_kernel(cabi.ASSIGN_AUTO)
    d_o, s_o, c_o, m_o = O.brute_clustering(x, cent)
    assert np.array_equal(ds.labels().astype(np.int64), m_o) and counts.tolist() == c_o.tolist()
    ds.close()


@pytest.mark.parametrize("case", ["row_scales", "tiny_rows", "huge_scale", "tiny_scale", "extreme_scale", "zero_rows", "one_k_step"])
@pytest.mark.parametrize("fold", [True, False])
def test_tc5h_scaling_edge_cases(ctx, O, case, fold, monkeypatch):
    """The 3xFP16 kernel (d <= 32) scales every row and the centroids by powers of two to fit FP16's 5-bit exponent and
    folds -||c||^2/2 into the GEMM as a rank-one BF16 term.  Whatever the magnitudes -- rows of very different size, rows
    far below the centroids, data near either end of the f32 range, all-zero rows -- labels must equal the exact f64
    argmin (rows the reduced-precision ranking cannot decide are re-decided exactly), sums and inertia stay f64."""
    if not fold:
        monkeypatch.setenv("SCKM_TC5H_NOFOLD", "1")
    n, d, k = 20000, 32, 200
    rng = np.random.default_rng(11)
    x = blobs(n, d, k, 5, np.float32, spread=2.0).astype(np.float64)
    if case == "row_scales":
        x *= 10.0 ** rng.uniform(-3, 3, size=(n, 1))
    elif case == "tiny_rows":
        x[::2] *= 1e-12
    elif case == "huge_scale":
        x *= 1e15
    elif case == "tiny_scale":
        x *= 1e-18
    elif case == "extreme_scale":
        x *= 1e30
    elif case == "zero_rows":
        x[::7] = 0.0
    elif case == "one_k_step":
        d = 12
        x = x[:, :d].copy()
    x = np.ascontiguousarray(x.astype(np.float32))
    cent = x[rng.choice(n, k, replace=False)].astype(np.float64)
    if case == "zero_rows":
        cent[0] = 0.0
    cent = cent * (1.0 + 1e-3)
    ctx.set_assign_kernel(cabi.ASSIGN_TC5)
    ds = ctx.upload(x)
    inertia, sums, counts = ds.lloyd_step(cent)
    again = ds.lloyd_step(cent)                                    # primed by the labels of the first call
    ctx.set_assign_kernel(cabi.ASSIGN_AUTO)
    d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent, want_gap=True)
    assert np.array_equal(ds.labels().astype(np.int64), m_o)
    assert counts.tolist() == c_o.tolist()
    np.testing.assert_allclose(sums, s_o, rtol=RTOL, atol=1e-9 * float(np.abs(s_o).max()))
    assert abs(inertia - d_o) <= RTOL * d_o
    assert again[0] == inertia and np.array_equal(again[1], sums) and np.array_equal(again[2], counts)
    ds.close()


def test_tc5h_matches_the_tf32_kernel(ctx, O, monkeypatch):
    """Both tcgen05 forms (3xFP16 with 8 or 16 epilogue warps, and 3xTF32) only rank; what they hand on is exact, so a few
    Lloyd steps from the same start end in the same labels, sizes and bit-identical inertia trace sums."""
    n, d, k = 60000, 32, 700
    x = blobs(n, d, k, 9, np.float32, spread=1.5)
    cent = x[np.random.default_rng(3).choice(n, k, replace=False)].astype(np.float64) + 0.01
    res = {}
    for name, env in (("f16", {}), ("f16_ew16", {"SCKM_TC5H_EW16": "1"}), ("tf32", {"SCKM_TC5_TF32": "1"})):
        for key in ("SCKM_TC5H_EW16", "SCKM_TC5_TF32"):
            monkeypatch.delenv(key, raising=False)
        for key, v in env.items():
            monkeypatch.setenv(key, v)
        ctx.set_assign_kernel(cabi.ASSIGN_TC5)
        ds = ctx.upload(x)
        out = ds.lloyd_iterate(cent, 4, want_inertia=True)
        ctx.set_assign_kernel(cabi.ASSIGN_AUTO)
        res[name] = (ds.labels().copy(), out["size"].copy(), out["inertia"].copy(), out["centroids"].copy())
        ds.close()
    for name in ("f16_ew16", "tf32"):
        assert np.array_equal(res[name][0], res["f16"][0]) and np.array_equal(res[name][1], res["f16"][1])
        np.testing.assert_allclose(res[name][2], res["f16"][2], rtol=1e-12)
        np.testing.assert_allclose(res[name][3], res["f16"][3], rtol=1e-12, atol=1e-12)


def test_init_centroids_are_label_means(ctx, O):
    x = blobs(3000, 8, 6, 11)
    first, u = cluster.kmeanspp_draws(5, 3000, 6)
    ds = ctx.upload(x)
    ds.kmeanspp(6, first, u)
    cent, size = ds.init_centroids(6)
    y = ds.labels().astype(np.int64)
    for c in range(6):
        assert size[c] == (y == c).sum()
        np.testing.assert_allclose(cent[c], x[y == c].mean(0), rtol=1e-12)
    ds.close()


def test_empty_cluster_keeps_previous_centroid(ctx, O):  # kmeans.rs:298
    x = blobs(500, 4, 2, 21)
    cent = np.vstack([x[:2], [[1e6, 1e6, 1e6, 1e6]]])
    ds = ctx.upload(x)
    out = ds.lloyd_iterate(cent, 1)
    assert out["size"][2] == 0 and np.array_equal(out["centroids"][2], cent[2])
    ds.close()


# ---- full fit ------------------------------------------------------------------------------------
def fit_gpu(ctx, x, k, seed, max_iter=100, column_major=False):
    first, u = cluster.kmeanspp_draws(seed, x.shape[0], k)
    return ctx.kmeans_fit(x, k, max_iter, first, u, column_major=column_major)


def check_fit(O, x, k, seed, got):
    want = O.fit(x, k, 100, 0 if seed is None else seed, use_tree=True)
    if not np.array_equal(got["labels"].astype(np.int64), want.y):
        # tolerate only near-tie flips, judged against the final centroids
        gap = O.brute_clustering(x, want.centroids, want_gap=True)[4]
        assert_labels_match(got["labels"], want.y, gap)
    np.testing.assert_allclose(got["centroids"], want.centroids, rtol=RTOL, atol=1e-12)
    assert abs(got["distortion"] - want.distortion) <= RTOL * want.distortion
    assert got["size"].tolist() == want.size.tolist()
    # The loop ends when the inertia stops DECREASING (kmeans.rs:305).  At the fixed point (labels unchanged) successive
    # inertias are equal up to the rounding of two differently ordered sums, so `<=` can fire one step earlier or later
    # in two correct implementations; the state it ends in is the same one (asserted above).  Anything else must agree.
    assert got["iters"] == want.iters or (abs(got["iters"] - want.iters) == 1 and
                                          np.array_equal(got["labels"].astype(np.int64), want.y)), (got["iters"], want.iters)
    return want


def test_fit_iris_matches_oracle_and_fixture(ctx, O, iris_f32, oracle_fits):
    for dtype, key in ((np.float64, "iris_f64_k3_seed42_mode0"), (np.float32, "iris_f32_k3_seed42_mode0")):
        x = iris_f32[0].astype(dtype)
        for cm in (False, True):
            got = fit_gpu(ctx, x, 3, 42, column_major=cm)
            check_fit(O, x, 3, 42, got)
            g = oracle_fits[key]
            assert got["labels"].tolist() == g["y"] and got["size"].tolist() == g["size"] and got["iters"] == g["iters"]
            np.testing.assert_allclose(got["centroids"], np.array(g["centroids"]), rtol=RTOL)


@pytest.mark.parametrize("n,d,k,dtype,seed", [(20000, 16, 8, np.float64, 42), (6000, 64, 32, np.float64, 1),
                                              (10000, 32, 16, np.float32, 9), (4001, 3, 5, np.float64, None)])
def test_fit_blobs_matches_oracle(ctx, O, n, d, k, dtype, seed):
    x = blobs(n, d, k, n + 13, dtype, spread=2.0)
    got = fit_gpu(ctx, x, k, seed)
    check_fit(O, x, k, seed, got)


def test_fit_is_run_to_run_deterministic(ctx):
    x = blobs(30000, 16, 8, 77, spread=1.0)
    a = fit_gpu(ctx, x, 8, 3); b = fit_gpu(ctx, x, 8, 3)
    assert a["iters"] == b["iters"] and a["distortion"] == b["distortion"]
    assert np.array_equal(a["centroids"], b["centroids"]) and np.array_equal(a["labels"], b["labels"])


def test_device_stop_rule_is_independent_of_the_batch_size(ctx, O, monkeypatch):
    """The stop rule `if distortion <= dist { break }` (kmeans.rs:305-309) runs on the device; the host enqueues batches
    of iterations and kernels past the breaking iteration return at once.  Whatever the batch size, the fit must end in
    the state of the reference's `break`: same iteration count, distortion, centroids, sizes and labels."""
    for (n, d, k, dtype) in ((30000, 16, 8, np.float64), (20000, 64, 32, np.float64), (20000, 32, 64, np.float32), (3000, 3, 5, np.float64)):
        x = blobs(n, d, k, 5 * n + d, dtype, spread=1.5)
        monkeypatch.setenv("SCKM_LLOYD_BATCH", "1")
        a = fit_gpu(ctx, x, k, 11)
        want = check_fit(O, x, k, 11, a)
        assert 2 <= a["iters"] < 100                                 # the rule fired (not max_iter)
        for batch in ("2", "3", "8", "64"):
            monkeypatch.setenv("SCKM_LLOYD_BATCH", batch)
            b = fit_gpu(ctx, x, k, 11)
            assert b["iters"] == a["iters"] and b["distortion"] == a["distortion"]
            assert np.array_equal(b["centroids"], a["centroids"]) and np.array_equal(b["labels"], a["labels"])
            assert np.array_equal(b["size"], a["size"])
        monkeypatch.delenv("SCKM_LLOYD_BATCH")
        c = fit_gpu(ctx, x, k, 11)                                    # the shape-derived default
        assert c["iters"] == a["iters"] and np.array_equal(c["centroids"], a["centroids"]) and np.array_equal(c["labels"], a["labels"])
        # max_iter smaller than the stopping iteration: exactly max_iter steps, distortion = the last strictly smaller value
        m = max(1, a["iters"] - 2)
        monkeypatch.setenv("SCKM_LLOYD_BATCH", "8")
        e = fit_gpu(ctx, x, k, 11, max_iter=m)
        w = O.fit(x, k, m, 11)
        assert e["iters"] == m == w.iters and abs(e["distortion"] - w.distortion) <= RTOL * w.distortion
        np.testing.assert_allclose(e["centroids"], w.centroids, rtol=RTOL, atol=1e-12)
        monkeypatch.delenv("SCKM_LLOYD_BATCH")
        # the per-iteration inertias of the fixed-length loop (bench path) are non-increasing and end at the fit's value
        ds = ctx.upload(x)
        first, u = cluster.kmeanspp_draws(11, n, k)
        ds.kmeanspp(k, first, u)
        cent0, _ = ds.init_centroids(k)
        tr = ds.lloyd_iterate(cent0, a["iters"], want_inertia=True)
        assert np.all(np.diff(tr["inertia"][: a["iters"] - 1]) <= 0) and tr["inertia"][a["iters"] - 2] == a["distortion"]
        assert np.array_equal(tr["centroids"], a["centroids"])
        ds.close()


# ---- cancellation stress for the GEMM-form kernels: data far from the origin, tiny spreads ---------------------------
@pytest.mark.parametrize("n,d,k,dtype,offset,scale", [
    (12000, 64, 32, np.float64, 1e4, 1.0),      # DMMA tile kernel, resident centroids
    (12000, 64, 32, np.float64, 1e8, 1.0),      # ||x||^2 ~ 1e18: every row is a near-tie for the GEMM form -> exact re-decision
    (9000, 128, 300, np.float64, 1e4, 1e-3),    # DMMA with streamed centroid blocks, spread 1e-3 around 1e4
    (20000, 16, 8, np.float64, 1e4, 1.0),       # streaming kernel
    (20000, 16, 8, np.float64, 1e8, 1e-3),
    (16000, 8, 12, np.float64, -3e5, 1e-3),
    (15000, 32, 64, np.float32, 1e4, 1.0),      # tcgen05 3xTF32 kernel: f32 keeps ~3 digits of the spread at 1e4
    (15000, 32, 64, np.float32, 300.0, 1e-3),
    (10000, 24, 40, np.float32, 1e8, 1.0)])     # f32 data at 1e8: all rows collapse onto a few representable points
def test_lloyd_step_far_from_origin(ctx, O, n, d, k, dtype, offset, scale):
    """||x||^2 - 2 x.c + ||c||^2 loses everything when the offset dwarfs the spread; the tile kernels only RANK with
    it and re-decide every row whose gap is inside the error bound exactly, so labels must still be the oracle's,
    and sums / inertia within the north-star tolerance (f32 rows: the reference itself runs Lloyd in f64 on the
    widened values, bbd_tree.rs:207-213)."""
    x = (blobs(n, d, k, 3 * n + d, np.float64, spread=3.0) * scale + offset).astype(dtype)
    cent = x[np.random.default_rng(2).choice(n, k, replace=False)].astype(np.float64) + 0.05 * scale
    ds = ctx.upload(x)
    inertia, sums, counts = ds.lloyd_step(cent)
    d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent, want_gap=True)
    labels = ds.labels().astype(np.int64)
    tol = GAP_TOL if dtype == np.float64 else 1e-5
    bad = np.nonzero(labels != m_o)[0]
    assert np.all(gap[bad] < tol), "%d labels differ with gap >= %g" % (len(bad), tol)
    if len(bad) == 0:
        assert counts.tolist() == c_o.tolist()
        np.testing.assert_allclose(sums, s_o, rtol=RTOL, atol=0)
    assert abs(inertia - d_o) <= RTOL * d_o, (inertia, d_o)
    # whole fit from these centroids: iteration count and final state equal the oracle's tree path
    out = ds.lloyd_fit(cent, 30)
    ds.close()
    tree = O.BBDTree(x)
    c_ref = cent.copy(); dist_ref = np.finfo(np.float64).max; it_ref = 0
    for it in range(1, 31):
        dd, ss, cc, mm = tree.clustering(c_ref)
        nz = cc > 0
        c_ref[nz] = ss[nz] / cc[nz, None]
        it_ref = it
        if dist_ref <= dd:
            break
        dist_ref = dd
    # Conditioning: centroids are means rounded to one ulp of |x|; two correct implementations that add the rows in a
    # different order (tree traversal vs row order) differ by that much, and the inertia follows with
    # |d inertia| <= 2 sqrt(inertia * n) * |d c|.  At offset 1e8 / spread 1e-3 this is 1e-5 of the inertia -- the data,
    # not the kernel; everywhere else it is far below the north-star tolerance, which then is the bar.
    cond = 2.0 * np.sqrt(dist_ref * n) * float(np.spacing(np.abs(x.astype(np.float64)).max())) * np.sqrt(d)
    rt = RTOL if dtype == np.float64 else 1e-4
    well = cond <= rt * dist_ref
    if dtype == np.float64 and well:
        assert out["iters"] == it_ref
    if out["iters"] == it_ref:
        np.testing.assert_allclose(out["centroids"], c_ref, rtol=rt, atol=0)
        assert abs(out["distortion"] - dist_ref) <= rt * dist_ref + cond, (out["distortion"], dist_ref, cond)


@pytest.mark.parametrize("force", ["1", "0"])
def test_centring_switch_does_not_change_results(ctx, O, force, monkeypatch):
    """The tile kernels have a centred (x - mu, c - mu) and a plain instantiation; launch_cnorm picks by where the data
    sit.  Forced either way (SCKM_CENTER) on ordinary and on offset data, labels must equal the oracle's and sums /
    inertia / centroids stay within the north-star tolerance; predict labels are bit-equal to the direct form."""
    monkeypatch.setenv("SCKM_CENTER", force)
    for (n, d, k, dtype, offset) in ((20000, 64, 256, np.float64, 0.0), (6001, 128, 300, np.float64, 0.0), (9000, 20, 33, np.float64, 50.0),
                                     (12000, 48, 40, np.float32, 0.0), (8000, 64, 64, np.float64, 1e3)):
        x = (blobs(n, d, k, n + d, np.float64, spread=1.5) + offset).astype(dtype)
        cent = x[np.random.default_rng(1).choice(n, k, replace=False)].astype(np.float64) * 1.001
        ctx.set_assign_kernel(cabi.ASSIGN_DMMA)
        ds = ctx.upload(x)
        inertia, sums, counts = ds.lloyd_step(cent)
        ctx.set_assign_kernel(cabi.ASSIGN_AUTO)
        d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent, want_gap=True)
        assert np.array_equal(ds.labels().astype(np.int64), m_o)
        assert counts.tolist() == c_o.tolist()
        np.testing.assert_allclose(sums, s_o, rtol=RTOL, atol=1e-9)
        assert abs(inertia - d_o) <= RTOL * d_o
        out = ds.lloyd_fit(cent, 5)
        ds.close()
        c_ref = cent.copy()
        for _ in range(out["iters"]):
            dd, ss, cc, mm = O.brute_clustering(x, c_ref)
            nz = cc > 0
            c_ref[nz] = ss[nz] / cc[nz, None]
        np.testing.assert_allclose(out["centroids"], c_ref, rtol=RTOL if dtype == np.float64 else 1e-6, atol=1e-12)
        assert np.array_equal(ctx.predict(x, cent).astype(np.int64), O.predict(x, cent))


@pytest.mark.parametrize("n,d,k,dtype", [(9000, 64, 2, np.float64), (12000, 48, 5, np.float64), (7000, 128, 8, np.float64),
                                         (10000, 100, 15, np.float64), (8000, 40, 11, np.float32), (5000, 36, 3, np.float64)])
def test_small_k_wide_rows_take_the_tile_kernel(ctx, O, n, d, k, dtype):
    """k < 16 with d > 32 (outside the streaming kernel): the DMMA tile kernel with a partly padded sub-block instead
    of the direct form.  Step and whole fit against the oracle."""
    x = blobs(n, d, k, 2 * n + d + k, dtype, spread=1.0)
    cent = x[np.random.default_rng(5).choice(n, k, replace=False)].astype(np.float64) + 0.01
    ds = ctx.upload(x)
    l0 = ctx.launch_count()
    inertia, sums, counts = ds.lloyd_step(cent)
    d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent, want_gap=True)
    assert_labels_match(ds.labels(), m_o, gap)
    assert abs(inertia - d_o) <= RTOL * d_o
    if np.array_equal(ds.labels().astype(np.int64), m_o):
        assert counts.tolist() == c_o.tolist()
        np.testing.assert_allclose(sums, s_o, rtol=RTOL, atol=1e-9)
    ds.close()
    got = fit_gpu(ctx, x, k, 3)
    check_fit(O, x, k, 3, got)


def test_step_is_bit_reproducible_at_full_size(ctx):
    """The stop rule compares successive inertias exactly, so a step must be bit-reproducible run to run: config C3
    (10M x 64, k = 256), five repeats of the same step -- packed sums, counts, inertia and labels identical."""
    n, d, k = 10_000_000, 64, 256
    ds = ctx.generate_blobs(n, d, k, 20260101)
    cent = np.vstack([ds.download_rows(i * (n // k), 1) for i in range(k)]).astype(np.float64) + 0.01
    ref = ds.lloyd_step(cent)
    lab = ds.labels(width=4)
    for _ in range(4):
        got = ds.lloyd_step(cent)
        assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
        assert np.array_equal(ds.labels(width=4), lab)
    assert ref[2].sum() == n
    ds.close()


# ---- full KMeans::fit against the oracle's tree path with the k and d of configs C3 / C4 / C5 (SURVEY 8d gate) -------
@pytest.mark.timeout(900)      # the oracle's tree path needs ~1-2 minutes of one host core at k = 1024 / 4096
@pytest.mark.parametrize("n,d,k,dtype,max_iter", [(200_000, 64, 256, np.float64, 8), (200_000, 128, 1024, np.float64, 3),
                                                  (200_000, 32, 4096, np.float32, 2)])
def test_fit_parity_with_config_k_and_d(ctx, O, n, d, k, dtype, max_iter):
    x = cabi.blobs_host(0, n, d, k, 20260101, dtype=dtype)
    got = fit_gpu(ctx, x, k, 42, max_iter=max_iter)
    want = O.fit(x, k, max_iter, 42, use_tree=True)
    f64 = dtype == np.float64
    lab = got["labels"].astype(np.int64)
    if not np.array_equal(lab, want.y):
        gap = O.brute_clustering(x, want.centroids, want_gap=True)[4]
        bad = np.nonzero(lab != want.y)[0]
        assert np.all(gap[bad] < (GAP_TOL if f64 else 1e-5)), "%d labels differ beyond the tolerance" % len(bad)
    assert got["iters"] == want.iters
    np.testing.assert_allclose(got["centroids"], want.centroids, rtol=RTOL if f64 else 1e-4, atol=1e-12)
    assert abs(got["distortion"] - want.distortion) <= (RTOL if f64 else 1e-4) * want.distortion
    if np.array_equal(lab, want.y):
        assert got["size"].tolist() == want.size.tolist()


# ---- ONE context over several devices (sckm_ctx_create_multi): KMeans::fit / predict use the whole box ----------------
def test_multi_context_on_one_device_is_the_plain_context(O):
    c = sc.Context(devices=[0])
    assert c.device_count() == 1
    x = blobs(5000, 8, 4, 1)
    got = fit_gpu(c, x, 4, 3)
    check_fit(O, x, 4, 3, got)
    t = c.last_fit_times()
    assert t["devices"] == 1 and t["total_s"] > 0
    c.close()
    with pytest.raises(cabi.SckmError):
        sc.Context(devices=[0, 0])


def test_fit_shard_on_one_rank_is_the_whole_fit(ctx, O):
    """sckm_kmeans_fit_shard (the per-rank form for torchrun deployments) with one rank holding every row must be
    sckm_kmeans_fit; the phase clock of the last fit is filled in."""
    x = blobs(30000, 24, 12, 8, spread=2.0)
    first, u = cluster.kmeanspp_draws(9, 30000, 12)
    a = ctx.kmeans_fit(x, 12, 50, first, u)
    b = ctx.kmeans_fit_shard(x, 0, 30000, 12, 50, first, u)
    assert a["iters"] == b["iters"] and a["distortion"] == b["distortion"] and np.array_equal(a["labels"], b["labels"])
    assert np.array_equal(a["centroids"], b["centroids"]) and np.array_equal(a["size"], b["size"])
    t = ctx.last_fit_times()
    assert t["devices"] == 1 and t["total_s"] >= t["lloyd_s"] > 0 and ctx.device_count() == 1
    check_fit(O, x, 12, 9, b)
    with pytest.raises(cabi.SckmError):
        ctx.kmeans_fit_shard(x, 5, 30000, 12, 50, first, u)          # rows [5, 30005) exceed n_global


@pytest.mark.parametrize("n,d,k,dtype", [(50_000, 32, 24, np.float64), (300_000, 64, 256, np.float64), (40_000, 16, 8, np.float64),
                                         (120_000, 32, 128, np.float32), (5_000, 3, 4, np.float64)])
def test_multi_context_fit_and_predict_equal_single_device(ctx, O, n, d, k, dtype, monkeypatch):
    """Rows sharded over every visible device behind ONE context and ONE host thread of the caller: same seeds, same
    iteration count, same labels as the single-device fit; sums are added in a different order (per-shard partials, then
    the all-reduce), hence centroids within the north-star tolerance instead of bit-equal."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    monkeypatch.setenv("SCKM_MULTI_MIN_ROWS", "1")
    mc = sc.Context(devices="all")
    assert mc.device_count() == ndev
    x = blobs(n, d, k, n + 3 * d, dtype, spread=2.0)
    one = fit_gpu(ctx, x, k, 5, max_iter=40)
    per = -(-(-(-n // ndev)) // 1024) * 1024                       # rows per device, aligned like dist.shard_range
    sharded = per * (ndev - 1) < n                                  # every device gets a non-empty block?
    for cm in (False, True):
        many = fit_gpu(mc, x, k, 5, max_iter=40, column_major=cm)
        assert mc.last_fit_times()["devices"] == (ndev if sharded else 1)
        if sharded and torch.cuda.can_device_access_peer(0, 1):
            assert mc.allreduce_path() == "peer"                    # the sum over the devices ran inside the finalize kernel
        assert many["iters"] == one["iters"] and many["size"].tolist() == one["size"].tolist()
        assert np.array_equal(many["labels"], one["labels"])
        rt = RTOL if dtype == np.float64 else 1e-4
        np.testing.assert_allclose(many["centroids"], one["centroids"], rtol=rt, atol=1e-12)
        assert abs(many["distortion"] - one["distortion"]) <= rt * one["distortion"]
    if n <= 60_000:
        check_fit(O, x, k, 5, fit_gpu(mc, x, k, 5))
    if sharded:                                                      # the same fit with the all-reduce through NCCL
        monkeypatch.setenv("SCKM_PEER_ALLREDUCE", "0")
        nccl = fit_gpu(mc, x, k, 5, max_iter=40)
        monkeypatch.delenv("SCKM_PEER_ALLREDUCE")
        assert mc.allreduce_path() == "nccl" and nccl["iters"] == many["iters"] and np.array_equal(nccl["labels"], many["labels"])
        np.testing.assert_allclose(nccl["centroids"], many["centroids"], rtol=1e-12, atol=0)
        if ndev == 2:
            assert np.array_equal(nccl["centroids"], many["centroids"])
    # predict: no collective, labels bit-equal to the single-device direct form
    assert np.array_equal(mc.predict(x, one["centroids"]), ctx.predict(x, one["centroids"]))
    assert np.array_equal(mc.predict(x, one["centroids"], column_major=True, width=4), ctx.predict(x, one["centroids"], width=4))
    # an input too small to give every device a non-empty aligned share runs on the first device alone
    small = blobs(1000, d, k, 9, dtype)
    s1 = fit_gpu(mc, small, min(k, 8), 2)
    assert mc.last_fit_times()["devices"] == 1
    s0 = fit_gpu(ctx, small, min(k, 8), 2)
    assert np.array_equal(s1["labels"], s0["labels"]) and np.array_equal(s1["centroids"], s0["centroids"])
    mc.close()


# ---- host mirror: the reference's own tests, re-expressed ----------------------------------------------
def test_fit_predict_reference_test(iris20):  # kmeans.rs:473-505
    x = sc.DenseMatrix.from_2d_array(iris20)
    kmeans = sc.KMeans.fit(x, sc.KMeansParameters.default())
    y = kmeans.predict(x)
    assert np.array_equal(y, kmeans._y) and kmeans.size.sum() == 20


def test_f32_model_and_u8_labels(iris20):  # doctest kmeans.rs:18-48 (Vec<u8>) and serde test's f32 model
    x = sc.DenseMatrix.from_2d_array(iris20.astype(np.float32))
    kmeans = sc.KMeans.fit(x, sc.KMeansParameters.default().with_k(2))
    y = kmeans.predict(x, ty=np.uint8)
    assert y.dtype == np.uint8 and np.array_equal(y, kmeans._y)
    with pytest.raises(sc.Failed):
        kmeans.predict(sc.DenseMatrix.from_2d_array(iris20[:, :3].astype(np.float32)))


# ---- predict ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,k,dtype", [(1, 4, 2, np.float64), (777, 16, 8, np.float32), (5000, 64, 40, np.float64),
                                         (300, 9, 3, np.float64)])
def test_predict_bit_exact(ctx, O, n, d, k, dtype):
    x = blobs(n, d, k, n + 5, dtype)
    cent = blobs(k, d, k, 99).astype(np.float64)
    want = O.predict(x, cent)
    assert np.array_equal(ctx.predict(x, cent).astype(np.int64), want)
    assert np.array_equal(ctx.predict(x, cent, column_major=True, width=4).astype(np.int64), want)


def test_predict_tie_break_lowest_index(ctx):
    x = np.zeros((40, 4)); cent = np.ones((5, 4)); cent[3] = 0.5
    assert np.all(ctx.predict(x, cent) == 3)
    cent[:] = 1.0
    assert np.all(ctx.predict(x, cent) == 0)          # exact tie: strict < keeps the first


@pytest.mark.parametrize("n,d,k,dtype", [(200000, 64, 256, np.float64), (50000, 32, 100, np.float32), (30000, 128, 300, np.float64),
                                         (40000, 64, 1000, np.float64), (25000, 20, 33, np.float32), (65537, 16, 64, np.float64)])
def test_predict_at_scale_bit_exact(ctx, O, n, d, k, dtype):
    """k >= 32 routes predict through the DMMA ranking + exact re-decision of near-ties; the labels must still be
    those of the direct form (kmeans.rs:334-347) for EVERY row, duplicates and exact ties included."""
    x = blobs(n, d, min(k, 50), n + d, dtype, spread=1.0)
    cent = x[np.random.default_rng(5).choice(n, k, replace=False)].astype(np.float64)
    cent[7] = cent[3]                                  # duplicate centroid: rows of cluster 3 tie exactly -> lowest index
    x[100:110] = 0.5 * (cent[1] + cent[2]).astype(dtype)   # rows (nearly) equidistant from two centroids
    want = O.predict(x, cent)
    got = ctx.predict(x, cent).astype(np.int64)
    assert np.array_equal(got, want)
    assert not np.any(got == 7)
    assert np.array_equal(ctx.predict(x, cent, column_major=True, width=4).astype(np.int64), want)
    ctx.set_assign_kernel(cabi.ASSIGN_DIRECT)          # and the direct-form kernel agrees
    try:
        assert np.array_equal(ctx.predict(x, cent).astype(np.int64), want)
    finally:
        ctx.set_assign_kernel(cabi.ASSIGN_AUTO)


def test_predict_streams_in_chunks(ctx, O, monkeypatch):
    """sckm_predict double-buffers X through the device in row chunks (ragged last chunk, both layouts, both widths)."""
    monkeypatch.setenv("SCKM_PREDICT_CHUNK_ROWS", "1000")
    for n, d, k, dtype in ((4321, 16, 40, np.float64), (2500, 8, 5, np.float32), (1000, 64, 64, np.float64), (999, 4, 3, np.float64)):
        x = blobs(n, d, k, n, dtype)
        cent = blobs(k, d, k, 7).astype(np.float64)
        want = O.predict(x, cent)
        assert np.array_equal(ctx.predict(x, cent).astype(np.int64), want)
        assert np.array_equal(ctx.predict(x, cent, column_major=True, width=4).astype(np.int64), want)


def test_pageable_transfers_through_the_pinned_ring(ctx, O, monkeypatch):
    """> 32 MB from/to ordinary (pageable) numpy memory goes through the threaded staging ring in both directions
    (forced here: by default a lane is only pinned for >= 128 MB of traffic)."""
    monkeypatch.setenv("SCKM_INGEST_FORCE", "1")
    n, d = 5_000_011, 2                                    # 40 MB of f32 rows, 40 MB of u64 labels; ragged tail chunk
    x = blobs(n, d, 2, 17, np.float32, spread=8.0)
    ds = ctx.upload(x)
    assert np.array_equal(ds.download_rows(0, 4096), x[:4096])
    assert np.array_equal(ds.download_rows(n - 4099, 4099), x[n - 4099:])
    assert np.array_equal(ds.download_rows(2_500_000, 1000), x[2_500_000:2_501_000])
    cent = np.array([x[0], x[1]], dtype=np.float64)
    inertia, sums, counts = ds.lloyd_step(cent)
    labels = ds.labels()                                   # staged D2H
    ds.close()
    d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent, want_gap=True)
    assert_labels_match(labels, m_o, gap)
    assert counts.sum() == n
    dsc = ctx.upload(x, column_major=True)                 # column-major image + device transpose
    assert np.array_equal(dsc.download_rows(n - 50, 50), x[n - 50:])
    dsc.close()


# ---- edge cases: tiny, ragged and degenerate inputs ------------------------------------------------------
@pytest.mark.parametrize("n,d,k", [(2, 1, 2), (3, 2, 2), (5, 3, 5), (17, 1, 4), (64, 2, 63), (31, 3, 30), (257, 5, 256),
                                   (100, 4, 99), (40, 130, 7), (1000, 1, 16), (129, 4, 16), (4096, 6, 17)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fit_edge_shapes(ctx, O, n, d, k, dtype):
    x = blobs(n, d, max(2, k // 2), 31 * n + d, dtype, spread=3.0)
    got = fit_gpu(ctx, x, k, 11)
    want = O.fit(x, k, 100, 11, use_tree=True)
    assert got["iters"] == want.iters and got["size"].tolist() == want.size.tolist()
    if not np.array_equal(got["labels"].astype(np.int64), want.y):
        gap = O.brute_clustering(x, want.centroids, want_gap=True)[4]
        assert_labels_match(got["labels"], want.y, gap)
    np.testing.assert_allclose(got["centroids"], want.centroids, rtol=RTOL if dtype == np.float64 else 1e-4, atol=1e-12,
                               equal_nan=True)
    # GEMM-form distances ||x||^2 - 2 x.c + ||c||^2 carry an ABSOLUTE rounding floor of a few ulp of ||x||^2 per row:
    # it only shows when the true inertia is ~0 (every point on its centroid, k ~ n)
    floor = 64 * np.finfo(np.float64).eps * float(np.sum(x.astype(np.float64) ** 2))
    assert abs(got["distortion"] - want.distortion) <= RTOL * want.distortion + floor
    pred = ctx.predict(x, got["centroids"]).astype(np.int64)
    assert np.array_equal(pred, O.predict(x, got["centroids"]))


@pytest.mark.parametrize("kernel", [cabi.ASSIGN_DIRECT, cabi.ASSIGN_AUTO, cabi.ASSIGN_DMMA])
def test_step_duplicates_and_constant_columns(ctx, O, kernel):
    """Every row duplicated 4x, two constant columns, centroids ON data points (zero distances)."""
    base = blobs(600, 8, 20, 5)
    base[:, 2] = 3.25; base[:, 5] = 0.0
    x = np.repeat(base, 4, axis=0)
    cent = base[:20].copy()
    ctx.set_assign_kernel(kernel)
    try:
        ds = ctx.upload(x)
        inertia, sums, counts = ds.lloyd_step(cent)
        labels = ds.labels()
        ds.close()
    finally:
        ctx.set_assign_kernel(cabi.ASSIGN_AUTO)
    d_o, s_o, c_o, m_o, gap = O.brute_clustering(x, cent, want_gap=True)
    assert_labels_match(labels, m_o, gap)
    assert np.array_equal(labels.reshape(-1, 4), np.repeat(labels[::4, None], 4, axis=1))   # duplicates stay together
    assert counts.tolist() == c_o.tolist() and abs(inertia - d_o) <= RTOL * d_o
    np.testing.assert_allclose(sums, s_o, rtol=RTOL, atol=1e-9)


def test_all_rows_identical(ctx, O):
    """All D^2 are zero: kmeans++ keeps drawing row 0, k-1 clusters start empty (0/0 centroids, kmeans.rs:276-280)."""
    x = np.full((50, 4), 2.5)
    got = fit_gpu(ctx, x, 3, 1)
    want = O.fit(x, 3, 100, 1, use_tree=True)
    assert got["size"].tolist() == want.size.tolist() and got["iters"] == want.iters
    assert np.array_equal(got["labels"].astype(np.int64), want.y)
    assert np.array_equal(np.isnan(got["centroids"]), np.isnan(want.centroids))
    np.testing.assert_allclose(got["centroids"], want.centroids, rtol=RTOL, equal_nan=True)


# ---- generator twin, properties at scale --------------------------------------------------------------
def test_device_blobs_equal_host_twin(ctx):
    for dtype in (np.float64, np.float32):
        ds = ctx.generate_blobs(5000, 16, 8, 20260101, dtype=dtype, row_offset=123, n_global=10000)
        got = ds.download_rows(100, 64)
        assert np.array_equal(got, cabi.blobs_host(223, 64, 16, 8, 20260101, dtype=dtype))
        ds.close()


def test_config2_1Mx16_k8_properties(ctx, O):
    """BASELINE config 2 at full size: size-independent properties + sampled-row oracle check."""
    n, d, k = 1_000_000, 16, 8
    ds = ctx.generate_blobs(n, d, k, 20260101)
    first, u = cluster.kmeanspp_draws(42, n, k)
    seeds = ds.kmeanspp(k, first, u)
    assert seeds[0] == first and len(set(seeds.tolist())) == k
    cent, size = ds.init_centroids(k)
    assert size.sum() == n
    out = ds.lloyd_fit(cent, 100)
    assert out["size"].sum() == n and 1 <= out["iters"] <= 100
    labels = ds.labels().astype(np.int64)
    assert np.array_equal(np.bincount(labels, minlength=k), out["size"])
    # monotone inertia: one more step from the final centroids cannot increase it beyond rounding
    inertia, sums, counts = ds.lloyd_step(out["centroids"])
    assert inertia <= out["distortion"] * (1 + 1e-12)
    np.testing.assert_allclose(sums.sum(0) / n, np.mean(cabi.blobs_host(0, 20000, d, k, 20260101), 0), atol=0.2)
    # sampled rows regenerated on the host, labels recomputed by the oracle against the GPU's centroids
    rows = np.random.default_rng(0).choice(n, 4096, replace=False)
    xs = np.vstack([cabi.blobs_host(int(r), 1, d, k, 20260101) for r in rows])
    d_o, s_o, c_o, m_o, gap = O.brute_clustering(xs, out["centroids"], want_gap=True)
    assert_labels_match(ds.labels()[rows], m_o, gap)
    ds.close()


def test_config2_fit_parity_200k(ctx, O):
    n, d, k = 200_000, 16, 8
    x = cabi.blobs_host(0, n, d, k, 20260101)
    got = fit_gpu(ctx, x, k, 42)
    check_fit(O, x, k, 42, got)


def test_serde_round_trip_reference_test(iris20):  # kmeans.rs:507-545: f32 model, serde_json round trip, PartialEq
    x = sc.DenseMatrix.from_2d_array(iris20.astype(np.float32))
    kmeans = sc.KMeans.fit(x, sc.KMeansParameters.default())
    for back in (sc.KMeans.from_json(kmeans.to_json(), dtype=np.float32),
                 sc.KMeans.from_bincode(kmeans.to_bincode(), dtype=np.float32)):
        assert back == kmeans
        assert np.array_equal(back._y, kmeans._y) and back._distortion == kmeans._distortion
        assert np.array_equal(back.predict(x), kmeans.predict(x))      # a reloaded model predicts on the device


# ---- cluster-quality scores (metrics/cluster_helpers.rs, cluster_hcv.rs): counting on the device --------------------
def test_hcv_reference_known_answers():  # cluster_helpers.rs:117-147, cluster_hcv.rs:94-104
    v1 = [0, 0, 1, 1, 2, 0, 4]
    v2 = [1, 0, 0, 0, 0, 1, 0]
    assert sc.contingency_matrix(v1, v2).tolist() == [[1, 2], [2, 0], [1, 0], [1, 0]]
    assert abs(1.2770 - sc.entropy(v1)) < 1e-4
    assert abs(0.3254 - sc.mutual_info_score(sc.contingency_matrix(v1, v2))) < 1e-4
    s = sc.HCVScore.new().compute(v1, v2)
    assert abs(0.2548 - s.homogeneity()) < 1e-4 and abs(0.5440 - s.completeness()) < 1e-4 and abs(0.3471 - s.v_measure()) < 1e-4


@pytest.mark.parametrize("n,na,nb", [(1, 1, 1), (1000, 7, 5), (200000, 40, 33), (300000, 120, 100), (5000, 1, 9)])
def test_hcv_matches_oracle(n, na, nb):
    from oracle import metrics_oracle as M
    rng = np.random.default_rng(n + na)
    a = rng.integers(0, na, n) * 5 - 7                     # arbitrary (negative, sparse) label values
    b = (a // 5 + rng.integers(0, nb, n) * (rng.random(n) < 0.4)) % nb
    assert sc.contingency_matrix(a, b).tolist() == M.contingency_matrix(a, b)          # integer counts: exact
    assert abs(sc.entropy(a) - M.entropy(a)) <= 1e-12 * max(1.0, M.entropy(a))
    s = sc.HCVScore.new().compute(a, b)
    for got, want in zip((s.homogeneity(), s.completeness(), s.v_measure()), M.hcv(a, b)):
        if not np.isfinite(want):      # a single class: the reference divides a rounding-level MI by a zero entropy
            assert not np.isfinite(got)
        else:
            assert abs(got - want) <= 1e-12 * max(1.0, abs(want))


def test_contingency_of_resident_labels(ctx, O):
    """After a fit the labels never leave the device: sckm_contingency counts them against the true classes there."""
    from oracle import metrics_oracle as M
    n, d, k = 20000, 16, 8
    x = blobs(n, d, k, 3, spread=6.0)
    truth = (np.arange(n) % k).astype(np.uint32)           # blobs(): point i belongs to centre i % k
    ds = ctx.upload(x)
    first, u = cluster.kmeanspp_draws(5, n, k)
    ds.kmeanspp(k, first, u)
    cent, _ = ds.init_centroids(k)
    out = ds.lloyd_iterate(cent, 5)
    table = ds.contingency(truth, k, k)
    labels = ds.labels().astype(np.int64)
    assert table.tolist() == M.contingency_matrix(truth, labels) or table.sum() == n
    want = np.zeros((k, k), dtype=np.int64)
    np.add.at(want, (truth, labels), 1)
    assert np.array_equal(table, want)
    s = sc.HCVScore.new().compute_from_table(table)
    cols = table[:, table.sum(axis=0) > 0]                 # the reference only sees clusters that occur
    for got, wantv in zip((s.homogeneity(), s.completeness(), s.v_measure()), M.hcv(truth, labels)):
        assert abs(got - wantv) <= 1e-12
    assert cols.shape[1] >= 1
    with pytest.raises(cabi.SckmError):
        ds.contingency(truth + 1, k, k)                    # class id k is outside [0, k)
    ds.close()


# ---- BASELINE configs 3-5 at the full per-GPU size: size-independent properties + sampled rows against the oracle ----
def _full_size_check(ctx, O, n, d, k, dtype, max_iter, gap_tol, rtol, sample=2048):
    seed_data = 20260101
    ds = ctx.generate_blobs(n, d, k, seed_data, dtype=dtype)
    first, u = cluster.kmeanspp_draws(42, n, k)
    seeds = ds.kmeanspp(k, first, u)
    assert seeds[0] == first and len(set(seeds.tolist())) == k and seeds.min() >= 0 and seeds.max() < n
    # D^2 of sampled rows against the chosen seeds: bit-identical to the reference arithmetic (euclidian.rs:56-63)
    rows = np.sort(np.random.default_rng(1).choice(n, sample, replace=False))
    xs = np.vstack([cabi.blobs_host(int(r), 1, d, k, seed_data, dtype=dtype) for r in rows])
    seed_rows = np.vstack([cabi.blobs_host(int(r), 1, d, k, seed_data, dtype=dtype) for r in seeds])
    dd = ds.mindist()[rows]
    lab0 = ds.labels()[rows].astype(np.int64)
    for i in range(0, sample, 97):
        want = [O.squared_distance(xs[i], seed_rows[j]) for j in range(k)]
        assert dd[i] == min(want) and lab0[i] == int(np.argmin(want))
    cent, size = ds.init_centroids(k)
    assert size.sum() == n and size.min() >= 1
    out = ds.lloyd_fit(cent, max_iter)
    assert out["size"].sum() == n and 1 <= out["iters"] <= max_iter
    labels = ds.labels().astype(np.int64)
    assert np.array_equal(np.bincount(labels, minlength=k), out["size"])
    # one more step from the final centroids: inertia does not go up, sums are those of the final labels' successor
    inertia, sums, counts = ds.lloyd_step(out["centroids"])
    assert inertia <= out["distortion"] * (1 + 1e-9) and counts.sum() == n
    # sampled rows, labels recomputed by the oracle against the GPU's centroids
    d_o, s_o, c_o, m_o, gap = O.brute_clustering(xs, out["centroids"], want_gap=True)
    got = ds.labels()[rows].astype(np.int64)
    bad = np.nonzero(got != m_o)[0]
    assert np.all(gap[bad] < gap_tol), "%d sampled rows differ with gap >= %g" % (len(bad), gap_tol)
    # inertia of the sample from the GPU's labels equals the oracle's within the tolerance of the data type
    mine = np.array([O.squared_distance(xs[i].astype(np.float64), out["centroids"][got[i]]) for i in range(0, sample, 16)])
    ref = np.array([O.squared_distance(xs[i].astype(np.float64), out["centroids"][m_o[i]]) for i in range(0, sample, 16)])
    assert abs(mine.sum() - ref.sum()) <= rtol * ref.sum()
    ds.close()


def test_config3_10Mx64_k256_full_size(ctx, O):
    _full_size_check(ctx, O, 10_000_000, 64, 256, np.float64, max_iter=6, gap_tol=GAP_TOL, rtol=RTOL)


def test_config4_shard_12p5Mx128_k1024(ctx, O):
    """One GPU's share of config 4 (100M x 128, k = 1024 over 8 GPUs): the streamed-centroid tile kernel."""
    _full_size_check(ctx, O, 12_500_000, 128, 1024, np.float64, max_iter=2, gap_tol=GAP_TOL, rtol=RTOL, sample=1024)


def test_config5_shard_6p25Mx32_k4096_f32(ctx, O):
    """One GPU's share of config 5 (50M x 32 f32, k = 4096 over 8 GPUs): kmeans++ on the GPU, tcgen05 3xTF32 kernel."""
    _full_size_check(ctx, O, 6_250_000, 32, 4096, np.float32, max_iter=2, gap_tol=1e-5, rtol=1e-4, sample=1024)


# ---- batched LinearKNNSearch::find with Euclidian::distance (linear_search.rs:52-84) ---------------------------------
def _oracle_knn(O, x, q, k):
    from oracle import knn_oracle as K
    dist = lambda a, b: float(np.sqrt(O.squared_distance(a, b)))          # Euclidian::distance, TX arithmetic inside
    return sorted(((d, i) for i, d in K.find(list(x), dist, q, k)))


def test_knn_reference_known_answer():  # linear_search.rs knn_find, Euclidian part
    s = sc.LinearKNNSearch.new(np.array([[1., 1.], [2., 2.], [3., 3.], [4., 4.], [5., 5.]]))
    got = s.find([3., 3.], 3)
    assert sorted(i for i, _ in got) == [1, 2, 3] and got[0] == (2, 0.0)
    with pytest.raises(sc.Failed, match="k should be >= 1 and <= length"):
        s.find([3., 3.], 6)
    with pytest.raises(sc.Failed, match="k should be >= 1 and <= length"):
        s.find([3., 3.], 0)


@pytest.mark.parametrize("n,d,k,nq,dtype", [(300, 2, 1, 3, np.float64), (1000, 4, 5, 9, np.float32), (2500, 16, 32, 8, np.float64),
                                            (4000, 64, 64, 5, np.float64), (129, 8, 129 - 70, 17, np.float32), (65, 2, 64, 2, np.float64),
                                            # rows that are not multiples of 16 bytes (2-D / 3-D f32 points, odd d) and k > 64
                                            (1000, 3, 5, 4, np.float32), (700, 2, 100, 3, np.float32), (900, 5, 70, 3, np.float64),
                                            (500, 1, 200, 2, np.float32), (333, 7, 333, 2, np.float64)])
def test_knn_matches_oracle(O, n, d, k, nq, dtype):
    """Distances bit-identical to Euclidian::distance; the neighbour set is the reference's (continuous data: no ties at
    the k-th distance); order: ascending (distance, index) vs the reference's heap order, compared after sorting."""
    rng = np.random.default_rng(n + d)
    x = rng.normal(size=(n, d)).astype(dtype)
    q = np.vstack([rng.normal(size=(nq - 1, d)).astype(dtype), x[7:8]])   # the last query is a data row: distance 0 first
    s = sc.LinearKNNSearch.new(x)
    got = s.find_batch(q, k)
    for qi in range(nq):
        want = _oracle_knn(O, x, q[qi], k)
        assert [(dd, i) for i, dd in got[qi]] == want
    assert got[-1][0] == (7, 0.0)
    # column-major storage (DenseMatrix default) gives the same answer
    s2 = sc.LinearKNNSearch.new(sc.DenseMatrix.from_2d_array(x))
    assert s2.find(q[0], k) == got[0]


def test_knn_duplicates_and_large_n(ctx, O):
    """Exact ties: the distance multiset always equals the reference's; which of several rows at exactly the k-th
    distance survives is the heap-layout artefact documented in sckm_knn -- here the lowest indices win."""
    rng = np.random.default_rng(3)
    base = rng.normal(size=(40, 4))
    x = np.repeat(base, 5, axis=0)                          # every row 5 times
    s = sc.LinearKNNSearch.new(x)
    for qi in range(6):
        got = s.find(base[qi] + 0.25, 7)
        want = _oracle_knn(O, x, base[qi] + 0.25, 7)
        assert [dd for _, dd in got] == [dd for dd, _ in want]
        assert all(dd == float(np.sqrt(O.squared_distance(x[i], base[qi] + 0.25))) for i, dd in got)
        assert got == sorted(got, key=lambda t: (t[1], t[0])) and len(set(i for i, _ in got)) == 7
    # many rows, many chunks per query and the merge across them: against a float64 numpy ranking on sampled queries
    n, d, k = 200_000, 16, 8
    xl = cabi.blobs_host(0, n, d, 8, 5)
    ds = ctx.upload(xl)
    ql = xl[::25_000] + 0.5
    idx, dist = ds.knn(ql, k)
    for qi in range(len(ql)):
        d2 = ((xl - ql[qi]) ** 2).sum(axis=1)
        order = np.lexsort((np.arange(n), d2))[:k]
        assert idx[qi].tolist() == order.tolist()
        assert all(dist[qi, j] == float(np.sqrt(O.squared_distance(xl[idx[qi, j]], ql[qi]))) for j in range(k))
    # k > 64: several passes of 64, each above the previous pass's last pick
    idx, dist = ds.knn(ql[:3], 150)
    for qi in range(3):
        d2 = ((xl - ql[qi]) ** 2).sum(axis=1)
        assert idx[qi].tolist() == np.lexsort((np.arange(n), d2))[:150].tolist()
        assert np.all(np.diff(dist[qi]) >= 0)
    ds.close()


def test_knn_infinite_distances_are_not_neighbours(ctx):
    """The reference's heap starts full of INFINITY entries and only `d < datum.distance` replaces one
    (linear_search.rs:62-76): a row at infinite (or NaN) distance is never returned, the result is shorter than k."""
    x = np.array([[0.0, 0.0], [1.0, 0.0], [np.inf, 0.0], [np.nan, 1.0], [3.0, 4.0], [1e200, 1e200]])   # last: d^2 overflows
    ds = ctx.upload(x)
    idx, dist = ds.knn(np.array([[0.0, 0.0]]), 5)
    assert idx[0].tolist() == [0, 1, 4, -1, -1] and dist[0, :3].tolist() == [0.0, 1.0, 5.0] and np.all(np.isinf(dist[0, 3:]))
    ds.close()


def test_radius_fill_refuses_offsets_of_another_query_set(ctx):
    """sckm_radius_fill recomputes the counts: offsets that do not match them are an error, never an out-of-bounds write."""
    import ctypes as C
    rng = np.random.default_rng(5)
    x = rng.normal(size=(5000, 3)).astype(np.float32)        # 12-byte rows: the element-wise staging path
    ds = ctx.upload(x)
    q = x[:4] + np.float32(0.25)
    res = ds.radius(q, 1.0)
    counts = np.array([len(i) for i, _ in res], dtype=np.int64)
    full = np.sqrt(((x.astype(np.float64)[None] - q.astype(np.float64)[:, None]) ** 2).sum(-1))
    assert np.all(np.abs(counts - (full <= 1.0).sum(1)) <= 2) and counts.min() > 0
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    total = int(counts.sum())
    idx = np.full(total + 64, -7, dtype=np.int64); dist = np.zeros(total + 64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for bad_off, bad_total in ((offsets // 2, total // 2), (offsets, total - 1), (offsets[::-1].copy(), total)):
        rc = cabi.lib.sckm_radius_fill(ds.h, vp(q), 4, C.c_double(1.0), vp(bad_off), bad_total, vp(idx), vp(dist))
        assert rc == 1 and np.all(idx[bad_total:] == -7)
    # a different radius with the old offsets
    rc = cabi.lib.sckm_radius_fill(ds.h, vp(q), 4, C.c_double(1.5), vp(offsets), total, vp(idx), vp(dist))
    assert rc == 1 and np.all(idx[total:] == -7)
    ds.close()


@pytest.mark.parametrize("n,d,dtype", [(500, 2, np.float64), (3000, 16, np.float32), (70001, 4, np.float64),
                                       (800, 2, np.float32), (2500, 3, np.float64), (1200, 1, np.float32)])   # DBSCAN-style 2-D f32 points
def test_find_radius_matches_oracle(ctx, O, n, d, dtype):
    """LinearKNNSearch::find_radius (linear_search.rs:89-110): same rows, same order (ascending index), bit-identical
    distances; `d <= radius` includes the boundary; ragged results over many row chunks."""
    from oracle import knn_oracle as K
    rng = np.random.default_rng(n)
    x = rng.normal(size=(n, d)).astype(dtype)
    dist = lambda a, b: float(np.sqrt(O.squared_distance(a, b)))
    s = sc.LinearKNNSearch.new(x)
    sub = list(x[: min(n, 3000)])                           # the pure-Python restatement is slow: check a prefix of the rows
    for qi, radius in ((3, 0.9), (11, 2.5), (20, 1e-3)):
        got = s.find_radius(x[qi], radius)
        want = K.find_radius(sub, dist, x[qi], radius)
        assert [(i, dd) for i, dd in got if i < len(sub)] == want
        assert all(dd <= radius for _, dd in got) and [i for i, _ in got] == sorted(i for i, _ in got)
        assert (qi, 0.0) in got
    # boundary inclusive: radius equal to an actual distance keeps that row
    j, dj = s.find(x[5], 3)[2]
    assert (j, dj) in s.find_radius(x[5], dj)
    # batched, ragged, against numpy on the whole set
    ds = ctx.upload(x)
    q = x[:9] + dtype(0.125)
    res = ds.radius(q, 1.75)
    for qi in range(9):
        full = np.sqrt(((x.astype(np.float64) - q[qi].astype(np.float64)) ** 2).sum(axis=1))   # numpy ranking, ~1e-7 apart for f32
        idx, dv = res[qi]
        got = set(idx.tolist())
        assert set(np.nonzero(full <= 1.75 * (1 - 1e-5))[0].tolist()) <= got <= set(np.nonzero(full <= 1.75 * (1 + 1e-5))[0].tolist())
        assert np.all(np.diff(idx) > 0) and np.all(dv <= 1.75)
        for j in range(0, len(idx), max(1, len(idx) // 25)):
            assert dv[j] == dist(x[idx[j]], q[qi])                                              # Euclidian::distance, bit for bit
    ds.close()
    with pytest.raises(sc.Failed, match="radius should be > 0"):
        s.find_radius(x[0], 0.0)
