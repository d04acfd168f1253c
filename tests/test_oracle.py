"""CPU: the oracle restatement against every known-answer the reference's own tests hold for the
hot path and against the committed regression fixtures."""
import numpy as np
import pytest


def test_xoshiro256pp_upstream_vector(O, kat):
    r = O.Rng(0, state=[1, 2, 3, 4])
    assert [r.next_u64() for _ in range(10)] == kat["xoshiro256pp_state_1234"]


def test_splitmix_seeding_candidate(O, kat):
    r = O.Rng(0, mode=O.SEED_MODE_SPLITMIX)
    assert list(r.state) == kat["splitmix_seed0_state"]
    assert [r.next_u64() for _ in range(4)] == kat["splitmix_seed0_out"]


def test_pcg_seeding_is_deterministic_and_distinct(O):
    a, b = O.Rng(42), O.Rng(42)
    assert [a.next_u64() for _ in range(8)] == [b.next_u64() for _ in range(8)]
    assert list(O.Rng(42).state) != list(O.Rng(42, mode=O.SEED_MODE_SPLITMIX).state)
    assert list(O.Rng(0).state) != [0, 0, 0, 0]


def test_gen_f64_and_gen_range_bounds(O):
    r = O.Rng(7)
    us = [r.gen_f64() for _ in range(2000)]
    assert 0.0 <= min(us) and max(us) < 1.0 and 0.4 < np.mean(us) < 0.6
    for n in (1, 2, 3, 150, 10**7, 2**40 + 17):
        vals = [r.gen_range(n) for _ in range(200)]
        assert 0 <= min(vals) and max(vals) < n
    # Standard f64 = top 53 bits * 2^-53
    r1, r2 = O.Rng(9), O.Rng(9)
    assert r1.gen_f64() == (r2.next_u64() >> 11) * 2.0 ** -53


def test_squared_distance_kat(O, kat):  # euclidian.rs:84-91 (i32 inputs -> f64)
    k = kat["squared_distance"]
    a = np.array(k["a"], dtype=np.int32); b = np.array(k["b"], dtype=np.int32)
    assert abs(O.squared_distance(a, b) ** 0.5 - k["l2"]) < k["tol"]
    assert O.squared_distance(a.astype(np.float64), b.astype(np.float64)) == 27.0


def test_squared_distance_f32_rounds_in_f32(O):
    a = np.array([0.1, 0.7, 1e-3], dtype=np.float32); b = np.array([0.3, -0.2, 5.0], dtype=np.float32)
    want = 0.0
    for x, y in zip(a, b):
        r = np.float32(x - y)
        want += float(np.float32(r * r))
    assert O.squared_distance(a, b) == want


def test_bbdtree_iris_golden(O, kat, iris20):  # bbd_tree.rs:324-364
    g = kat["bbdtree_iris"]
    tree = O.BBDTree(iris20)
    dist, sums, counts, mem = tree.clustering(g["centroids"])
    assert abs(dist - g["cost"]) < g["cost_tol"]
    assert abs(sums[0][0] - g["sums_0_0"]) < g["sums_tol"]
    assert abs(sums[1][3] - g["sums_1_3"]) < g["sums_tol"]
    assert mem[17] == g["membership_17"]
    assert counts.sum() == 20


def test_tree_equals_dense_step(O, iris_f32):
    x = iris_f32[0].astype(np.float64)
    rng = np.random.default_rng(1)
    tree = O.BBDTree(x)
    for k in (2, 3, 7):
        c = x[rng.choice(150, k, replace=False)] + 0.01
        dt, st, ct, mt = tree.clustering(c)
        db, sb, cb, mb, gap = O.brute_clustering(x, c, want_gap=True)
        bad = mt != mb
        assert np.all(gap[bad] < 1e-12)
        assert np.array_equal(ct, cb) or bad.any()
        np.testing.assert_allclose(st, sb, rtol=1e-12, atol=1e-12)
        assert abs(dt - db) <= 1e-11 * abs(db)


def test_fit_predict_self_consistency(O, iris20):  # kmeans.rs:473-505, any seed
    for seed in (0, 1, 42, 12345):
        for mode in (0, 1):
            r = O.fit(iris20, 2, 100, seed, mode)
            assert np.array_equal(O.predict(iris20, r.centroids), r.y)
            assert r.size.sum() == 20 and r.iters >= 1


def test_invalid_parameters(O, kat):  # kmeans.rs:426-443
    x = np.array([[1, 2, 3], [4, 5, 6]], dtype=np.float64)
    with pytest.raises(ValueError):
        O.fit(x, 0)
    with pytest.raises(ValueError) as e:
        O.fit(x, 1)
    assert str(e.value) == kat["invalid_k_message"]
    with pytest.raises(ValueError) as e:
        O.fit(x, 2, max_iter=0)
    assert str(e.value) == "Fit failed: invalid maximum number of iterations: 0"


def test_regression_fixtures(O, oracle_fits, iris_f32, iris20):
    data = {"iris_f64_k3_seed42": iris_f32[0].astype(np.float64), "iris_f32_k3_seed42": iris_f32[0],
            "iris20_f64_k2_seedNone": iris20}
    for name, g in oracle_fits.items():
        base, mode = name.rsplit("_mode", 1)
        r = O.fit(data[base], g["k"], 100, g["seed"], int(mode))
        assert r.y.tolist() == g["y"] and r.size.tolist() == g["size"] and r.iters == g["iters"]
        assert r.seed_idx.tolist() == g["seed_idx"]
        assert r.distortion == g["distortion"]
        np.testing.assert_array_equal(r.centroids, np.array(g["centroids"]))


def test_injected_seed_rows_bypass_rng(O, iris_f32):
    x = iris_f32[0].astype(np.float64)
    r = O.fit(x, 3, 100, 42)
    r2 = O.fit(x, 3, 100, 999, inject=r.seed_idx)
    assert np.array_equal(r.y, r2.y) and r.distortion == r2.distortion


def test_duplicate_rows_leaf_rule(O):
    # duplicate points form a leaf whose sum is first point x count (bbd_tree.rs:234-246)
    x = np.array([[1.0, 1.0]] * 5 + [[4.0, 4.0]] * 3 + [[9.0, 0.5]])
    tree = O.BBDTree(x)
    dist, sums, counts, mem = tree.clustering([[1.0, 1.1], [4.2, 4.0], [9.0, 0.0]])
    assert counts.tolist() == [5, 3, 1] and mem.tolist() == [0] * 5 + [1] * 3 + [2]
    np.testing.assert_allclose(sums, [[5, 5], [12, 12], [9, 0.5]])
    db = O.brute_clustering(x, [[1.0, 1.1], [4.2, 4.0], [9.0, 0.0]])[0]
    assert abs(dist - db) < 1e-12


def test_kmeanspp_labels_are_nearest_seed(O, iris_f32):
    x = iris_f32[0]
    y, idx, dd = O.kmeanspp(x, 4, seed=3)
    seeds = x[idx].astype(np.float32)
    for i in range(150):
        ds = [O.squared_distance(x[i], s) for s in seeds]
        assert dd[i] == min(ds)
        assert y[i] == int(np.argmin(ds))  # strict <: first minimum wins


# ---- cluster-quality scores: the reference's own known answers pin the restatement ----------------------------
def test_metrics_oracle_against_reference_kats():
    from oracle import metrics_oracle as M
    v1 = [0, 0, 1, 1, 2, 0, 4]
    v2 = [1, 0, 0, 0, 0, 1, 0]
    assert M.contingency_matrix(v1, v2) == [[1, 2], [2, 0], [1, 0], [1, 0]]   # cluster_helpers.rs:117-125
    assert abs(1.2770 - M.entropy(v1)) < 1e-4                                  # cluster_helpers.rs:131-135
    assert abs(0.3254 - M.mutual_info_score(M.contingency_matrix(v1, v2))) < 1e-4   # cluster_helpers.rs:141-147
    h, c, v = M.hcv(v1, v2)                                                    # cluster_hcv.rs:94-104
    assert abs(0.2548 - h) < 1e-4 and abs(0.5440 - c) < 1e-4 and abs(0.3471 - v) < 1e-4


def test_metrics_oracle_against_sklearn():
    """Second opinion: scikit-learn computes the same three scores (natural logs cancel in the ratios)."""
    from oracle import metrics_oracle as M
    from sklearn.metrics import homogeneity_completeness_v_measure, mutual_info_score
    rng = np.random.default_rng(5)
    for n, na, nb in ((50, 3, 4), (1000, 7, 5), (3000, 20, 31)):
        a = rng.integers(0, na, n) * 3 - 4
        b = (a // 3 + rng.integers(0, nb, n) * (rng.random(n) < 0.3)) % nb
        h, c, v = M.hcv(a, b)
        hs, cs, vs = homogeneity_completeness_v_measure(a, b)
        assert abs(h - hs) < 1e-10 and abs(c - cs) < 1e-10 and abs(v - vs) < 1e-10
        assert abs(M.mutual_info_score(M.contingency_matrix(a, b)) - mutual_info_score(a, b)) < 1e-10


# ---- brute-force neighbour search: the reference's own known answers pin the restatement ------------------------
def test_knn_oracle_against_reference_kats():
    from oracle import knn_oracle as K
    h = K.HeapSelection(3)                                   # heap_select.rs test_add
    for v in (-5, 333, 13, 10, 2, 0, 40, 30):
        h.add(v)
    assert h.n == 8 and h.get() == [2, 0, -5]
    h = K.HeapSelection(3)                                   # test_add1
    for v in (float("inf"), -5.0, 4.0, -1.0, 2.0, 1.0, 0.0):
        h.add(v)
    assert h.n == 7 and h.get() == [0.0, -1.0, -5.0]
    h = K.HeapSelection(3)                                   # test_add2
    for v in (float("inf"), 0.0, 8.4852, 5.6568, 2.8284):
        h.add(v)
    assert h.get() == [5.6568, 2.8284, 0.0]
    h = K.HeapSelection(3)                                   # test_add_ordered
    for v in (1.0, 2.0, 3.0, 4.0, 5.0, 6.0):
        h.add(v)
    assert h.get() == [3.0, 2.0, 1.0]
    simple = lambda a, b: float(abs(a - b))                  # linear_search.rs knn_find
    data1 = list(range(1, 11))
    assert sorted(i for i, _ in K.find(data1, simple, 2, 3)) == [0, 1, 2]
    assert sorted(data1[i] for i, _ in K.find_radius(data1, simple, 5, 3.0)) == [2, 3, 4, 5, 6, 7, 8]
    import math
    eu = lambda a, b: math.sqrt(sum((x - y) * (x - y) for x, y in zip(a, b)))
    data2 = [[1., 1.], [2., 2.], [3., 3.], [4., 4.], [5., 5.]]
    assert sorted(i for i, _ in K.find(data2, eu, [3., 3.], 3)) == [1, 2, 3]
    with pytest.raises(ValueError):
        K.find(data1, simple, 2, 11)


# ---- RNG of the `std_rand` build and the seeding path: published known answers ----------------------------------------
def _xsh_rr(state):
    xs = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
    rot = state >> 59
    return ((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF


PCG_MUL, MASK64 = 6364136223846793005, (1 << 64) - 1


def test_pcg32_output_function_reproduces_the_pcg_reference_vector():
    """PCG32 (XSH-RR 64/32) with the reference seeding (initstate 42, initseq 54) -- the demo vector of the PCG
    distribution: pins the multiplier and the output permutation that rand_core's seed_from_u64 is built from."""
    state, inc = 0, ((54 << 1) | 1)
    out = []

    def step():
        nonlocal state
        old = state
        state = (old * PCG_MUL + inc) & MASK64
        return _xsh_rr(old)
    step(); state = (state + 42) & MASK64; step()
    out = [step() for _ in range(6)]
    assert out == [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]


def test_seed_from_u64_fill_matches_the_python_restatement(O):
    """rand_core 0.6 SeedableRng::seed_from_u64: advance first (MUL, INC = 11634580027462260723), XSH-RR of the NEW state,
    little-endian words -- restated here in Python on top of the output function pinned above, against the oracle (C++)
    and the host mirror (C++), for both RngImpl kinds."""
    import smartcore_b200.cluster as cl
    for seed in (0, 1, 42, 2**63 + 12345, 2**64 - 1):
        st, words = seed, []
        for _ in range(8):
            st = (st * PCG_MUL + 11634580027462260723) & MASK64
            words.append(_xsh_rr(st))
        assert O.pcg32_fill(seed).view("<u4").tolist() == words
        # StdRng: the words are the ChaCha12 key; first outputs = block 0 of that key
        blk = O.chacha_block(np.array(words, dtype=np.uint32), 0, 12)
        want = [int(blk[2 * i]) | (int(blk[2 * i + 1]) << 32) for i in range(8)]
        assert O.rand_next_u64(seed, O.SEED_MODE_STDRNG, 8).tolist() == want
        assert cl.rng_next_u64(seed, 8, std_rand=True).tolist() == want
        # SmallRng: the words are the xoshiro256++ state (LE u64s); mirror == oracle
        assert cl.rng_next_u64(seed, 8, std_rand=False).tolist() == O.rand_next_u64(seed, O.SEED_MODE_PCG, 8).tolist()


def test_chacha_block_published_vectors(O):
    """ChaCha with an all-zero 256-bit key and IV, block 0 (TC1 of the published ChaCha test vectors): 8, 12 (StdRng's
    variant in rand 0.8: ChaCha12Rng) and 20 rounds; oracle and host mirror."""
    import smartcore_b200.cluster as cl
    kat = {8: "3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e984ce172b9216f419f445367456d5619314a42a3da86b001387bfdb80e0cfe42",
           12: "9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be",
           20: "76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586"}
    zero = np.zeros(8, dtype=np.uint32)
    for rounds, hexv in kat.items():
        assert O.chacha_block(zero, 0, rounds).astype("<u4").tobytes().hex() == hexv
        assert cl.chacha_block(zero, 0, rounds).astype("<u4").tobytes().hex() == hexv
    # the block counter occupies words 12-13 (64 bit): block 1 differs from block 0 and is what the 9th next_u64 reads
    w = O.rand_next_u64(7, O.SEED_MODE_STDRNG, 9)
    key = O.pcg32_fill(7).view("<u4")
    b1 = O.chacha_block(key, 1, 12)
    assert int(w[8]) == int(b1[0]) | (int(b1[1]) << 32)


def test_std_rand_draw_sequence_host_mirror_equals_oracle(O):
    import smartcore_b200 as sc
    from smartcore_b200 import cluster as cl
    try:
        sc.set_std_rand(True)
        assert sc.get_std_rand()
        for seed, n, k in ((42, 150, 3), (0, 10_000_000, 256), (2**40 + 3, 77, 12)):
            f, u = cl.kmeanspp_draws(seed, n, k)
            fo, uo = O.draws(seed, n, k, O.SEED_MODE_STDRNG)
            assert f == fo and np.array_equal(u, uo) and 0 <= f < n and np.all((u >= 0) & (u < 1))
        sc.set_std_rand(False)
        f, u = cl.kmeanspp_draws(42, 150, 3)
        fo, uo = O.draws(42, 150, 3, O.SEED_MODE_PCG)
        assert f == fo and np.array_equal(u, uo)
        assert (f, u.tolist()) != tuple([O.draws(42, 150, 3, O.SEED_MODE_STDRNG)[0], O.draws(42, 150, 3, O.SEED_MODE_STDRNG)[1].tolist()])
    finally:
        sc.set_std_rand(False)


def test_numpy_blob_generator_equals_the_host_twin():
    """oracle/blobs_np.py (what the CPU arm of bench.py builds its input with) against sckm_blobs_fill_host, bit for bit."""
    from oracle import blobs_np
    from smartcore_b200 import cabi
    for (row0, n, d, k, seed, dt) in ((0, 3000, 64, 256, 20260101, np.float64), ((1 << 33) + 5, 500, 7, 3, 99, np.float32), (123, 70000, 5, 8, 1, np.float64)):
        assert np.array_equal(blobs_np.blobs(row0, n, d, k, seed, dtype=dt), cabi.blobs_host(row0, n, d, k, seed, dtype=dt))
