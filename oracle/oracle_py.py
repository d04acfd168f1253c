"""ctypes loader for oracle/liboracle_kmeans.so -- TEST INFRASTRUCTURE ONLY.

Thin numpy-facing wrappers over the C restatement in kmeans_oracle.cpp (which cites the
reference file:line for every function).  Not imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SEED_MODE_PCG = 0       # rand_core 0.6 default seed_from_u64 (PCG32 fill)  -- default
SEED_MODE_SPLITMIX = 1  # xoshiro's own seed_from_u64 (SplitMix64)
SEED_MODE_STDRNG = 2    # `std_rand` builds: StdRng = ChaCha12Rng, key = PCG32 fill of the seed (rand_custom.rs:1-4)


def build(force=False):
    so = os.path.join(_HERE, "liboracle_kmeans.so")
    src = os.path.join(_HERE, "kmeans_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_kmeans.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u64p = C.POINTER(C.c_uint64)
        L.orc_rng_seed.argtypes = [C.c_uint64, C.c_int, u64p]
        L.orc_rng_next_u64.argtypes = [u64p]; L.orc_rng_next_u64.restype = C.c_uint64
        L.orc_rng_gen_f64.argtypes = [u64p]; L.orc_rng_gen_f64.restype = C.c_double
        L.orc_rng_gen_range.argtypes = [u64p, C.c_uint64]; L.orc_rng_gen_range.restype = C.c_uint64
        L.orc_draws.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_size_t, C.c_void_p, C.c_void_p]; L.orc_draws.restype = None
        L.orc_rand_next_u64.argtypes = [C.c_uint64, C.c_int, C.c_size_t, C.c_void_p]; L.orc_rand_next_u64.restype = None
        L.orc_chacha_block.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]; L.orc_chacha_block.restype = None
        L.orc_pcg32_fill.argtypes = [C.c_uint64, C.c_void_p]; L.orc_pcg32_fill.restype = None
        for nm in ("f64", "f32", "i32"):
            f = getattr(L, "orc_squared_distance_" + nm)
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]; f.restype = C.c_double
        L.orc_bbd_new.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]; L.orc_bbd_new.restype = C.c_void_p
        L.orc_bbd_free.argtypes = [C.c_void_p]
        L.orc_bbd_num_nodes.argtypes = [C.c_void_p]; L.orc_bbd_num_nodes.restype = C.c_size_t
        L.orc_bbd_clustering.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_bbd_clustering.restype = C.c_double
        L.orc_brute_clustering.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int64,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_brute_clustering.restype = C.c_double
        L.orc_kmeanspp.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_fit.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64,
                              C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_fit.restype = C.c_int
        L.orc_predict.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dtype_code(x):
    if x.dtype == np.float32:
        return 0
    if x.dtype == np.float64:
        return 1
    raise TypeError("oracle supports f32/f64 data, got %s" % x.dtype)


class Rng:
    """rand 0.8.5 SmallRng (xoshiro256++) restatement; mode selects the seed_from_u64 candidate."""

    def __init__(self, seed, mode=SEED_MODE_PCG, state=None):
        self.state = (C.c_uint64 * 4)()
        if state is not None:
            for i in range(4):
                self.state[i] = state[i]
        else:
            lib().orc_rng_seed(seed, mode, self.state)

    def next_u64(self):
        return lib().orc_rng_next_u64(self.state)

    def gen_f64(self):
        return lib().orc_rng_gen_f64(self.state)

    def gen_range(self, n):
        return lib().orc_rng_gen_range(self.state, n)


def draws(seed, n, k, mode=SEED_MODE_PCG):
    """RNG draw sequence of kmeans_plus_plus (kmeans.rs:359,385): (first_index, uniforms[k-1]) for either RngImpl."""
    first = C.c_uint64(0); u = np.zeros(max(k - 1, 0))
    lib().orc_draws(seed, mode, n, k, C.addressof(first), _p(u))
    return first.value, u


def rand_next_u64(seed, mode, count):
    out = np.zeros(count, dtype=np.uint64)
    lib().orc_rand_next_u64(seed, mode, count, _p(out))
    return out


def chacha_block(key_words, counter, rounds, stream=0):
    key = np.ascontiguousarray(key_words, dtype=np.uint32); out = np.zeros(16, dtype=np.uint32)
    lib().orc_chacha_block(_p(key), counter, stream, rounds, _p(out))
    return out


def pcg32_fill(seed):
    out = np.zeros(32, dtype=np.uint8)
    lib().orc_pcg32_fill(seed, _p(out))
    return out


def squared_distance(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape
    nm = {np.dtype("float64"): "f64", np.dtype("float32"): "f32", np.dtype("int32"): "i32"}[a.dtype]
    return getattr(lib(), "orc_squared_distance_" + nm)(_p(a), _p(b), a.size)


class BBDTree:
    """BBDTree::new + clustering (bbd_tree.rs:42-163)."""

    def __init__(self, x):
        self.x = np.ascontiguousarray(x, dtype=np.float64)  # to_f64 view, as the reference reads it
        self.n, self.d = self.x.shape
        self.h = lib().orc_bbd_new(_p(self.x), self.n, self.d)

    def num_nodes(self):
        return lib().orc_bbd_num_nodes(self.h)

    def clustering(self, centroids):
        c = np.ascontiguousarray(centroids, dtype=np.float64)
        k = c.shape[0]
        sums = np.zeros((k, self.d)); counts = np.zeros(k, dtype=np.int64); mem = np.zeros(self.n, dtype=np.int64)
        dist = lib().orc_bbd_clustering(self.h, _p(c), k, _p(sums), _p(counts), _p(mem))
        return dist, sums, counts, mem

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_bbd_free(self.h); self.h = None


def brute_clustering(x, centroids, want_gap=False):
    x = np.ascontiguousarray(x, dtype=np.float64); c = np.ascontiguousarray(centroids, dtype=np.float64)
    n, d = x.shape; k = c.shape[0]
    sums = np.zeros((k, d)); counts = np.zeros(k, dtype=np.int64); mem = np.zeros(n, dtype=np.int64)
    gap = np.zeros(n) if want_gap else None
    dist = lib().orc_brute_clustering(_p(x), n, d, _p(c), k, _p(sums), _p(counts), _p(mem), _p(gap))
    return (dist, sums, counts, mem, gap) if want_gap else (dist, sums, counts, mem)


def kmeanspp(x, k, seed=0, seed_mode=SEED_MODE_PCG, inject=None):
    x = np.ascontiguousarray(x); n, d = x.shape
    y = np.zeros(n, dtype=np.int64); idx = np.zeros(k, dtype=np.int64); dd = np.zeros(n)
    inj = None if inject is None else np.ascontiguousarray(inject, dtype=np.int64)
    lib().orc_kmeanspp(_p(x), _dtype_code(x), n, d, k, seed, seed_mode, _p(inj), _p(y), _p(idx), _p(dd))
    return y, idx, dd


class FitResult:
    pass


def fit(x, k, max_iter=100, seed=0, seed_mode=SEED_MODE_PCG, inject=None, use_tree=True):
    """KMeans::fit restatement.  Returns FitResult or raises ValueError with the reference message."""
    x = np.ascontiguousarray(x); n, d = x.shape
    r = FitResult()
    r.y = np.zeros(n, dtype=np.int64); r.size = np.zeros(max(k, 1), dtype=np.int64)
    r.centroids = np.zeros((max(k, 1), d)); dist = C.c_double(0); iters = C.c_int64(0)
    r.seed_idx = np.zeros(max(k, 1), dtype=np.int64); times = np.zeros(3)
    inj = None if inject is None else np.ascontiguousarray(inject, dtype=np.int64)
    rc = lib().orc_fit(_p(x), _dtype_code(x), n, d, k, max_iter, seed, seed_mode, _p(inj), 1 if use_tree else 0,
                       _p(r.y), _p(r.size), _p(r.centroids), C.addressof(dist), C.addressof(iters),
                       _p(r.seed_idx), _p(times))
    if rc == 1:
        raise ValueError("Fit failed: invalid number of clusters: %d" % k)
    if rc == 2:
        raise ValueError("Fit failed: invalid maximum number of iterations: %d" % max_iter)
    r.distortion = dist.value; r.iters = iters.value
    r.t_tree, r.t_kmeanspp, r.t_lloyd = times
    return r


def predict(x, centroids):
    x = np.ascontiguousarray(x); n, d = x.shape
    c = np.ascontiguousarray(centroids, dtype=np.float64)
    out = np.zeros(n, dtype=np.int64)
    lib().orc_predict(_p(x), _dtype_code(x), n, d, _p(c), c.shape[0], _p(out))
    return out
