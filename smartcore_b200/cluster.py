"""Python face of the host mirror (host/smartcore_kmeans.hpp via libsmartcore_host.so).

Same names, argument meaning and error behaviour as smartcore's Rust API:
  KMeans.fit(x, parameters) / .predict(x)      src/cluster/kmeans.rs:254-352
  KMeansParameters {k, max_iter, seed}         src/cluster/kmeans.rs:109-146
  KMeansSearchParameters (grid iterator)       src/cluster/kmeans.rs:148-232
  DenseMatrix.new / from_2d_array              src/linalg/basic/matrix.rs:187-237
  Failed ("Fit failed: ...")                   src/error/mod.rs:109-128
The computation happens in libsmartcore_kmeans_cuda.so; this module holds no arithmetic.
"""
import ctypes as C
import os

import numpy as np

from . import cabi  # noqa: F401  (loads the CUDA library first; raises when it is missing)

_HOST_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libsmartcore_host.so")
if not os.path.exists(_HOST_PATH):
    raise ImportError("%s is missing: run __graft_entry__.build()" % _HOST_PATH)
_h = C.CDLL(_HOST_PATH)
_vp = C.c_void_p
_h.sch_kmeans_fit.argtypes = [C.c_int, _vp, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_int,
                              C.c_uint64, C.POINTER(_vp), C.c_char_p, C.c_size_t]
_h.sch_kmeans_predict.argtypes = [_vp, _vp, C.c_size_t, C.c_size_t, C.c_int, _vp, C.c_char_p, C.c_size_t]
_h.sch_kmeans_dims.argtypes = [_vp, _vp, _vp, _vp]; _h.sch_kmeans_dims.restype = None
_h.sch_kmeans_get.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp]; _h.sch_kmeans_get.restype = None
_h.sch_kmeans_free.argtypes = [_vp]; _h.sch_kmeans_free.restype = None
_h.sch_kmeans_serialize.argtypes = [_vp, C.c_int, _vp, C.c_size_t]; _h.sch_kmeans_serialize.restype = C.c_size_t
_h.sch_kmeans_deserialize.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_size_t, _vp, C.c_char_p, C.c_size_t]
_h.sch_kmeans_eq.argtypes = [_vp, _vp]
_h.sch_search_parameters.argtypes = [_vp, C.c_size_t, _vp, C.c_size_t, _vp, _vp, C.c_size_t, _vp, _vp, _vp, _vp, C.c_size_t]
_h.sch_search_parameters.restype = C.c_size_t
_h.sch_kmeanspp_draws.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_size_t, _vp, _vp]; _h.sch_kmeanspp_draws.restype = None
_h.sch_set_std_rand.argtypes = [C.c_int]; _h.sch_set_std_rand.restype = None
_h.sch_get_std_rand.restype = C.c_int
_h.sch_rng_next_u64.argtypes = [C.c_uint64, C.c_int, C.c_size_t, _vp]; _h.sch_rng_next_u64.restype = None
_h.sch_chacha_block.argtypes = [_vp, C.c_uint64, C.c_int, _vp]; _h.sch_chacha_block.restype = None
_h.sch_dense_get_f64.argtypes = [_vp, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t]
_h.sch_dense_get_f64.restype = C.c_double

_DT = {np.dtype("float32"): 0, np.dtype("float64"): 1, np.dtype("int32"): 2, np.dtype("int64"): 3}


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


class Failed(Exception):
    """smartcore::error::Failed; str() is the reference's Display form."""


class DenseMatrix:
    """Contiguous values + layout flag, like smartcore's DenseMatrix<T>."""

    def __init__(self, nrows, ncols, values, column_major=True):
        values = np.ascontiguousarray(values).reshape(-1)
        if nrows * ncols != values.size:
            raise Failed("Error in input, check parameters: The specified shape: (cols: %d, rows: %d) does not "
                         "align with data len: %d" % (ncols, nrows, values.size))
        if values.dtype not in _DT:
            raise TypeError("unsupported element type %s" % values.dtype)
        self.nrows, self.ncols, self.values, self.column_major = nrows, ncols, values, bool(column_major)

    new = classmethod(lambda cls, nrows, ncols, values, column_major: cls(nrows, ncols, values, column_major))

    @classmethod
    def from_2d_array(cls, rows, dtype=None):
        a = np.asarray(rows, dtype=dtype)
        if a.ndim != 2 or a.size == 0:
            raise Failed("Error in input, check parameters: The 2d vec provided is empty; cannot instantiate the matrix")
        return cls(a.shape[0], a.shape[1], np.ascontiguousarray(a.T).reshape(-1), True)  # column-major (matrix.rs:230-236)

    @classmethod
    def from_numpy(cls, a):
        """Row-major view of a C-contiguous (n, d) array: DenseMatrix::new(n, d, values, false)."""
        a = np.ascontiguousarray(a)
        return cls(a.shape[0], a.shape[1], a.reshape(-1), False)

    def shape(self):
        return (self.nrows, self.ncols)

    def get(self, pos):
        r, c = pos
        return self.values[c * self.nrows + r] if self.column_major else self.values[c + self.ncols * r]


class KMeansParameters:
    def __init__(self, k=2, max_iter=100, seed=None):
        self.k, self.max_iter, self.seed = k, max_iter, seed

    @classmethod
    def default(cls):
        return cls()

    def with_k(self, k):
        return KMeansParameters(k, self.max_iter, self.seed)

    def with_max_iter(self, max_iter):
        return KMeansParameters(self.k, max_iter, self.seed)


class KMeansSearchParameters:
    def __init__(self, k=None, max_iter=None, seed=None):
        d = KMeansParameters()
        self.k = [d.k] if k is None else list(k)
        self.max_iter = [d.max_iter] if max_iter is None else list(max_iter)
        self.seed = [d.seed] if seed is None else list(seed)

    def __iter__(self):
        k = np.array(self.k, dtype=np.uint64); m = np.array(self.max_iter, dtype=np.uint64)
        s = np.array([0 if v is None else v for v in self.seed], dtype=np.uint64)
        hs = np.array([0 if v is None else 1 for v in self.seed], dtype=np.int32)
        cap = len(k) * len(m) * len(s) + 4
        ok = np.zeros(cap, dtype=np.uint64); om = np.zeros(cap, dtype=np.uint64); os_ = np.zeros(cap, dtype=np.uint64)
        oh = np.zeros(cap, dtype=np.int32)
        n = _h.sch_search_parameters(_p(k), len(k), _p(m), len(m), _p(s), _p(hs), len(s), _p(ok), _p(om), _p(os_), _p(oh), cap)
        for i in range(n):
            yield KMeansParameters(int(ok[i]), int(om[i]), int(os_[i]) if oh[i] else None)


def set_std_rand(on):
    """Follow smartcore built with feature `std_rand` (RngImpl = StdRng = ChaCha12; forced by `datasets`) instead of the
    default-feature build (SmallRng = xoshiro256++): src/rand_custom.rs:1-4."""
    _h.sch_set_std_rand(1 if on else 0)


def get_std_rand():
    return bool(_h.sch_get_std_rand())


def rng_next_u64(seed, count, std_rand=False):
    out = np.zeros(count, dtype=np.uint64)
    _h.sch_rng_next_u64(seed, 1 if std_rand else 0, count, _p(out))
    return out


def chacha_block(key_words, counter, rounds):
    key = np.ascontiguousarray(key_words, dtype=np.uint32); out = np.zeros(16, dtype=np.uint32)
    _h.sch_chacha_block(_p(key), counter, rounds, _p(out))
    return out


def kmeanspp_draws(seed, n, k):
    """The host RNG sequence of kmeans_plus_plus: (first_index, uniforms[k-1])."""
    first = C.c_uint64(0); u = np.zeros(max(k - 1, 0))
    _h.sch_kmeanspp_draws(0 if seed is None else 1, 0 if seed is None else seed, n, k, C.addressof(first), _p(u))
    return first.value, u


class KMeans:
    """KMeans<TX, TY, DenseMatrix<TX>, Vec<TY>> fitted on the GPU."""

    def __init__(self, handle, dtype):
        self._h, self._dtype = handle, dtype
        k = C.c_size_t(0); n = C.c_size_t(0); d = C.c_size_t(0)
        _h.sch_kmeans_dims(handle, C.addressof(k), C.addressof(n), C.addressof(d))
        self.k = k.value
        self._y = np.zeros(n.value, dtype=np.int64); self.size = np.zeros(self.k, dtype=np.int64)
        self.centroids = np.zeros((self.k, d.value)); dist = C.c_double(0); it = C.c_int64(0)
        _h.sch_kmeans_get(handle, _p(self._y), _p(self.size), _p(self.centroids), C.addressof(dist), C.addressof(it))
        self._distortion, self._iterations = dist.value, it.value

    def __del__(self):
        if getattr(self, "_h", None):
            _h.sch_kmeans_free(self._h); self._h = None

    @classmethod
    def fit(cls, data, parameters=None):
        p = parameters or KMeansParameters()
        model = _vp(); err = C.create_string_buffer(1024)
        rc = _h.sch_kmeans_fit(_DT[data.values.dtype], _p(data.values), data.nrows, data.ncols,
                               1 if data.column_major else 0, p.k, p.max_iter, 0 if p.seed is None else 1,
                               0 if p.seed is None else p.seed, C.byref(model), err, len(err))
        if rc:
            raise Failed(err.value.decode())
        return cls(model, data.values.dtype)

    # ---- serde (kmeans.rs:70-83): images interchangeable with the reference's derive(Serialize, Deserialize) ----
    def _serialize(self, fmt):
        size = _h.sch_kmeans_serialize(self._h, fmt, None, 0)
        buf = C.create_string_buffer(size)
        _h.sch_kmeans_serialize(self._h, fmt, buf, size)
        return buf.raw[:size]

    @classmethod
    def _deserialize(cls, fmt, image, dtype):
        model = _vp(); err = C.create_string_buffer(1024)
        dtype = np.dtype(dtype)
        if _h.sch_kmeans_deserialize(_DT[dtype], fmt, image, len(image), C.byref(model), err, len(err)):
            raise Failed(err.value.decode())
        return cls(model, dtype)

    def to_json(self):
        """serde_json::to_string(&kmeans)"""
        return self._serialize(0).decode()

    def to_bincode(self):
        """bincode::serialize(&kmeans) (bincode 1.3 default options)"""
        return self._serialize(1)

    @classmethod
    def from_json(cls, text, dtype=np.float64):
        """serde_json::from_str::<KMeans<TX, ..>>(text); dtype = TX"""
        return cls._deserialize(0, text.encode() if isinstance(text, str) else text, dtype)

    @classmethod
    def from_bincode(cls, image, dtype=np.float64):
        return cls._deserialize(1, bytes(image), dtype)

    def __eq__(self, other):  # PartialEq (kmeans.rs:85-107)
        return isinstance(other, KMeans) and bool(_h.sch_kmeans_eq(self._h, other._h))

    __hash__ = None

    def predict(self, x, ty=np.int64):
        out = np.zeros(x.nrows, dtype=np.int64); err = C.create_string_buffer(1024)
        if x.values.dtype != self._dtype:
            raise Failed("Predict failed: element type differs from the fitted model")
        rc = _h.sch_kmeans_predict(self._h, _p(x.values), x.nrows, x.ncols, 1 if x.column_major else 0, _p(out), err, len(err))
        if rc:
            raise Failed(err.value.decode())
        return out.astype(ty)  # TY::from_usize
