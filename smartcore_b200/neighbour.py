"""smartcore::algorithm::neighbour::linear_search::LinearKNNSearch for row vectors under the Euclidian metric
(src/algorithm/neighbour/linear_search.rs:35-84, src/metrics/distance/euclidian.rs:51-76) over the C++ host mirror;
the search runs on the GPU (sckm_knn).  No CPU fallback."""
import ctypes as C

import numpy as np

from .cluster import DenseMatrix, Failed, _DT, _h, _p

_vp = C.c_void_p
_h.sch_knn_new.argtypes = [C.c_int, _vp, C.c_size_t, C.c_size_t, C.c_int, _vp, C.c_char_p, C.c_size_t]
_h.sch_knn_find.argtypes = [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, _vp, _vp, C.c_char_p, C.c_size_t]
_h.sch_knn_find_radius.argtypes = [_vp, _vp, C.c_size_t, C.c_double, _vp, _vp, C.c_size_t, C.c_char_p, C.c_size_t]
_h.sch_knn_find_radius.restype = C.c_longlong
_h.sch_knn_free.argtypes = [_vp]; _h.sch_knn_free.restype = None


class LinearKNNSearch:
    """LinearKNNSearch::new(data, Distances::euclidian()): data = DenseMatrix (or 2-D array) of f32 / f64 rows."""

    def __init__(self, data):
        if not isinstance(data, DenseMatrix):
            data = DenseMatrix.from_numpy(np.ascontiguousarray(data))
        if data.values.dtype not in (np.dtype("float32"), np.dtype("float64")):
            raise Failed("Error in input, check parameters: LinearKNNSearch on the GPU takes f32 or f64 rows")
        self._dtype, self._d, self._n = data.values.dtype, data.ncols, data.nrows
        h = _vp(); err = C.create_string_buffer(1024)
        if _h.sch_knn_new(_DT[self._dtype], _p(data.values), data.nrows, data.ncols, 1 if data.column_major else 0,
                          C.byref(h), err, len(err)):
            raise Failed(err.value.decode())
        self._h = h

    @classmethod
    def new(cls, data):
        return cls(data)

    def __del__(self):
        if getattr(self, "_h", None):
            _h.sch_knn_free(self._h); self._h = None

    def find_batch(self, queries, k):
        """[(index, distance), ...] per query row, ascending by (distance, index)."""
        q = np.ascontiguousarray(queries, dtype=self._dtype).reshape(-1, self._d)
        nq = q.shape[0]
        kk = max(int(k), 1)
        idx = np.zeros((nq, kk), dtype=np.int64); dist = np.zeros((nq, kk)); counts = np.zeros(nq, dtype=np.int64)
        err = C.create_string_buffer(1024)
        if _h.sch_knn_find(self._h, _p(q), nq, self._d, int(k), _p(idx), _p(dist), _p(counts), err, len(err)):
            raise Failed(err.value.decode())
        return [list(zip(idx[i, :counts[i]].tolist(), dist[i, :counts[i]].tolist())) for i in range(nq)]

    def find_radius(self, frm, radius):
        """LinearKNNSearch::find_radius(&from, radius) (linear_search.rs:89-110): [(index, distance)] in row order."""
        q = np.ascontiguousarray(frm, dtype=self._dtype).reshape(-1)
        if q.size != self._d:
            raise Failed("Find failed: query length differs from the data")
        cap = 1024
        while True:
            idx = np.zeros(cap, dtype=np.int64); dist = np.zeros(cap); err = C.create_string_buffer(1024)
            m = _h.sch_knn_find_radius(self._h, _p(q), self._d, float(radius), _p(idx), _p(dist), cap, err, len(err))
            if m < 0:
                raise Failed(err.value.decode())
            if m <= cap:
                return list(zip(idx[:m].tolist(), dist[:m].tolist()))
            cap = int(m)

    def find(self, frm, k):
        """LinearKNNSearch::find(&from, k) (linear_search.rs:52-84)."""
        return self.find_batch(np.asarray(frm).reshape(1, -1), k)[0]
