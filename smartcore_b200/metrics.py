"""smartcore::metrics::{cluster_helpers, cluster_hcv} -- the reference's cluster-quality scores
(src/metrics/cluster_helpers.rs:7-104, src/metrics/cluster_hcv.rs:12-55) over the C++ host mirror; the pair
counting runs on the GPU (sckm_contingency / sckm_contingency_host).  No CPU fallback."""
import ctypes as C

import numpy as np

from .cluster import Failed, _h, _p

_vp = C.c_void_p
_h.sch_contingency_matrix.argtypes = [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _vp, _vp, C.c_char_p, C.c_size_t]
_h.sch_entropy.argtypes = [_vp, C.c_size_t]; _h.sch_entropy.restype = C.c_double
_h.sch_mutual_info_score.argtypes = [_vp, C.c_size_t, C.c_size_t]; _h.sch_mutual_info_score.restype = C.c_double
_h.sch_hcv_score.argtypes = [_vp, _vp, C.c_size_t, _vp, C.c_char_p, C.c_size_t]
_h.sch_hcv_from_table.argtypes = [_vp, C.c_size_t, C.c_size_t, _vp]; _h.sch_hcv_from_table.restype = None


def _i64(v):
    return np.ascontiguousarray(v, dtype=np.int64)


def contingency_matrix(labels_true, labels_pred):
    """cluster_helpers.rs:7-25; rows = sorted unique true labels, columns = sorted unique predicted labels."""
    a, b = _i64(labels_true), _i64(labels_pred)
    if a.shape != b.shape:
        raise Failed("Error in input, check parameters: label vectors differ in length")
    nr, nc = C.c_size_t(0), C.c_size_t(0)
    err = C.create_string_buffer(1024)
    cap = 1 << 12
    while True:
        out = np.zeros(cap, dtype=np.int64)
        rc = _h.sch_contingency_matrix(_p(a), _p(b), a.size, _p(out), cap, C.addressof(nr), C.addressof(nc), err, len(err))
        if rc == 3:
            cap = nr.value * nc.value
            continue
        if rc:
            raise Failed(err.value.decode())
        return out[: nr.value * nc.value].reshape(nr.value, nc.value)


def entropy(data):
    """cluster_helpers.rs:27-48"""
    a = _i64(data)
    return _h.sch_entropy(_p(a), a.size)


def mutual_info_score(contingency):
    """cluster_helpers.rs:50-104"""
    t = _i64(contingency)
    return _h.sch_mutual_info_score(_p(t), t.shape[0], t.shape[1])


class HCVScore:
    """cluster_hcv.rs:12-55: homogeneity, completeness and V-measure."""

    def __init__(self):
        self._h = self._c = self._v = None

    @classmethod
    def new(cls):
        return cls()

    def homogeneity(self):
        return self._h

    def completeness(self):
        return self._c

    def v_measure(self):
        return self._v

    def compute(self, y_true, y_pred):
        a, b = _i64(y_true), _i64(y_pred)
        out = np.zeros(3); err = C.create_string_buffer(1024)
        if a.shape != b.shape:
            raise Failed("Error in input, check parameters: label vectors differ in length")
        if _h.sch_hcv_score(_p(a), _p(b), a.size, _p(out), err, len(err)):
            raise Failed(err.value.decode())
        self._h, self._c, self._v = out.tolist()
        return self

    def compute_from_table(self, contingency):
        """Scores from a table counted on the device against resident labels (Dataset.contingency)."""
        t = _i64(contingency)
        out = np.zeros(3)
        _h.sch_hcv_from_table(_p(t), t.shape[0], t.shape[1], _p(out))
        self._h, self._c, self._v = out.tolist()
        return self
