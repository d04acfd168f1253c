"""ctypes binding of include/smartcore_kmeans_cuda.h (libsmartcore_kmeans_cuda.so).

Thin and mechanical on purpose: every method is one C-ABI call with numpy buffers.  The library is
the product; this file only marshals pointers.  Missing library => ImportError (no fallback).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsmartcore_kmeans_cuda.so")
if os.environ.get("SCKM_LIB_VARIANT"):      # A/B builds of the same sources with experiment flags (tools/build_variant.sh)
    LIB_PATH = os.path.join(_HERE, "lib", "libsmartcore_kmeans_cuda.%s.so" % os.environ["SCKM_LIB_VARIANT"])

F32, F64 = 0, 1
ASSIGN_AUTO, ASSIGN_DIRECT, ASSIGN_DMMA, ASSIGN_STREAM, ASSIGN_TC5 = 0, 1, 2, 3, 4

# every symbol include/smartcore_kmeans_cuda.h declares (checked by tests/test_cabi_symbols.py)
SYMBOLS = [
    "sckm_abi_version", "sckm_ctx_create", "sckm_ctx_destroy", "sckm_last_error", "sckm_ctx_set_assign_kernel",
    "sckm_ctx_launch_count", "sckm_comm_unique_id", "sckm_comm_init_rank", "sckm_dataset_upload",
    "sckm_dataset_generate_blobs", "sckm_blobs_fill_host", "sckm_dataset_download_rows", "sckm_dataset_destroy",
    "sckm_kmeanspp", "sckm_init_centroids", "sckm_lloyd_step", "sckm_lloyd_fit", "sckm_lloyd_iterate",
    "sckm_labels_download", "sckm_mindist_download", "sckm_predict", "sckm_kmeans_fit", "sckm_device_peaks",
    "sckm_contingency", "sckm_contingency_host", "sckm_knn", "sckm_radius_count", "sckm_radius_fill",
    "sckm_flush_l2", "sckm_ctx_create_multi", "sckm_ctx_device_count", "sckm_ctx_last_fit_times",
    "sckm_kmeans_fit_shard", "sckm_ctx_allreduce_path",
]


class SckmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sckm error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(smartcore_b200 has no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.sckm_abi_version.restype = i32
    L.sckm_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.sckm_ctx_create_multi.argtypes = [i32, vp, C.POINTER(vp)]
    L.sckm_ctx_device_count.argtypes = [vp]; L.sckm_ctx_device_count.restype = i32
    L.sckm_ctx_last_fit_times.argtypes = [vp, vp]
    L.sckm_ctx_allreduce_path.argtypes = [vp]; L.sckm_ctx_allreduce_path.restype = i32
    L.sckm_ctx_destroy.argtypes = [vp]; L.sckm_ctx_destroy.restype = None
    L.sckm_last_error.argtypes = [vp]; L.sckm_last_error.restype = C.c_char_p
    L.sckm_ctx_set_assign_kernel.argtypes = [vp, i32]
    L.sckm_ctx_launch_count.argtypes = [vp]; L.sckm_ctx_launch_count.restype = u64
    L.sckm_comm_unique_id.argtypes = [vp, vp]
    L.sckm_comm_init_rank.argtypes = [vp, i32, i32, vp]
    L.sckm_dataset_upload.argtypes = [vp, vp, u64, u64, i32, i32, u64, u64, C.POINTER(vp)]
    L.sckm_dataset_generate_blobs.argtypes = [vp, u64, u64, u64, u64, i32, u64, u64, C.POINTER(vp)]
    L.sckm_blobs_fill_host.argtypes = [vp, i32, u64, u64, u64, u64, u64]
    L.sckm_dataset_download_rows.argtypes = [vp, u64, u64, vp]
    L.sckm_dataset_destroy.argtypes = [vp]; L.sckm_dataset_destroy.restype = None
    L.sckm_kmeanspp.argtypes = [vp, u64, u64, vp, vp, vp]
    L.sckm_init_centroids.argtypes = [vp, u64, vp, vp]
    L.sckm_lloyd_step.argtypes = [vp, vp, u64, vp, vp, vp]
    L.sckm_lloyd_fit.argtypes = [vp, u64, u64, vp, vp, vp, vp]
    L.sckm_lloyd_iterate.argtypes = [vp, u64, u64, vp, vp, vp, vp, vp]
    L.sckm_labels_download.argtypes = [vp, vp, i32]
    L.sckm_mindist_download.argtypes = [vp, vp]
    L.sckm_predict.argtypes = [vp, vp, u64, u64, i32, i32, vp, u64, vp, i32]
    L.sckm_kmeans_fit.argtypes = [vp, vp, u64, u64, i32, i32, u64, u64, u64, vp, vp, i32, vp, vp, vp, vp]
    L.sckm_kmeans_fit_shard.argtypes = [vp, vp, u64, u64, i32, i32, u64, u64, u64, u64, u64, vp, vp, i32, vp, vp, vp, vp]
    L.sckm_device_peaks.argtypes = [vp, vp]
    L.sckm_contingency.argtypes = [vp, vp, u64, u64, vp]
    L.sckm_knn.argtypes = [vp, vp, u64, u64, vp, vp]
    L.sckm_radius_count.argtypes = [vp, vp, u64, C.c_double, vp]
    L.sckm_radius_fill.argtypes = [vp, vp, u64, C.c_double, vp, u64, vp, vp]
    L.sckm_contingency_host.argtypes = [vp, vp, vp, u64, u64, u64, vp]
    L.sckm_flush_l2.argtypes = [vp]
    return L


lib = _load()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dtype_code(a):
    if a.dtype == np.float32:
        return F32
    if a.dtype == np.float64:
        return F64
    raise TypeError("X must be float32 or float64 (got %s)" % a.dtype)


def blobs_host(row0, nrows, d, n_centers, seed, dtype=np.float64):
    """Host twin of the device blob generator (bit-identical rows; pure CPU)."""
    out = np.empty((nrows, d), dtype=dtype)
    rc = lib.sckm_blobs_fill_host(_p(out), dtype_code(out), row0, nrows, d, n_centers, seed)
    if rc:
        raise SckmError(rc, "sckm_blobs_fill_host: invalid argument")
    return out


class Context:
    """sckm_ctx: one CUDA device + stream (+ NCCL communicator when joined)."""

    def __init__(self, device=0, devices=None):
        """device: one CUDA device.  devices: a list of devices (or "all") behind ONE context (sckm_ctx_create_multi):
        kmeans_fit / predict then shard the rows of the host matrix over them."""
        h = C.c_void_p()
        if devices is None:
            rc = lib.sckm_ctx_create(device, C.byref(h))
        elif isinstance(devices, str):
            rc = lib.sckm_ctx_create_multi(0, None, C.byref(h))
        else:
            ids = (C.c_int * len(devices))(*devices)
            rc = lib.sckm_ctx_create_multi(len(devices), ids, C.byref(h))
        if rc:
            raise SckmError(rc, (lib.sckm_last_error(None) or b"").decode())
        self.h = h
        self.nranks, self.rank = 1, 0

    def device_count(self):
        return int(lib.sckm_ctx_device_count(self.h))

    def allreduce_path(self):
        """How the last Lloyd loop summed over the ranks: 'none' (one rank), 'nccl' or 'peer' (inside the finalize kernel)."""
        return ("none", "nccl", "peer")[int(lib.sckm_ctx_allreduce_path(self.h))]

    def last_fit_times(self):
        out = np.zeros(6)
        self._check(lib.sckm_ctx_last_fit_times(self.h, _p(out)))
        return dict(upload_s=out[0], kmeanspp_init_s=out[1], lloyd_s=out[2], download_s=out[3], total_s=out[4], devices=int(out[5]))

    def _check(self, rc):
        if rc:
            raise SckmError(rc, (lib.sckm_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            lib.sckm_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_assign_kernel(self, which):
        self._check(lib.sckm_ctx_set_assign_kernel(self.h, which))

    def launch_count(self):
        return int(lib.sckm_ctx_launch_count(self.h))

    def comm_unique_id(self):
        buf = (C.c_ubyte * 128)()
        self._check(lib.sckm_comm_unique_id(self.h, buf))
        return bytes(buf)

    def comm_init_rank(self, nranks, rank, id128):
        buf = (C.c_ubyte * 128).from_buffer_copy(id128) if id128 is not None else None
        self._check(lib.sckm_comm_init_rank(self.h, nranks, rank, buf))
        self.nranks, self.rank = nranks, rank

    def upload(self, x, column_major=False, row_offset=0, n_global=0):
        """x: (n_local, d) array.  column_major=True: x.T is what lies in memory (DenseMatrix default)."""
        n, d = x.shape
        buf = np.ascontiguousarray(x.T) if column_major else np.ascontiguousarray(x)
        h = C.c_void_p()
        self._check(lib.sckm_dataset_upload(self.h, _p(buf), n, d, dtype_code(buf), 1 if column_major else 0,
                                            row_offset, n_global or n, C.byref(h)))
        return Dataset(self, h, n, d, buf.dtype)

    def upload_colmajor_image(self, image, n, d):
        """image: C-contiguous (d, n) array = the column-major image of an n x d matrix (no host transpose here)."""
        assert image.shape == (d, n) and image.flags.c_contiguous
        h = C.c_void_p()
        self._check(lib.sckm_dataset_upload(self.h, _p(image), n, d, dtype_code(image), 1, 0, n, C.byref(h)))
        return Dataset(self, h, n, d, image.dtype)

    def generate_blobs(self, n_local, d, n_centers, seed, dtype=np.float64, row_offset=0, n_global=0):
        h = C.c_void_p()
        code = F32 if np.dtype(dtype) == np.float32 else F64
        self._check(lib.sckm_dataset_generate_blobs(self.h, n_local, d, n_centers, seed, code, row_offset,
                                                    n_global or n_local, C.byref(h)))
        return Dataset(self, h, n_local, d, np.dtype(dtype))

    def predict(self, x, centroids, column_major=False, width=8):
        n, d = x.shape
        buf = np.ascontiguousarray(x.T) if column_major else np.ascontiguousarray(x)
        c = np.ascontiguousarray(centroids, dtype=np.float64)
        out = np.empty(n, dtype=np.uint64 if width == 8 else np.uint32)
        self._check(lib.sckm_predict(self.h, _p(buf), n, d, dtype_code(buf), 1 if column_major else 0, _p(c),
                                     c.shape[0], _p(out), width))
        return out

    def kmeans_fit(self, x, k, max_iter, first_index, uniforms, column_major=False):
        """sckm_kmeans_fit: the whole KMeans::fit from host buffers.  Returns dict."""
        n, d = x.shape
        buf = np.ascontiguousarray(x.T) if column_major else np.ascontiguousarray(x)
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        labels = np.empty(n, dtype=np.uint64); size = np.zeros(k, dtype=np.int64); cent = np.zeros((k, d))
        dist = C.c_double(0); iters = C.c_int64(0)
        self._check(lib.sckm_kmeans_fit(self.h, _p(buf), n, d, dtype_code(buf), 1 if column_major else 0, k, max_iter,
                                        first_index, _p(u), _p(labels), 8, _p(size), _p(cent), C.addressof(dist),
                                        C.addressof(iters)))
        return dict(labels=labels, size=size, centroids=cent, distortion=dist.value, iters=iters.value)

    def kmeans_fit_shard(self, x_local, row_offset, n_global, k, max_iter, first_index, uniforms, width=8):
        """sckm_kmeans_fit_shard: this rank's rows [row_offset, row_offset + len(x_local)) of an n_global-row matrix."""
        n, d = x_local.shape
        buf = np.ascontiguousarray(x_local)
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        labels = np.empty(n, dtype=np.uint64 if width == 8 else np.uint32); size = np.zeros(k, dtype=np.int64); cent = np.zeros((k, d))
        dist = C.c_double(0); iters = C.c_int64(0)
        self._check(lib.sckm_kmeans_fit_shard(self.h, _p(buf), n, d, dtype_code(buf), 0, row_offset, n_global, k, max_iter,
                                              first_index, _p(u), _p(labels), width, _p(size), _p(cent), C.addressof(dist),
                                              C.addressof(iters)))
        return dict(labels=labels, size=size, centroids=cent, distortion=dist.value, iters=iters.value)

    def device_peaks(self):
        out = np.zeros(3)
        self._check(lib.sckm_device_peaks(self.h, _p(out)))
        return dict(hbm_copy_gbs=out[0], fp64_dfma_tflops=out[1], fp64_dmma_tflops=out[2])

    def flush_l2(self):
        self._check(lib.sckm_flush_l2(self.h))


class Dataset:
    """sckm_dataset: this rank's rows of X resident in HBM with their labels and D^2 array."""

    def __init__(self, ctx, h, n, d, dtype):
        self.ctx, self.h, self.n, self.d, self.dtype = ctx, h, n, d, dtype

    def close(self):
        if self.h:
            lib.sckm_dataset_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def download_rows(self, row0, nrows):
        out = np.empty((nrows, self.d), dtype=self.dtype)
        self.ctx._check(lib.sckm_dataset_download_rows(self.h, row0, nrows, _p(out)))
        return out

    def kmeanspp(self, k, first_index=0, uniforms=None, inject_rows=None):
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        inj = None if inject_rows is None else np.ascontiguousarray(inject_rows, dtype=np.int64)
        seeds = np.zeros(k, dtype=np.int64)
        self.ctx._check(lib.sckm_kmeanspp(self.h, k, first_index, _p(u), _p(inj), _p(seeds)))
        return seeds

    def init_centroids(self, k):
        cent = np.zeros((k, self.d)); size = np.zeros(k, dtype=np.int64)
        self.ctx._check(lib.sckm_init_centroids(self.h, k, _p(cent), _p(size)))
        return cent, size

    def lloyd_step(self, centroids):
        c = np.ascontiguousarray(centroids, dtype=np.float64); k = c.shape[0]
        sums = np.zeros((k, self.d)); counts = np.zeros(k, dtype=np.int64); inertia = C.c_double(0)
        self.ctx._check(lib.sckm_lloyd_step(self.h, _p(c), k, _p(sums), _p(counts), C.addressof(inertia)))
        return inertia.value, sums, counts

    def lloyd_fit(self, centroids, max_iter):
        c = np.array(centroids, dtype=np.float64, order="C"); k = c.shape[0]
        size = np.zeros(k, dtype=np.int64); dist = C.c_double(0); iters = C.c_int64(0)
        self.ctx._check(lib.sckm_lloyd_fit(self.h, k, max_iter, _p(c), _p(size), C.addressof(dist), C.addressof(iters)))
        return dict(centroids=c, size=size, distortion=dist.value, iters=iters.value)

    def lloyd_iterate(self, centroids, n_iters, want_inertia=False, per_step_events=True):
        """per_step_events=False: time the n_iters steps as a whole (two events, as a fit runs); `ms` then carries the mean
        in every slot and `assign_ms` is None."""
        c = np.array(centroids, dtype=np.float64, order="C"); k = c.shape[0]
        size = np.zeros(k, dtype=np.int64); ms = np.zeros(n_iters, dtype=np.float32)
        ams = np.zeros(n_iters, dtype=np.float32) if per_step_events else None
        inertia = np.zeros(n_iters) if want_inertia else None
        self.ctx._check(lib.sckm_lloyd_iterate(self.h, k, n_iters, _p(c), _p(size), _p(inertia), _p(ms), _p(ams)))
        return dict(centroids=c, size=size, ms=ms, assign_ms=ams, inertia=inertia)

    def labels(self, width=8):
        out = np.empty(self.n, dtype=np.uint64 if width == 8 else np.uint32)
        self.ctx._check(lib.sckm_labels_download(self.h, _p(out), width))
        return out

    def contingency(self, class_ids, n_classes, k):
        """sckm_contingency: table[n_classes][k] of (class id, resident cluster label) pairs, counted on the device."""
        a = np.ascontiguousarray(class_ids, dtype=np.uint32)
        assert a.size == self.n
        out = np.zeros((n_classes, k), dtype=np.int64)
        self.ctx._check(lib.sckm_contingency(self.h, _p(a), n_classes, k, _p(out)))
        return out

    def knn(self, queries, k):
        """sckm_knn: for each query row the k nearest resident rows -> (idx [nq, k] int64, dist [nq, k] float64)."""
        q = np.ascontiguousarray(queries, dtype=self.dtype)
        assert q.ndim == 2 and q.shape[1] == self.d
        idx = np.zeros((q.shape[0], k), dtype=np.int64); dist = np.zeros((q.shape[0], k))
        self.ctx._check(lib.sckm_knn(self.h, _p(q), q.shape[0], k, _p(idx), _p(dist)))
        return idx, dist

    def radius(self, queries, radius):
        """sckm_radius_count + sckm_radius_fill: per query (idx, dist) arrays of the rows within `radius`, row order."""
        q = np.ascontiguousarray(queries, dtype=self.dtype)
        assert q.ndim == 2 and q.shape[1] == self.d
        nq = q.shape[0]
        counts = np.zeros(nq, dtype=np.int64)
        self.ctx._check(lib.sckm_radius_count(self.h, _p(q), nq, float(radius), _p(counts)))
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        total = int(offsets[-1])
        idx = np.zeros(max(total, 1), dtype=np.int64); dist = np.zeros(max(total, 1))
        self.ctx._check(lib.sckm_radius_fill(self.h, _p(q), nq, float(radius), _p(offsets[:-1].copy()), total, _p(idx), _p(dist)))
        return [(idx[offsets[i]:offsets[i + 1]], dist[offsets[i]:offsets[i + 1]]) for i in range(nq)]

    def mindist(self):
        out = np.empty(self.n)
        self.ctx._check(lib.sckm_mindist_download(self.h, _p(out)))
        return out
