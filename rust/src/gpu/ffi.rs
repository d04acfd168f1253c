//! Raw bindings of `include/smartcore_kmeans_cuda.h` (libsmartcore_kmeans_cuda.so).
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Rust toolchain.  The signatures
//! mirror the C header one to one; the tested twin of this file is `smartcore_b200/cabi.py`
//! (ctypes) and `smartcore_b200/host/smartcore_kmeans.hpp` (C++), which call the same symbols.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct sckm_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct sckm_dataset {
    _private: [u8; 0],
}

pub const SCKM_F32: c_int = 0;
pub const SCKM_F64: c_int = 1;
pub const SCKM_OK: c_int = 0;

#[link(name = "smartcore_kmeans_cuda")]
extern "C" {
    pub fn sckm_abi_version() -> c_int;
    pub fn sckm_ctx_create(device: c_int, out: *mut *mut sckm_ctx) -> c_int;
    pub fn sckm_ctx_destroy(ctx: *mut sckm_ctx);
    pub fn sckm_last_error(ctx: *const sckm_ctx) -> *const c_char;
    pub fn sckm_comm_unique_id(ctx: *mut sckm_ctx, id128: *mut c_void) -> c_int;
    pub fn sckm_comm_init_rank(ctx: *mut sckm_ctx, nranks: c_int, rank: c_int, id128: *const c_void) -> c_int;
    pub fn sckm_dataset_upload(
        ctx: *mut sckm_ctx, host: *const c_void, n_local: u64, d: u64, dtype: c_int, column_major: c_int,
        row_offset: u64, n_global: u64, out: *mut *mut sckm_dataset,
    ) -> c_int;
    pub fn sckm_dataset_destroy(ds: *mut sckm_dataset);
    pub fn sckm_kmeanspp(
        ds: *mut sckm_dataset, k: u64, first_index: u64, uniforms: *const f64, inject_rows: *const i64,
        seed_rows_out: *mut i64,
    ) -> c_int;
    pub fn sckm_init_centroids(ds: *mut sckm_dataset, k: u64, centroids_out: *mut f64, size_out: *mut i64) -> c_int;
    pub fn sckm_lloyd_step(
        ds: *mut sckm_dataset, centroids: *const f64, k: u64, sums_out: *mut f64, counts_out: *mut i64,
        inertia_out: *mut f64,
    ) -> c_int;
    pub fn sckm_lloyd_fit(
        ds: *mut sckm_dataset, k: u64, max_iter: u64, centroids_inout: *mut f64, size_out: *mut i64,
        distortion_out: *mut f64, iters_out: *mut i64,
    ) -> c_int;
    pub fn sckm_labels_download(ds: *mut sckm_dataset, out: *mut c_void, width: c_int) -> c_int;
    pub fn sckm_predict(
        ctx: *mut sckm_ctx, x_host: *const c_void, n: u64, d: u64, dtype: c_int, column_major: c_int,
        centroids: *const f64, k: u64, labels_out: *mut c_void, width: c_int,
    ) -> c_int;
    pub fn sckm_kmeans_fit(
        ctx: *mut sckm_ctx, x_host: *const c_void, n: u64, d: u64, dtype: c_int, column_major: c_int, k: u64,
        max_iter: u64, first_index: u64, uniforms: *const f64, labels_out: *mut c_void, width: c_int,
        size_out: *mut i64, centroids_out: *mut f64, distortion_out: *mut f64, iters_out: *mut i64,
    ) -> c_int;
    // cluster quality (metrics/cluster_helpers.rs:7-25): contingency table counted on the device
    pub fn sckm_contingency(
        ds: *mut sckm_dataset, class_ids_host: *const u32, n_classes: u64, k: u64, out: *mut i64,
    ) -> c_int;
    // batched LinearKNNSearch::find with Euclidian::distance (linear_search.rs:52-84)
    pub fn sckm_knn(
        ds: *mut sckm_dataset, queries_host: *const c_void, nq: u64, k: u64, idx_out: *mut i64, dist_out: *mut f64,
    ) -> c_int;
    pub fn sckm_radius_count(
        ds: *mut sckm_dataset, queries_host: *const c_void, nq: u64, radius: f64, counts_out: *mut i64,
    ) -> c_int;
    pub fn sckm_radius_fill(
        ds: *mut sckm_dataset, queries_host: *const c_void, nq: u64, radius: f64, offsets: *const i64, total: u64,
        idx_out: *mut i64, dist_out: *mut f64,
    ) -> c_int;
    pub fn sckm_contingency_host(
        ctx: *mut sckm_ctx, a_host: *const u32, b_host: *const u32, n: u64, na: u64, nb: u64, out: *mut i64,
    ) -> c_int;
}
