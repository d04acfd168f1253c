//! Raw bindings of `include/smartcore_kmeans_cuda.h` (libsmartcore_kmeans_cuda.so), ABI version 3.
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Rust toolchain.  The signatures mirror the C header
//! one to one -- every exported symbol is declared here (tests/test_host_logic.py checks this file against the header);
//! the tested twins of this file are `smartcore_b200/cabi.py` (ctypes) and `smartcore_b200/host/smartcore_kmeans.hpp`
//! (C++), which call the same symbols.
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

pub const SCKM_ABI_VERSION: c_int = 3;
pub const SCKM_F32: c_int = 0;
pub const SCKM_F64: c_int = 1;
pub const SCKM_OK: c_int = 0;

#[link(name = "smartcore_kmeans_cuda")]
extern "C" {
    pub fn sckm_abi_version() -> c_int;
    // ---- context
    pub fn sckm_ctx_create(device: c_int, out: *mut *mut sckm_ctx) -> c_int;
    pub fn sckm_ctx_create_multi(n_dev: c_int, dev_ids: *const c_int, out: *mut *mut sckm_ctx) -> c_int;
    pub fn sckm_ctx_device_count(ctx: *const sckm_ctx) -> c_int;
    pub fn sckm_ctx_last_fit_times(ctx: *const sckm_ctx, out6: *mut f64) -> c_int;
    pub fn sckm_ctx_allreduce_path(ctx: *const sckm_ctx) -> c_int;
    pub fn sckm_ctx_destroy(ctx: *mut sckm_ctx);
    pub fn sckm_last_error(ctx: *const sckm_ctx) -> *const c_char;
    pub fn sckm_ctx_set_assign_kernel(ctx: *mut sckm_ctx, which: c_int) -> c_int;
    pub fn sckm_ctx_launch_count(ctx: *const sckm_ctx) -> u64;
    // ---- one-process-per-GPU deployments
    pub fn sckm_comm_unique_id(ctx: *mut sckm_ctx, id128: *mut c_void) -> c_int;
    pub fn sckm_comm_init_rank(ctx: *mut sckm_ctx, nranks: c_int, rank: c_int, id128: *const c_void) -> c_int;
    // ---- dataset
    pub fn sckm_dataset_upload(
        ctx: *mut sckm_ctx, host: *const c_void, n_local: u64, d: u64, dtype: c_int, column_major: c_int,
        row_offset: u64, n_global: u64, out: *mut *mut sckm_dataset,
    ) -> c_int;
    pub fn sckm_dataset_generate_blobs(
        ctx: *mut sckm_ctx, n_local: u64, d: u64, n_centers: u64, seed: u64, dtype: c_int, row_offset: u64,
        n_global: u64, out: *mut *mut sckm_dataset,
    ) -> c_int;
    pub fn sckm_blobs_fill_host(
        out: *mut c_void, dtype: c_int, row0: u64, nrows: u64, d: u64, n_centers: u64, seed: u64,
    ) -> c_int;
    pub fn sckm_dataset_download_rows(ds: *mut sckm_dataset, local_row0: u64, nrows: u64, host_out: *mut c_void) -> c_int;
    pub fn sckm_dataset_destroy(ds: *mut sckm_dataset);
    // ---- the pieces of KMeans::fit
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
    pub fn sckm_lloyd_iterate(
        ds: *mut sckm_dataset, k: u64, n_iters: u64, centroids_inout: *mut f64, size_out: *mut i64,
        inertia_out: *mut f64, ms_per_iter_out: *mut f32, assign_ms_out: *mut f32,
    ) -> c_int;
    pub fn sckm_labels_download(ds: *mut sckm_dataset, out: *mut c_void, width: c_int) -> c_int;
    pub fn sckm_mindist_download(ds: *mut sckm_dataset, out: *mut f64) -> c_int;
    // ---- whole-matrix calls (what KMeans::fit / predict use)
    pub fn sckm_predict(
        ctx: *mut sckm_ctx, x_host: *const c_void, n: u64, d: u64, dtype: c_int, column_major: c_int,
        centroids: *const f64, k: u64, labels_out: *mut c_void, width: c_int,
    ) -> c_int;
    pub fn sckm_kmeans_fit(
        ctx: *mut sckm_ctx, x_host: *const c_void, n: u64, d: u64, dtype: c_int, column_major: c_int, k: u64,
        max_iter: u64, first_index: u64, uniforms: *const f64, labels_out: *mut c_void, width: c_int,
        size_out: *mut i64, centroids_out: *mut f64, distortion_out: *mut f64, iters_out: *mut i64,
    ) -> c_int;
    pub fn sckm_kmeans_fit_shard(
        ctx: *mut sckm_ctx, x_local: *const c_void, n_local: u64, d: u64, dtype: c_int, column_major: c_int,
        row_offset: u64, n_global: u64, k: u64, max_iter: u64, first_index: u64, uniforms: *const f64,
        labels_out: *mut c_void, width: c_int, size_out: *mut i64, centroids_out: *mut f64,
        distortion_out: *mut f64, iters_out: *mut i64,
    ) -> c_int;
    // ---- cluster quality (metrics/cluster_helpers.rs:7-25): contingency table counted on the device
    pub fn sckm_contingency(
        ds: *mut sckm_dataset, class_ids_host: *const u32, n_classes: u64, k: u64, out: *mut i64,
    ) -> c_int;
    pub fn sckm_contingency_host(
        ctx: *mut sckm_ctx, a_host: *const u32, b_host: *const u32, n: u64, na: u64, nb: u64, out: *mut i64,
    ) -> c_int;
    // ---- batched LinearKNNSearch::find / find_radius with Euclidian::distance (linear_search.rs:52-110)
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
    // ---- measurement helpers
    pub fn sckm_device_peaks(ctx: *mut sckm_ctx, out3: *mut f64) -> c_int;
    pub fn sckm_flush_l2(ctx: *mut sckm_ctx) -> c_int;
}
