//! `smartcore::gpu` -- safe wrapper over the CUDA k-means library (cargo feature `cuda`).
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI (no Rust toolchain in the build image); written to compile by inspection
//! against smartcore v0.4.0 with the small additions of `rust/smartcore.patch`:
//!   * `Number::GPU_DTYPE` (an associated const, `None` by default, overridden by the `f32` / `f64` impls) replaces
//!     any `TypeId` dispatch -- `Number` carries no `'static` bound (src/numbers/basenum.rs:8-23);
//!   * `Array2::as_contiguous()` (defaulted to `None`, overridden by `DenseMatrix` and the ndarray binding) hands the
//!     backing slice and its layout to the C ABI without a copy (src/linalg/basic/matrix.rs:27-32,
//!     src/linalg/ndarray/matrix.rs:18-45).
//! Everything with behaviour worth testing lives behind the C ABI; this layer only borrows or packs the matrix,
//! forwards the RNG draws of `kmeans_plus_plus` and maps status codes to `Failed`.
pub mod ffi;

use crate::error::{Failed, FailedError};
use crate::linalg::basic::arrays::Array2;
use crate::numbers::basenum::Number;
use std::borrow::Cow;
use std::cell::RefCell;
use std::ffi::CStr;
use std::os::raw::c_void;
use std::ptr;

/// RAII over `sckm_ctx`.  Not `Send`: the C context is not thread-safe, each calling thread keeps its own.
pub struct Ctx(*mut ffi::sckm_ctx);

impl Ctx {
    /// ONE context over every visible GPU (`SMARTCORE_CUDA_DEVICES="0,2"` restricts them): `KMeans::fit` / `predict`
    /// shard the rows over the whole box with no change to the caller.
    pub fn new() -> Result<Ctx, Failed> {
        let ids: Vec<i32> = std::env::var("SMARTCORE_CUDA_DEVICES")
            .map(|s| s.split(',').filter_map(|t| t.trim().parse().ok()).collect())
            .unwrap_or_default();
        let mut p = ptr::null_mut();
        let rc = unsafe {
            ffi::sckm_ctx_create_multi(ids.len() as i32, if ids.is_empty() { ptr::null() } else { ids.as_ptr() }, &mut p)
        };
        if rc != ffi::SCKM_OK {
            let msg = unsafe { CStr::from_ptr(ffi::sckm_last_error(ptr::null())) };
            return Err(Failed::fit(&format!("CUDA backend unavailable: {}", msg.to_string_lossy())));
        }
        Ok(Ctx(p))
    }
    fn err(&self) -> String {
        unsafe { CStr::from_ptr(ffi::sckm_last_error(self.0)) }.to_string_lossy().into_owned()
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { ffi::sckm_ctx_destroy(self.0) }
    }
}

thread_local! {
    // created on the first fit / predict of this thread and kept: stream, pinned staging ring, device memory pool and
    // (multi-GPU) NCCL communicators are set up once, not per call
    static CTX: RefCell<Option<Ctx>> = RefCell::new(None);
}

fn with_ctx<R>(f: impl FnOnce(&Ctx) -> Result<R, Failed>) -> Result<R, Failed> {
    CTX.with(|slot| {
        let mut slot = slot.borrow_mut();
        if slot.is_none() {
            *slot = Some(Ctx::new()?);
        }
        f(slot.as_ref().unwrap())
    })
}

/// X as the C ABI takes it: a contiguous buffer of f32 / f64 plus its layout.
pub struct Packed<'a> {
    bytes: Cow<'a, [u8]>,
    dtype: i32,
    column_major: i32,
    n: u64,
    d: u64,
}

fn as_bytes<T>(s: &[T]) -> &[u8] {
    // plain-old-data view of a numeric slice (f32 / f64 only reach this)
    unsafe { std::slice::from_raw_parts(s.as_ptr() as *const u8, std::mem::size_of_val(s)) }
}

/// Borrow the matrix when its container exposes contiguous f32 / f64 storage (`DenseMatrix`, C- or F-ordered
/// ndarray): zero copies on the host, the column-major case is transposed on the device.  Anything else (views,
/// integer element types) is packed row-major through `iterator(0)`, widening with `to_f64` like the reference's
/// own arithmetic does (bbd_tree.rs:207-213).
pub fn pack<'a, TX: Number, X: Array2<TX>>(x: &'a X) -> Packed<'a> {
    let (n, d) = x.shape();
    if let (Some(dtype), Some((slice, column_major))) = (TX::GPU_DTYPE, x.as_contiguous()) {
        return Packed { bytes: Cow::Borrowed(as_bytes(slice)), dtype, column_major: column_major as i32, n: n as u64, d: d as u64 };
    }
    if TX::GPU_DTYPE == Some(ffi::SCKM_F32) {
        let v: Vec<f32> = x.iterator(0).map(|v| v.to_f32().unwrap()).collect();
        return Packed { bytes: Cow::Owned(as_bytes(&v).to_vec()), dtype: ffi::SCKM_F32, column_major: 0, n: n as u64, d: d as u64 };
    }
    let v: Vec<f64> = x.iterator(0).map(|v| v.to_f64().unwrap()).collect();
    Packed { bytes: Cow::Owned(as_bytes(&v).to_vec()), dtype: ffi::SCKM_F64, column_major: 0, n: n as u64, d: d as u64 }
}

pub struct FitOutput {
    pub y: Vec<usize>,
    pub size: Vec<usize>,
    pub centroids: Vec<Vec<f64>>,
    pub distortion: f64,
    pub iterations: usize,
}

/// The device part of `KMeans::fit`: everything after validation and the RNG draws (kmeans.rs:271-322).
pub fn kmeans_fit(x: &Packed, k: usize, max_iter: usize, first: usize, uniforms: &[f64]) -> Result<FitOutput, Failed> {
    with_ctx(|ctx| {
        let (n, d) = (x.n as usize, x.d as usize);
        let mut y = vec![0usize; n]; // usize == u64 on every target CUDA supports: labels land in place (width 8)
        let mut size = vec![0i64; k];
        let mut c = vec![0f64; k * d];
        let (mut distortion, mut iters) = (0f64, 0i64);
        let rc = unsafe {
            ffi::sckm_kmeans_fit(
                ctx.0, x.bytes.as_ptr() as *const c_void, x.n, x.d, x.dtype, x.column_major, k as u64, max_iter as u64,
                first as u64, uniforms.as_ptr(), y.as_mut_ptr() as *mut c_void, 8, size.as_mut_ptr(), c.as_mut_ptr(),
                &mut distortion, &mut iters,
            )
        };
        if rc != ffi::SCKM_OK {
            return Err(Failed::fit(&ctx.err()));
        }
        Ok(FitOutput {
            y,
            size: size.into_iter().map(|v| v as usize).collect(),
            centroids: c.chunks(d).map(|r| r.to_vec()).collect(),
            distortion,
            iterations: iters as usize,
        })
    })
}

/// The device part of `KMeans::predict` (kmeans.rs:327-352).
pub fn kmeans_predict(x: &Packed, centroids: &[Vec<f64>]) -> Result<Vec<u32>, Failed> {
    with_ctx(|ctx| {
        let flat: Vec<f64> = centroids.iter().flatten().copied().collect();
        let mut out = vec![0u32; x.n as usize];
        let rc = unsafe {
            ffi::sckm_predict(
                ctx.0, x.bytes.as_ptr() as *const c_void, x.n, x.d, x.dtype, x.column_major, flat.as_ptr(),
                centroids.len() as u64, out.as_mut_ptr() as *mut c_void, 4,
            )
        };
        if rc != ffi::SCKM_OK {
            return Err(Failed::predict(&ctx.err()));
        }
        Ok(out)
    })
    .map_err(|e| match e.error() {
        FailedError::PredictFailed => e,
        _ => Failed::predict(&e.to_string()), // a context that could not be created reports through `fit`'s constructor
    })
}
