//! `smartcore::gpu` -- safe wrapper over the CUDA k-means library (cargo feature `cuda`).
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI (no Rust toolchain in the build image).  Everything with
//! behaviour worth testing lives behind the C ABI; this layer only packs `Array2` data into a
//! contiguous buffer, forwards the RNG draws of `kmeans_plus_plus` and maps status codes to `Failed`.
pub mod ffi;

use crate::error::Failed;
use crate::linalg::basic::arrays::Array2;
use crate::numbers::basenum::Number;
use std::any::TypeId;
use std::ffi::CStr;
use std::os::raw::c_void;
use std::ptr;

/// One context per calling thread (the C context is not thread-safe).
pub struct Ctx(*mut ffi::sckm_ctx);

impl Ctx {
    pub fn new(device: i32) -> Result<Ctx, Failed> {
        let mut p = ptr::null_mut();
        let rc = unsafe { ffi::sckm_ctx_create(device, &mut p) };
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

/// X packed for the C ABI: f32 / f64 pass through (`iterator(0)` yields row-major order,
/// `src/linalg/basic/matrix.rs:391-404`); every other `Number` is widened with `to_f64`.
pub struct Packed {
    buf: Vec<u8>,
    dtype: i32,
    n: u64,
    d: u64,
}

pub fn pack<TX: Number + 'static, X: Array2<TX>>(x: &X) -> Packed {
    let (n, d) = x.shape();
    if TypeId::of::<TX>() == TypeId::of::<f32>() {
        let mut buf = Vec::with_capacity(n * d * 4);
        for v in x.iterator(0) {
            buf.extend_from_slice(&(v.to_f32().unwrap()).to_le_bytes());
        }
        Packed { buf, dtype: ffi::SCKM_F32, n: n as u64, d: d as u64 }
    } else {
        let mut buf = Vec::with_capacity(n * d * 8);
        for v in x.iterator(0) {
            buf.extend_from_slice(&(v.to_f64().unwrap()).to_le_bytes());
        }
        Packed { buf, dtype: ffi::SCKM_F64, n: n as u64, d: d as u64 }
    }
}

pub struct FitOutput {
    pub y: Vec<usize>,
    pub size: Vec<usize>,
    pub centroids: Vec<Vec<f64>>,
    pub distortion: f64,
}

/// The device part of `KMeans::fit`: everything after validation and the RNG draws.
pub fn kmeans_fit(x: &Packed, k: usize, max_iter: usize, first: usize, uniforms: &[f64]) -> Result<FitOutput, Failed> {
    let ctx = Ctx::new(0)?;
    let (n, d) = (x.n as usize, x.d as usize);
    let mut y = vec![0usize; n];
    let mut size = vec![0i64; k];
    let mut c = vec![0f64; k * d];
    let (mut distortion, mut iters) = (0f64, 0i64);
    let rc = unsafe {
        ffi::sckm_kmeans_fit(
            ctx.0, x.buf.as_ptr() as *const c_void, x.n, x.d, x.dtype, 0, k as u64, max_iter as u64, first as u64,
            uniforms.as_ptr(), y.as_mut_ptr() as *mut c_void, 8, size.as_mut_ptr(), c.as_mut_ptr(), &mut distortion,
            &mut iters,
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
    })
}

/// The device part of `KMeans::predict`.
pub fn kmeans_predict(x: &Packed, centroids: &[Vec<f64>]) -> Result<Vec<u32>, Failed> {
    let ctx = Ctx::new(0).map_err(|e| Failed::predict(&e.to_string()))?;
    let flat: Vec<f64> = centroids.iter().flatten().copied().collect();
    let mut out = vec![0u32; x.n as usize];
    let rc = unsafe {
        ffi::sckm_predict(
            ctx.0, x.buf.as_ptr() as *const c_void, x.n, x.d, x.dtype, 0, flat.as_ptr(), centroids.len() as u64,
            out.as_mut_ptr() as *mut c_void, 4,
        )
    };
    if rc != ffi::SCKM_OK {
        return Err(Failed::predict(&ctx.err()));
    }
    Ok(out)
}
