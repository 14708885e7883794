//! dxt-lossless-transform-cuda — thin `extern "C"` FFI over libdxt_lossless_transform_cuda.so.
//!
//! SOURCE ONLY: the build image has no Rust toolchain, so this file documents the binding a
//! maintainer would compile; the C ABI it binds is exercised by the Python/ctypes tests instead.
//!
//! The public functions keep the signatures of the reference crates so the crate is a drop-in for
//! the hot path:
//!   dxt_lossless_transform_bc1::{transform_bc1_with_settings, untransform_bc1_with_settings}
//!   (core/dxt-lossless-transform-bc1/src/transform/transform_with_settings.rs:31,92), the bc2/bc3
//!   equivalents, and transform_bcN_auto (transform_auto.rs:200 / 196 / 196).
//! No CPU fallback: a CUDA failure surfaces as an error (safe API) or a panic (raw-pointer API,
//! whose reference signature has no return value).
#![allow(non_camel_case_types)]

pub mod extensions; // zstd estimator (host threads) and experimental::normalize_blocks
pub mod file_formats; // TransformBundle / dispatch / DdsHandler over the dltff_* and dltdds_* symbols

use core::ffi::c_void;
use dxt_lossless_transform_api_common::estimate::SizeEstimationOperations;
use dxt_lossless_transform_bc1::Bc1TransformSettings;
use dxt_lossless_transform_bc2::Bc2TransformSettings;
use dxt_lossless_transform_bc3::Bc3TransformSettings;
use dxt_lossless_transform_common::color_565::YCoCgVariant;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct DltResult {
    pub error_code: i32,
}

/// Core-crate settings layout (bool first, INTERNAL variant numbering).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct DltCoreSettings {
    pub split_colour_endpoints: bool,
    pub decorrelation_mode: u8,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct DltCoreBc3Settings {
    pub split_alpha_endpoints: bool,
    pub split_colour_endpoints: bool,
    pub decorrelation_mode: u8,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct DltCoreAutoSettings {
    pub use_all_modes: bool,
}

#[repr(C)]
pub struct DltSizeEstimator {
    pub context: *mut c_void,
    pub max_compressed_size: unsafe extern "C" fn(*mut c_void, usize, *mut usize) -> u32,
    pub estimate_compressed_size:
        unsafe extern "C" fn(*mut c_void, *const u8, usize, *mut u8, usize, *mut usize) -> u32,
}

extern "C" {
    fn dltbc1core_transform(i: *const u8, il: usize, o: *mut u8, ol: usize, s: DltCoreSettings) -> DltResult;
    fn dltbc1core_untransform(i: *const u8, il: usize, o: *mut u8, ol: usize, s: DltCoreSettings) -> DltResult;
    fn dltbc2core_transform(i: *const u8, il: usize, o: *mut u8, ol: usize, s: DltCoreSettings) -> DltResult;
    fn dltbc2core_untransform(i: *const u8, il: usize, o: *mut u8, ol: usize, s: DltCoreSettings) -> DltResult;
    fn dltbc3core_transform(i: *const u8, il: usize, o: *mut u8, ol: usize, s: DltCoreBc3Settings) -> DltResult;
    fn dltbc3core_untransform(i: *const u8, il: usize, o: *mut u8, ol: usize, s: DltCoreBc3Settings) -> DltResult;
    fn dltbc1core_transform_auto(
        d: *const u8, dl: usize, o: *mut u8, ol: usize, e: *const DltSizeEstimator, s: DltCoreAutoSettings,
        out: *mut DltCoreSettings,
    ) -> DltResult;
    fn dltbc2core_transform_auto(
        d: *const u8, dl: usize, o: *mut u8, ol: usize, e: *const DltSizeEstimator, s: DltCoreAutoSettings,
        out: *mut DltCoreSettings,
    ) -> DltResult;
    fn dltbc3core_transform_auto(
        d: *const u8, dl: usize, o: *mut u8, ol: usize, e: *const DltSizeEstimator, s: DltCoreAutoSettings,
        out: *mut DltCoreBc3Settings,
    ) -> DltResult;
    pub fn dltltu_new_size_estimator() -> *mut DltSizeEstimator;
    pub fn dltltu_free_size_estimator(e: *mut DltSizeEstimator);
    pub fn dltcuda_alloc_pinned(bytes: usize) -> *mut c_void;
    pub fn dltcuda_free_pinned(p: *mut c_void);
    pub fn dltcuda_set_device(device: i32);
}

#[derive(Debug, thiserror::Error, PartialEq, Eq)]
pub enum CudaTransformError {
    #[error("Invalid input length: {0}")]
    InvalidLength(usize),
    #[error("Output buffer too small: needed {needed}, got {actual}")]
    OutputBufferTooSmall { needed: usize, actual: usize },
    #[error("Size estimation failed")]
    SizeEstimationFailed,
    #[error("CUDA failure (core error code {0})")]
    Device(i32),
}

fn check(r: DltResult, needed: usize, actual: usize) -> Result<(), CudaTransformError> {
    match r.error_code {
        0 => Ok(()),
        5 => Err(CudaTransformError::InvalidLength(needed)),
        6 => Err(CudaTransformError::OutputBufferTooSmall { needed, actual }),
        7 => Err(CudaTransformError::SizeEstimationFailed),
        c => Err(CudaTransformError::Device(c)),
    }
}

fn core12(variant: YCoCgVariant, split: bool) -> DltCoreSettings {
    DltCoreSettings { split_colour_endpoints: split, decorrelation_mode: variant as u8 }
}

/// Same contract as `dxt_lossless_transform_bc1::transform_bc1_with_settings` (raw pointers, `len % 8 == 0`), but a
/// device failure is RETURNED: prefer this over the signature-compatible function below.
///
/// # Safety
/// As the reference function: both pointers valid for `len` bytes, non-overlapping.
pub unsafe fn try_transform_bc1_with_settings(
    input: *const u8, output: *mut u8, len: usize, s: Bc1TransformSettings,
) -> Result<(), CudaTransformError> {
    check(dltbc1core_transform(input, len, output, len, core12(s.decorrelation_mode, s.split_colour_endpoints)), len, len)
}

/// # Safety
/// As the reference function.
pub unsafe fn try_untransform_bc1_with_settings(
    input: *const u8, output: *mut u8, len: usize, s: Bc1TransformSettings,
) -> Result<(), CudaTransformError> {
    check(dltbc1core_untransform(input, len, output, len, core12(s.decorrelation_mode, s.split_colour_endpoints)), len, len)
}

/// Signature-compatible with the reference (`unsafe fn(*const u8, *mut u8, usize, Bc1TransformSettings)`, no return
/// value — transform_with_settings.rs:31).  The reference cannot fail here; a GPU can (no device, out of memory), and
/// the signature leaves no way to say so: the failure panics, which under the reference's `panic = "abort"` release
/// profile ends the process.  Callers that can handle an error use `try_transform_bc1_with_settings` or the `_safe` form.
///
/// # Safety
/// As the reference function: both pointers valid for `len` bytes, non-overlapping.
pub unsafe fn transform_bc1_with_settings(input: *const u8, output: *mut u8, len: usize, s: Bc1TransformSettings) {
    if let Err(e) = try_transform_bc1_with_settings(input, output, len, s) {
        panic!("dxt-lossless-transform-cuda: {e}");
    }
}

/// # Safety
/// As the reference function.
pub unsafe fn untransform_bc1_with_settings(input: *const u8, output: *mut u8, len: usize, s: Bc1TransformSettings) {
    if let Err(e) = try_untransform_bc1_with_settings(input, output, len, s) {
        panic!("dxt-lossless-transform-cuda: {e}");
    }
}

/// `transform_bc1_with_settings_safe` (safe/transform_with_settings.rs:88).
pub fn transform_bc1_with_settings_safe(i: &[u8], o: &mut [u8], s: Bc1TransformSettings) -> Result<(), CudaTransformError> {
    let r = unsafe {
        dltbc1core_transform(i.as_ptr(), i.len(), o.as_mut_ptr(), o.len(), core12(s.decorrelation_mode, s.split_colour_endpoints))
    };
    check(r, i.len(), o.len())
}

/// `untransform_bc1_with_settings_safe` (safe/transform_with_settings.rs:192).
pub fn untransform_bc1_with_settings_safe(i: &[u8], o: &mut [u8], s: Bc1TransformSettings) -> Result<(), CudaTransformError> {
    let r = unsafe {
        dltbc1core_untransform(i.as_ptr(), i.len(), o.as_mut_ptr(), o.len(), core12(s.decorrelation_mode, s.split_colour_endpoints))
    };
    check(r, i.len(), o.len())
}

pub fn transform_bc2_with_settings_safe(i: &[u8], o: &mut [u8], s: Bc2TransformSettings) -> Result<(), CudaTransformError> {
    let r = unsafe {
        dltbc2core_transform(i.as_ptr(), i.len(), o.as_mut_ptr(), o.len(), core12(s.decorrelation_mode, s.split_colour_endpoints))
    };
    check(r, i.len(), o.len())
}

pub fn untransform_bc2_with_settings_safe(i: &[u8], o: &mut [u8], s: Bc2TransformSettings) -> Result<(), CudaTransformError> {
    let r = unsafe {
        dltbc2core_untransform(i.as_ptr(), i.len(), o.as_mut_ptr(), o.len(), core12(s.decorrelation_mode, s.split_colour_endpoints))
    };
    check(r, i.len(), o.len())
}

fn core3(s: Bc3TransformSettings) -> DltCoreBc3Settings {
    DltCoreBc3Settings {
        split_alpha_endpoints: s.split_alpha_endpoints,
        split_colour_endpoints: s.split_colour_endpoints,
        decorrelation_mode: s.decorrelation_mode as u8,
    }
}

pub fn transform_bc3_with_settings_safe(i: &[u8], o: &mut [u8], s: Bc3TransformSettings) -> Result<(), CudaTransformError> {
    check(unsafe { dltbc3core_transform(i.as_ptr(), i.len(), o.as_mut_ptr(), o.len(), core3(s)) }, i.len(), o.len())
}

pub fn untransform_bc3_with_settings_safe(i: &[u8], o: &mut [u8], s: Bc3TransformSettings) -> Result<(), CudaTransformError> {
    check(unsafe { dltbc3core_untransform(i.as_ptr(), i.len(), o.as_mut_ptr(), o.len(), core3(s)) }, i.len(), o.len())
}

// ---- transform_bcN_auto: any SizeEstimationOperations goes through a C trampoline ------------------
unsafe extern "C" fn tramp_max<T: SizeEstimationOperations>(ctx: *mut c_void, len: usize, out: *mut usize) -> u32 {
    match (*(ctx as *const T)).max_compressed_size(len) {
        Ok(v) => {
            *out = v;
            0
        }
        Err(_) => 2,
    }
}

unsafe extern "C" fn tramp_est<T: SizeEstimationOperations>(
    ctx: *mut c_void, input: *const u8, len: usize, scratch: *mut u8, scratch_len: usize, out: *mut usize,
) -> u32 {
    match (*(ctx as *const T)).estimate_compressed_size(input, len, scratch, scratch_len) {
        Ok(v) => {
            *out = v;
            0
        }
        Err(_) => 3,
    }
}

/// `transform_bc1_auto_safe` (safe/transform_auto.rs:95).  Pass [`LtuOnGpu`] to run the whole search,
/// estimator included, on the device; any other estimator is called back with host memory.
pub fn transform_bc1_auto_safe<T: SizeEstimationOperations>(
    input: &[u8], output: &mut [u8], estimator: &T, use_all_decorrelation_modes: bool,
) -> Result<Bc1TransformSettings, CudaTransformError> {
    let est = DltSizeEstimator {
        context: estimator as *const T as *mut c_void,
        max_compressed_size: tramp_max::<T>,
        estimate_compressed_size: tramp_est::<T>,
    };
    let mut out = DltCoreSettings { split_colour_endpoints: true, decorrelation_mode: 1 };
    let r = unsafe {
        dltbc1core_transform_auto(
            input.as_ptr(), input.len(), output.as_mut_ptr(), output.len(), &est,
            DltCoreAutoSettings { use_all_modes: use_all_decorrelation_modes }, &mut out,
        )
    };
    check(r, input.len(), output.len())?;
    Ok(Bc1TransformSettings {
        decorrelation_mode: match out.decorrelation_mode {
            1 => YCoCgVariant::Variant1,
            2 => YCoCgVariant::Variant2,
            3 => YCoCgVariant::Variant3,
            _ => YCoCgVariant::None,
        },
        split_colour_endpoints: out.split_colour_endpoints,
    })
}

/// The LTU estimator whose callbacks are the library's own: `transform_bcN_auto` recognises it and
/// keeps the search on the GPU.
pub struct LtuOnGpu(*mut DltSizeEstimator);

impl LtuOnGpu {
    pub fn new() -> Self {
        Self(unsafe { dltltu_new_size_estimator() })
    }
    pub fn as_c(&self) -> *const DltSizeEstimator {
        self.0
    }
}

impl Drop for LtuOnGpu {
    fn drop(&mut self) {
        unsafe { dltltu_free_size_estimator(self.0) }
    }
}

// bc2 / bc3 auto wrappers are identical in shape (dltbc2core_transform_auto, dltbc3core_transform_auto).
#[allow(dead_code)]
fn _link_check() {
    let _ = (dltbc2core_transform_auto as usize, dltbc3core_transform_auto as usize);
}
