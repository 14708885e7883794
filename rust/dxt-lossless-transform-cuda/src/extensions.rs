//! The two optional pieces around the hot path, with the reference's names.
//!
//! SOURCE ONLY, like the rest of the crate (no Rust toolchain in the build image).
//!
//! * [`ZStandardOnHostThreads`] — `ZStandardSizeEstimation` of crate dxt-lossless-transform-zstd
//!   (extensions/compressors/dxt-lossless-transform-zstd/src/lib.rs:54-140) as a `DltSizeEstimator` made by
//!   `dltzstd_new_size_estimator`.  `transform_bcN_auto` recognises it: candidates are transformed on the GPU and
//!   compressed concurrently, one host thread per candidate.  Estimates are real zstd sizes with the reference's
//!   parameters (magicless, no content size / checksum / dict id).
//! * `experimental::normalize_blocks` (core/dxt-lossless-transform-bc1/src/experimental/normalize_blocks/):
//!   `normalize_blocks`, `normalize_blocks_all_modes`, `normalize_split_blocks_in_place`,
//!   `transform_bc1_with_normalize_blocks`, `transform_bc1_auto_with_normalization` over the `dltcuda_bc1_*` symbols;
//!   in the fused entry points normalization happens inside the transform kernel.
use crate::{CudaTransformError, DltSizeEstimator};
use dxt_lossless_transform_bc1::experimental::normalize_blocks::ColorNormalizationMode;
use dxt_lossless_transform_bc1::experimental::normalize_blocks::Bc1TransformDetailsWithNormalization;
use dxt_lossless_transform_common::color_565::YCoCgVariant;

extern "C" {
    fn dltzstd_new_size_estimator(compression_level: i32) -> *mut DltSizeEstimator;
    fn dltzstd_free_size_estimator(e: *mut DltSizeEstimator);
    fn dltzstd_version_number() -> u32;
    fn dltcuda_bc1_normalize_blocks(input: *const u8, output: *mut u8, len: usize, mode: i32) -> i32;
    fn dltcuda_bc1_normalize_blocks_all_modes(
        input: *const u8, out_none: *mut u8, out_color0_only: *mut u8, out_replicate_color: *mut u8, len: usize,
        any_normalized: *mut bool,
    ) -> i32;
    fn dltcuda_bc1_normalize_split_blocks_in_place(colors: *mut u8, indices: *mut u8, num_blocks: usize, mode: i32) -> i32;
    fn dltcuda_bc1_transform_with_normalize_blocks(
        input: *const u8, output: *mut u8, len: usize, normalization_mode: i32, decorrelation_mode: u8,
        split_colour_endpoints: bool,
    ) -> i32;
    fn dltcuda_bc1_transform_auto_with_normalization(
        input: *const u8, output: *mut u8, len: usize, use_all_modes: bool, out_normalization_mode: *mut i32,
        out_decorrelation_mode: *mut u8, out_split_colour_endpoints: *mut bool, out_estimates: *mut usize,
    ) -> i32;
}

/// zstd estimator whose compression runs on host threads in parallel with the GPU search.
pub struct ZStandardOnHostThreads(*mut DltSizeEstimator);

impl ZStandardOnHostThreads {
    /// `ZStandardSizeEstimation::new` (lib.rs:60-69): `None` for a level outside 1..=22 or when no libzstd can be loaded.
    pub fn new(compression_level: i32) -> Option<Self> {
        let p = unsafe { dltzstd_new_size_estimator(compression_level) };
        if p.is_null() { None } else { Some(Self(p)) }
    }
    pub fn new_fast() -> Option<Self> { Self::new(1) }
    pub fn new_default() -> Option<Self> { Self::new(3) }
    pub fn new_best() -> Option<Self> { Self::new(22) }
    /// `ZSTD_versionNumber()` of the library bound at run time; sizes equal the reference's when this is 10507.
    pub fn library_version() -> u32 { unsafe { dltzstd_version_number() } }
    pub fn as_c(&self) -> *const DltSizeEstimator { self.0 }
}

impl Drop for ZStandardOnHostThreads {
    fn drop(&mut self) { unsafe { dltzstd_free_size_estimator(self.0) } }
}

fn mode_code(m: ColorNormalizationMode) -> i32 {
    match m {
        ColorNormalizationMode::None => 0,
        ColorNormalizationMode::Color0Only => 1,
        ColorNormalizationMode::ReplicateColor => 2,
    }
}

fn device(rc: i32) -> Result<(), CudaTransformError> {
    if rc == 0 { Ok(()) } else { Err(CudaTransformError::Device(rc)) }
}

/// `normalize_blocks` (normalize.rs:38).  `len` must be a multiple of 8; input and output may be the same buffer.
///
/// # Safety
/// Both pointers must be valid for `len` bytes.
pub unsafe fn normalize_blocks(input: *const u8, output: *mut u8, len: usize, mode: ColorNormalizationMode) {
    device(dltcuda_bc1_normalize_blocks(input, output, len, mode_code(mode))).expect("CUDA failure in normalize_blocks");
}

/// `normalize_blocks_all_modes` (normalize.rs:417): returns whether any block was normalized.
///
/// # Safety
/// `input` and the three outputs must be valid for `len` bytes.
pub unsafe fn normalize_blocks_all_modes(input: *const u8, outputs: &[*mut u8; 3], len: usize) -> bool {
    let mut any = false;
    device(dltcuda_bc1_normalize_blocks_all_modes(input, outputs[0], outputs[1], outputs[2], len, &mut any))
        .expect("CUDA failure in normalize_blocks_all_modes");
    any
}

/// `normalize_split_blocks_in_place` (normalize.rs:286).
///
/// # Safety
/// `colors` and `indices` must be valid for `4 * num_blocks` bytes each.
pub unsafe fn normalize_split_blocks_in_place(colors: *mut u8, indices: *mut u8, num_blocks: usize, mode: ColorNormalizationMode) {
    device(dltcuda_bc1_normalize_split_blocks_in_place(colors, indices, num_blocks, mode_code(mode)))
        .expect("CUDA failure in normalize_split_blocks_in_place");
}

/// `transform_bc1_with_normalize_blocks` (transform.rs:65) — one fused pass on the GPU.
pub fn transform_bc1_with_normalize_blocks_safe(
    input: &[u8], output: &mut [u8], details: Bc1TransformDetailsWithNormalization,
) -> Result<(), CudaTransformError> {
    if input.len() % 8 != 0 {
        return Err(CudaTransformError::InvalidLength(input.len()));
    }
    if output.len() < input.len() {
        return Err(CudaTransformError::OutputBufferTooSmall { needed: input.len(), actual: output.len() });
    }
    device(unsafe {
        dltcuda_bc1_transform_with_normalize_blocks(
            input.as_ptr(), output.as_mut_ptr(), input.len(), mode_code(details.color_normalization_mode),
            details.decorrelation_mode as u8, details.split_colour_endpoints,
        )
    })
}

/// `transform_bc1_auto_with_normalization` (transform.rs:222) with the LTU-semantics estimator on the GPU.
pub fn transform_bc1_auto_with_normalization_safe(
    input: &[u8], output: &mut [u8], use_all_decorrelation_modes: bool,
) -> Result<Bc1TransformDetailsWithNormalization, CudaTransformError> {
    if input.len() % 8 != 0 {
        return Err(CudaTransformError::InvalidLength(input.len()));
    }
    if output.len() < input.len() {
        return Err(CudaTransformError::OutputBufferTooSmall { needed: input.len(), actual: output.len() });
    }
    let (mut norm, mut var, mut split) = (0i32, 0u8, false);
    device(unsafe {
        dltcuda_bc1_transform_auto_with_normalization(
            input.as_ptr(), output.as_mut_ptr(), input.len(), use_all_decorrelation_modes, &mut norm, &mut var, &mut split,
            core::ptr::null_mut(),
        )
    })?;
    Ok(Bc1TransformDetailsWithNormalization {
        color_normalization_mode: match norm {
            1 => ColorNormalizationMode::Color0Only,
            2 => ColorNormalizationMode::ReplicateColor,
            _ => ColorNormalizationMode::None,
        },
        decorrelation_mode: match var {
            1 => YCoCgVariant::Variant1,
            2 => YCoCgVariant::Variant2,
            3 => YCoCgVariant::Variant3,
            _ => YCoCgVariant::None,
        },
        split_colour_endpoints: split,
    })
}
