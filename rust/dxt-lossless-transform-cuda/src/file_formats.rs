//! File-format layer over the C ABI of include/dxt_lossless_transform_file_formats.h and
//! include/dxt_lossless_transform_dds.h — SOURCE ONLY (no Rust toolchain in the build image).
//!
//! Drop-ins for (paths relative to the reference's `src/`):
//!   api/dxt-lossless-transform-file-formats-api/src/handlers/dispatch.rs   dispatch_transform / dispatch_untransform
//!   extensions/file-formats/dxt-lossless-transform-dds/src/handler/file_format_handler.rs   DdsHandler
//! The header value returned / consumed is the reference's `TransformHeader` (u32, little endian), so files written
//! through this crate are untransformed by the stock CPU crates and vice versa.
use core::ffi::c_void;

#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct DltffResult {
    pub error_code: i32, // DltffErrorCode
    pub detail_a: usize,
    pub detail_b: usize,
}

/// dds/parse_dds.rs:36-42
#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct DdsInfo {
    pub format: u8, // DdsFormat
    pub data_offset: u8,
    pub data_length: u32,
}

#[repr(C)]
pub struct DltddsFile {
    pub input: *const u8,
    pub input_len: usize,
    pub output: *mut u8,
    pub output_len: usize,
}

extern "C" {
    // the reference's own exports (dds/exports.rs)
    pub fn is_dds(ptr: *const u8, len: usize) -> bool;
    pub fn parse_dds(ptr: *const u8, len: usize) -> DdsInfo;

    fn dltff_new_TransformBundle() -> *mut c_void;
    fn dltff_TransformBundle_default_all() -> *mut c_void;
    fn dltff_free_TransformBundle(bundle: *mut c_void);
    fn dltff_TransformBundle_with_bc1_manual(bundle: *mut c_void, builder: *const c_void) -> DltffResult;
    fn dltff_TransformBundle_with_bc1_auto(bundle: *mut c_void, builder: *const c_void) -> DltffResult;
    fn dltff_TransformBundle_with_bc2_manual(bundle: *mut c_void, builder: *const c_void) -> DltffResult;
    fn dltff_TransformBundle_with_bc2_auto(bundle: *mut c_void, builder: *const c_void) -> DltffResult;
    fn dltff_dispatch_transform(format: i32, input: *const u8, input_len: usize, output: *mut u8, output_len: usize,
                                bundle: *const c_void, out_header: *mut u32) -> DltffResult;
    fn dltff_dispatch_untransform(header: u32, input: *const u8, input_len: usize, output: *mut u8,
                                  output_len: usize) -> DltffResult;
    fn dltdds_transform_bundle(input: *const u8, input_len: usize, output: *mut u8, output_len: usize,
                               bundle: *const c_void) -> DltffResult;
    fn dltdds_untransform(input: *const u8, input_len: usize, output: *mut u8, output_len: usize) -> DltffResult;
    fn dltdds_transform_bundle_batch(files: *const DltddsFile, count: usize, bundle: *const c_void,
                                     results: *mut DltffResult, devices: *const i32, num_devices: i32) -> i32;
    fn dltdds_untransform_batch(files: *const DltddsFile, count: usize, results: *mut DltffResult,
                                devices: *const i32, num_devices: i32) -> i32;
}

/// Owned `DltffTransformBundle`.
pub struct TransformBundle(*mut c_void);

impl TransformBundle {
    pub fn new() -> Self { Self(unsafe { dltff_new_TransformBundle() }) }
    pub fn default_all() -> Self { Self(unsafe { dltff_TransformBundle_default_all() }) }
    /// `builder` is a `Dltbc1ManualTransformBuilder*` of the stable C API; its settings are copied.
    pub unsafe fn with_bc1_manual(self, builder: *const c_void) -> Self { dltff_TransformBundle_with_bc1_manual(self.0, builder); self }
    pub unsafe fn with_bc1_auto(self, builder: *const c_void) -> Self { dltff_TransformBundle_with_bc1_auto(self.0, builder); self }
    pub unsafe fn with_bc2_manual(self, builder: *const c_void) -> Self { dltff_TransformBundle_with_bc2_manual(self.0, builder); self }
    pub unsafe fn with_bc2_auto(self, builder: *const c_void) -> Self { dltff_TransformBundle_with_bc2_auto(self.0, builder); self }
}
impl Drop for TransformBundle {
    fn drop(&mut self) { unsafe { dltff_free_TransformBundle(self.0) } }
}

fn check(r: DltffResult) -> Result<(), DltffResult> { if r.error_code == 0 { Ok(()) } else { Err(r) } }

/// handlers/dispatch.rs:131 — returns the `TransformHeader` to embed.
pub fn dispatch_transform(format: i32, input: &[u8], output: &mut [u8], bundle: &TransformBundle) -> Result<u32, DltffResult> {
    let mut header = 0u32;
    check(unsafe { dltff_dispatch_transform(format, input.as_ptr(), input.len(), output.as_mut_ptr(), output.len(), bundle.0, &mut header) })?;
    Ok(header)
}

/// handlers/dispatch.rs:41
pub fn dispatch_untransform(header: u32, input: &[u8], output: &mut [u8]) -> Result<(), DltffResult> {
    check(unsafe { dltff_dispatch_untransform(header, input.as_ptr(), input.len(), output.as_mut_ptr(), output.len()) })
}

/// `FileFormatHandler` for DDS (handler/file_format_handler.rs:17-145).
pub struct DdsHandler;

impl DdsHandler {
    pub fn transform_bundle(&self, input: &[u8], output: &mut [u8], bundle: &TransformBundle) -> Result<(), DltffResult> {
        check(unsafe { dltdds_transform_bundle(input.as_ptr(), input.len(), output.as_mut_ptr(), output.len(), bundle.0) })
    }
    pub fn untransform(&self, input: &[u8], output: &mut [u8]) -> Result<(), DltffResult> {
        check(unsafe { dltdds_untransform(input.as_ptr(), input.len(), output.as_mut_ptr(), output.len()) })
    }
    /// A directory at once: one pinned copy pipeline per GPU instead of one rayon task per file
    /// (tools/dxt-lossless-transform-cli/src/commands/transform/mod.rs:154-176).
    pub fn transform_bundle_batch(&self, files: &mut [(&[u8], &mut [u8])], bundle: &TransformBundle, devices: &[i32]) -> Vec<DltffResult> {
        let raw: Vec<DltddsFile> = files.iter_mut().map(|(i, o)| DltddsFile { input: i.as_ptr(), input_len: i.len(), output: o.as_mut_ptr(), output_len: o.len() }).collect();
        let mut results = vec![DltffResult { error_code: 0, detail_a: 0, detail_b: 0 }; raw.len()];
        unsafe { dltdds_transform_bundle_batch(raw.as_ptr(), raw.len(), bundle.0, results.as_mut_ptr(), if devices.is_empty() { core::ptr::null() } else { devices.as_ptr() }, devices.len() as i32) };
        results
    }
    pub fn untransform_batch(&self, files: &mut [(&[u8], &mut [u8])], devices: &[i32]) -> Vec<DltffResult> {
        let raw: Vec<DltddsFile> = files.iter_mut().map(|(i, o)| DltddsFile { input: i.as_ptr(), input_len: i.len(), output: o.as_mut_ptr(), output_len: o.len() }).collect();
        let mut results = vec![DltffResult { error_code: 0, detail_a: 0, detail_b: 0 }; raw.len()];
        unsafe { dltdds_untransform_batch(raw.as_ptr(), raw.len(), results.as_mut_ptr(), if devices.is_empty() { core::ptr::null() } else { devices.as_ptr() }, devices.len() as i32) };
        results
    }
}
