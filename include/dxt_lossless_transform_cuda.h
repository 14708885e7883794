/*
 * dxt_lossless_transform_cuda.h — ADDITIVE device-side entry points of libdxt_lossless_transform_cuda.so.
 *
 * The reference is a CPU library: its boundary is synchronous calls on host buffers, and the
 * headers dxt_lossless_transform_bc{1,2}_{api,core}.h / _bc3_core.h / _ltu.h reproduce exactly that.
 * The functions here exist because the data can already live in HBM: they take CUDA device
 * pointers, a block range for multi-GPU sharding, or hand out page-locked host memory (the role the
 * reference's allocate_align_64 / allocate_cache_line_aligned play for its SIMD paths,
 * core/dxt-lossless-transform-common/src/allocate.rs).
 *
 * Settings use the INTERNAL YCoCg numbering (DltCoreYCoCgVariant_*).
 */
#ifndef DXT_LOSSLESS_TRANSFORM_CUDA_H
#define DXT_LOSSLESS_TRANSFORM_CUDA_H

#include "dxt_lossless_transform_api_common.h"
#include "dxt_lossless_transform_bc1_api.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum DltcudaStatus {
  DltcudaStatus_Ok = 0,
  DltcudaStatus_InvalidLength = 1,   /* len not a multiple of the block size / range out of bounds */
  DltcudaStatus_InvalidSettings = 2,
  DltcudaStatus_CudaError = 3,       /* see dltcuda_last_error() */
  DltcudaStatus_NullPointer = 4,
  DltcudaStatus_OutOfMemory = 5,
} DltcudaStatus;

typedef struct DltcudaSettings {
  uint8_t format;                         /* 1 = BC1, 2 = BC2, 3 = BC3 */
  DltCoreYCoCgVariant decorrelation_mode; /* internal numbering */
  bool split_alpha_endpoints;             /* BC3 only */
  bool split_colour_endpoints;
} DltcudaSettings;

/* ---- device selection / diagnostics -------------------------------------------------------------- */
int dltcuda_device_count(void);
/* Device used by THIS host thread for every entry point of the library (-1 = the thread's current
 * CUDA device, the default). */
void dltcuda_set_device(int device);
/* The library keeps its per-device working set between calls (device buffers sized to the largest payload seen, the
 * estimator's scratch, pinned staging slots): this frees the buffers of every idle pooled context and returns the number
 * of bytes released.  Safe at any time; buffers are re-allocated on demand. */
size_t dltcuda_release_cached_memory(void);
/* Text of the last CUDA error seen by this thread. */
const char *dltcuda_last_error(void);
/* Kernels launched by the library in this process so far. */
uint64_t dltcuda_kernel_launch_count(void);

/* ---- page-locked host buffers ------------------------------------------------------------------- */
/* Host buffers from here are copied to/from the device directly and asynchronously by the host-
 * pointer entry points; any other host memory is staged through an internal pinned ring. */
void *dltcuda_alloc_pinned(size_t bytes);
void dltcuda_free_pinned(void *ptr);

/* ---- whole payload, device resident ------------------------------------------------------------ */
/* d_input / d_output: device pointers to `len` bytes; d_output receives the reference's single-buffer
 * layout (transform_bcN_with_settings: core/dxt-lossless-transform-bc1/src/transform/
 * transform_with_settings.rs:31, bc2 :30, bc3 :32).  Buffers must not overlap.  Asynchronous on
 * `stream` (a cudaStream_t; NULL = the legacy default stream).  Any alignment works; 16-byte aligned
 * block pointers and naturally aligned streams take the tiled kernels. */
int dltcuda_transform_device(const uint8_t *d_input, uint8_t *d_output, size_t len,
                             DltcudaSettings settings, void *stream);
/* untransform_bcN_with_settings (bc1 :92, bc2 :93, bc3 :162). */
int dltcuda_untransform_device(const uint8_t *d_input, uint8_t *d_output, size_t len,
                               DltcudaSettings settings, void *stream);

/* ---- block-range shards (multi-GPU) -------------------------------------------------------------- */
/* Blocks [first_block, first_block + num_blocks) of a payload of total_blocks blocks.
 *   transform  : d_blocks = the shard's first block; d_streams_base = base of the FULL transformed
 *                image (reference layout for total_blocks); only this shard's slice of every stream
 *                is written.
 *   untransform: the mirror.
 * Shards of one payload share nothing but (total_blocks, first_block): there is no collective. */
int dltcuda_transform_device_range(const uint8_t *d_blocks, uint8_t *d_streams_base,
                                   size_t total_blocks, size_t first_block, size_t num_blocks,
                                   DltcudaSettings settings, void *stream);
int dltcuda_untransform_device_range(const uint8_t *d_streams_base, uint8_t *d_blocks,
                                     size_t total_blocks, size_t first_block, size_t num_blocks,
                                     DltcudaSettings settings, void *stream);
/* Explicit per-stream device pointers, in the stream order of the layout (BC1: [c0c1 | c0, c1], idx;
 * BC2: alpha, [c0c1 | c0, c1], idx; BC3: [a0a1 | a0, a1], aidx, [c0c1 | c0, c1], idx), each already
 * pointing at this range's first element.  A rank that holds only its own shard uses these. */
int dltcuda_transform_device_streams(const uint8_t *d_blocks, uint8_t *const *d_streams,
                                     size_t num_blocks, DltcudaSettings settings, void *stream);
int dltcuda_untransform_device_streams(const uint8_t *const *d_streams, uint8_t *d_blocks,
                                       size_t num_blocks, DltcudaSettings settings, void *stream);
/* Layout introspection: number of streams, and bytes per block of stream k, for these settings. */
int dltcuda_stream_count(DltcudaSettings settings);
int dltcuda_stream_width(DltcudaSettings settings, int k);
/* First block of shard `shard` out of `num_shards` (shard == num_shards -> total_blocks): the host-
 * side prefix of shard offsets, rounded to the kernel tile. */
size_t dltcuda_shard_first_block(int format, size_t total_blocks, int shard, int num_shards);

/* ---- split_color_endpoints ----------------------------------------------------------------------- */
/* split_color_endpoints (core/dxt-lossless-transform-common/src/transforms/split_565_color_endpoints/mod.rs,
 * portable32.rs:18-64): [c0 c1] x n -> c0 x n | c1 x n; len_bytes = 4n.  Device pointers + stream, or
 * host pointers (synchronous). */
int dltcuda_split_color_endpoints_device(const uint8_t *d_colors, uint8_t *d_colors_out,
                                         size_t len_bytes, void *stream);
int dltcuda_split_color_endpoints(const uint8_t *colors, uint8_t *colors_out, size_t len_bytes);

/* ---- batches of host payloads ---------------------------------------------------------------------- */
/* The reference's CLI transforms a directory one file per rayon task
 * (tools/dxt-lossless-transform-cli/src/commands/transform/mod.rs:154-176).  Here a batch of independent
 * payloads (any mix of formats and settings; host pointers, same rules as the with-settings calls)
 * goes through one pipeline per device: chunks of consecutive payloads overlap, one wait at the end.
 * untransform = false: transform_bcN_with_settings on every payload; true: untransform_...
 * Payloads of up to 4 MiB (8 MiB for a single-payload call) in page-locked memory (dltcuda_alloc_pinned, or anything cudaHostRegister-ed) are not
 * copied at all: all of them that share a settings combination are processed by ONE kernel launch over the mapped
 * host memory — a directory of thousands of small textures read into a pinned pool moves at ~30 GB/s. */
typedef struct DltcudaPayload {
  const uint8_t *input;
  uint8_t *output;
  size_t len;
  DltcudaSettings settings;
} DltcudaPayload;
int dltcuda_transform_batch(const DltcudaPayload *payloads, size_t count, bool untransform);
/* Same, dealt out over several GPUs of one box at payload granularity (whole payloads, least-loaded
 * device first; one host thread per device; nothing is exchanged between devices). */
int dltcuda_transform_batch_multi_gpu(const DltcudaPayload *payloads, size_t count, bool untransform,
                                      const int *devices, int num_devices);

/* ---- estimator / best-settings search, device resident ------------------------------------------ */
/* LTU-semantics estimate (dxt_lossless_transform_ltu.h) of `len` device bytes.  Synchronous. */
int dltcuda_ltu_estimate_device(const uint8_t *d_data, size_t len, size_t *out_size);
/* The estimator restates `estimate_num_lz_matches_fast` of the third-party crate lossless-transform-utils 0.1.3, whose
 * source is not in the reference tree (PARITY UNPINNED).  The parts of the restatement that could not
 * be checked are run-time parameters, process-wide, so that pinning parity later is a call, not a redesign:
 *   hash_bits 12..17 (table of 1 << hash_bits entries; default 16)
 *   index_from_top_bits: true  -> index = (key * 0x9E3779B1) >> (32 - hash_bits)   (default)
 *                        false -> index = (key * 0x9E3779B1) & ((1 << hash_bits) - 1)
 *   group 4 or 1: positions per loop iteration (all compares of an iteration precede its updates; default 4)
 * Returns DltcudaStatus_Ok, or DltcudaStatus_InvalidSettings (nothing changed) for an unsupported combination.
 * Must not be called while another thread is inside an estimating call. */
int dltcuda_ltu_set_params(int hash_bits, bool index_from_top_bits, int group);
void dltcuda_ltu_get_params(int *hash_bits, bool *index_from_top_bits, int *group);
/* transform_bcN_auto (bc1 transform_auto.rs:200, bc2 :196, bc3 :196) with the LTU estimator, all on
 * the device.  out_estimates (optional, >= 16 entries) receives the per-candidate estimates in the
 * reference's test order.  Synchronous; d_output holds the winner's transform on return. */
int dltcuda_transform_auto_device(int format, const uint8_t *d_input, uint8_t *d_output, size_t len,
                                  bool use_all_modes, DltcudaSettings *out_settings,
                                  size_t *out_estimates);
/* The same search for a BATCH of independent host payloads (a directory of textures; BASELINE config "determine-best
 * over a batch of chunks"): one upload per payload, ONE set of estimator launches for all candidates of all payloads,
 * one download per payload.  jobs[i].out_settings / .status are written; returns the first failing status (or Ok). */
typedef struct DltcudaAutoJob {
  uint8_t format; /* 1 = BC1, 2 = BC2, 3 = BC3 */
  const uint8_t *input;
  uint8_t *output;
  size_t len;
  DltcudaSettings out_settings;
  int32_t status; /* DltcudaStatus */
} DltcudaAutoJob;
int dltcuda_transform_auto_batch(DltcudaAutoJob *jobs, size_t count, bool use_all_modes);
/* Same, payloads dealt out whole over several GPUs of one box (one host thread per device, nothing exchanged). */
int dltcuda_transform_auto_batch_multi_gpu(DltcudaAutoJob *jobs, size_t count, bool use_all_modes,
                                           const int *devices, int num_devices);
/* The candidate order of that search (FAST_/COMPREHENSIVE_TEST_ORDER, bc1 settings.rs:81-98,
 * bc3 settings.rs:91-121).  `out` needs 16 entries; returns the count. */
int dltcuda_auto_candidates(int format, bool use_all_modes, DltcudaSettings *out);

/* ---- experimental: BC1 block normalization -------------------------------------------------------- */
/* core/dxt-lossless-transform-bc1/src/experimental/normalize_blocks/ (Rust-only and experimental in the reference):
 * solid-colour and fully transparent blocks get one canonical representation before the transform (visually lossless,
 * NOT bit-lossless).  normalize.rs:38 normalize_blocks, :286 normalize_split_blocks_in_place, :417
 * normalize_blocks_all_modes; transform.rs:65 transform_bc1_with_normalize_blocks, :222
 * transform_bc1_auto_with_normalization.  In the fused transform entry points the normalization happens inside the
 * transform kernel: one pass over the data. */
typedef enum DltcudaColorNormalizationMode { /* ColorNormalizationMode::all_values() order, normalize.rs:487-500 */
  DltcudaColorNormalizationMode_None = 0,
  DltcudaColorNormalizationMode_Color0Only = 1,
  DltcudaColorNormalizationMode_ReplicateColor = 2,
} DltcudaColorNormalizationMode;
int dltcuda_bc1_normalize_blocks(const uint8_t *input, uint8_t *output, size_t len, int mode);
int dltcuda_bc1_normalize_blocks_device(const uint8_t *d_input, uint8_t *d_output, size_t len, int mode,
                                        void *stream);
int dltcuda_bc1_normalize_blocks_all_modes(const uint8_t *input, uint8_t *out_none, uint8_t *out_color0_only,
                                           uint8_t *out_replicate_color, size_t len, bool *any_normalized);
int dltcuda_bc1_normalize_split_blocks_in_place(uint8_t *colors, uint8_t *indices, size_t num_blocks,
                                                int mode);
int dltcuda_bc1_transform_with_normalize_blocks(const uint8_t *input, uint8_t *output, size_t len,
                                                int normalization_mode,
                                                DltCoreYCoCgVariant decorrelation_mode,
                                                bool split_colour_endpoints);
int dltcuda_bc1_transform_with_normalize_blocks_device(const uint8_t *d_input, uint8_t *d_output, size_t len,
                                                       int normalization_mode,
                                                       DltCoreYCoCgVariant decorrelation_mode,
                                                       bool split_colour_endpoints, void *stream);
/* LTU estimator on the GPU; out_estimates optional (>= 24 entries, normalization mode outermost). */
int dltcuda_bc1_transform_auto_with_normalization(const uint8_t *input, uint8_t *output, size_t len,
                                                  bool use_all_modes, int *out_normalization_mode,
                                                  DltCoreYCoCgVariant *out_decorrelation_mode,
                                                  bool *out_split_colour_endpoints, size_t *out_estimates);

/* Reads back the settings a manual builder holds (Bc1ManualTransformBuilder::get_settings is
 * Rust-only in the reference).  out_mode uses the STABLE numbering.  Works for dltbc1_ and dltbc2_
 * builders.  Returns non-zero on a null builder. */
int dltcuda_ManualTransformBuilder_GetSettings(const void *builder, YCoCgVariant *out_mode,
                                               bool *out_split_colour_endpoints);

#ifdef __cplusplus
}
#endif
#endif
