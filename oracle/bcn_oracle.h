/*
 * bcn_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain scalar C restatement of the BCn lossless-transform hot path of
 * Sewer56/dxt-lossless-transform (reference paths are relative to
 * /root/reference/src).  It exists to CHECK the CUDA product path; nothing in
 * the product (dxt_lossless_transform_b200/, include/) links, imports or calls
 * it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may use this library.
 *
 * Parity status
 *   - transform / untransform / YCoCg-R / auto search loops: restated from the
 *     reference's own scalar ground-truth files (generic.rs / portable32.rs) and
 *     pinned by the vectors in tests/golden (reference generators, the one
 *     golden byte vector of split_565_color_endpoints/tests.rs, the SURVEY §8c
 *     known answers) and exhaustive bijection checks.  The reference itself is
 *     Rust and cannot be compiled in this image (no cargo/rustc), so there is
 *     no oracle/_ref.
 *   - LTU size estimator (lossless-transform-utils 0.1.3, crates.io, NOT under
 *     /root/reference): **parity unpinned** — restated from the crate's
 *     published algorithm; every tunable lives in ltu_params.h.
 */
#ifndef BCN_ORACLE_H
#define BCN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Internal numbering of common/src/color_565/decorrelate.rs:72-84. */
enum { ORC_VARIANT_NONE = 0, ORC_VARIANT_1 = 1, ORC_VARIANT_2 = 2, ORC_VARIANT_3 = 3 };

/* Color565::decorrelate_ycocg_r_var{1,2,3} / recorrelate (decorrelate.rs:101-344). */
uint16_t orc_decorrelate(uint16_t v, int variant);
uint16_t orc_recorrelate(uint16_t v, int variant);

/* transform_bcN_with_settings / untransform_bcN_with_settings. len in bytes. */
void orc_bc1_transform(const uint8_t *in, uint8_t *out, size_t len, int variant, int split_colour);
void orc_bc1_untransform(const uint8_t *in, uint8_t *out, size_t len, int variant, int split_colour);
void orc_bc2_transform(const uint8_t *in, uint8_t *out, size_t len, int variant, int split_colour);
void orc_bc2_untransform(const uint8_t *in, uint8_t *out, size_t len, int variant, int split_colour);
void orc_bc3_transform(const uint8_t *in, uint8_t *out, size_t len, int variant, int split_alpha,
                       int split_colour);
void orc_bc3_untransform(const uint8_t *in, uint8_t *out, size_t len, int variant, int split_alpha,
                         int split_colour);

/* split_color_endpoints (common/src/transforms/split_565_color_endpoints/portable32.rs:18-64). */
void orc_split_color_endpoints(const uint8_t *in, uint8_t *out, size_t len_bytes);

/* lossless_transform_utils::match_estimator::estimate_num_lz_matches_fast (restated, unpinned). */
size_t orc_ltu_num_lz_matches(const uint8_t *data, size_t len);
/* The same with explicit parameters / process-wide parameters for everything that calls the estimator (see bcn_oracle.c). */
size_t orc_ltu_num_lz_matches_params(const uint8_t *data, size_t len, int hash_bits, int index_top, int group);
int orc_ltu_set_params(int hash_bits, int index_top, int group);
/* LosslessTransformUtilsSizeEstimation::estimate_compressed_size (ltu/src/lib.rs:67-119). */
size_t orc_ltu_estimate(const uint8_t *data, size_t len);

/* Estimator callback shaped like SizeEstimationOperations (api-common/src/estimate/mod.rs:24-64).
 * Returns 0 on success. */
typedef int (*orc_estimate_fn)(void *ctx, const uint8_t *data, size_t len, size_t *out_size);

/* transform_bcN_auto.  Writes the winner into *out_variant / *out_split_* and leaves `out`
 * holding the data transformed with it.  est == NULL selects the LTU restatement.
 * Returns 0 on success, the callback's non-zero code on estimator failure. */
int orc_bc1_transform_auto(const uint8_t *in, uint8_t *out, size_t len, int use_all_modes,
                           orc_estimate_fn est, void *ctx, int *out_variant, int *out_split_colour);
int orc_bc2_transform_auto(const uint8_t *in, uint8_t *out, size_t len, int use_all_modes,
                           orc_estimate_fn est, void *ctx, int *out_variant, int *out_split_colour);
int orc_bc3_transform_auto(const uint8_t *in, uint8_t *out, size_t len, int use_all_modes,
                           orc_estimate_fn est, void *ctx, int *out_variant, int *out_split_alpha,
                           int *out_split_colour);

/* Per-candidate estimates in test order (debug aid for parity tests): fills sizes[0..K). Returns K. */
int orc_bc1_auto_estimates(const uint8_t *in, uint8_t *scratch, size_t len, int use_all_modes,
                           size_t *sizes);
int orc_bc2_auto_estimates(const uint8_t *in, uint8_t *scratch, size_t len, int use_all_modes,
                           size_t *sizes);
int orc_bc3_auto_estimates(const uint8_t *in, uint8_t *scratch, size_t len, int use_all_modes,
                           size_t *sizes);

/* Reference test-data generators (test_prelude.rs of each core crate). */
void orc_generate_bc1_test_data(uint8_t *out, size_t num_blocks);
void orc_generate_bc2_test_data(uint8_t *out, size_t num_blocks);
void orc_generate_bc3_test_data(uint8_t *out, size_t num_blocks);

/* Multi-threaded drivers for the CPU baseline: split [0,len) into `threads` contiguous block
 * ranges, each thread writing its slice of every output stream (same bytes as the single call).
 * format: 1/2/3.  direction: 0 transform, 1 untransform. */
void orc_bcn_run_mt(int format, int direction, const uint8_t *in, uint8_t *out, size_t len,
                    int variant, int split_alpha, int split_colour, int threads);

/* Block range [b0, b1) only: reads/writes exactly the slices of every stream that range owns. */
void orc_bcn_run_range(int format, int direction, const uint8_t *in, uint8_t *out, size_t len,
                       int variant, int split_alpha, int split_colour, size_t b0, size_t b1);

/* experimental::normalize_blocks (BC1); ColorNormalizationMode::all_values() order (normalize.rs:487-500). */
enum { ORC_NORM_NONE = 0, ORC_NORM_COLOR0_ONLY = 1, ORC_NORM_REPLICATE = 2 };
void orc_bc1_normalize_blocks(const uint8_t *in, uint8_t *out, size_t len, int mode);
int orc_bc1_normalize_blocks_all_modes(const uint8_t *in, uint8_t *out_none, uint8_t *out_color0,
                                       uint8_t *out_replicate, size_t len);
void orc_bc1_normalize_split_blocks_in_place(uint8_t *colors, uint8_t *indices, size_t num_blocks, int mode);
void orc_bc1_transform_with_normalize_blocks(const uint8_t *in, uint8_t *out, size_t len, int norm, int variant,
                                             int split);
int orc_bc1_transform_auto_with_normalization(const uint8_t *in, uint8_t *out, size_t len, int use_all,
                                              orc_estimate_fn est, void *ctx, int *out_norm, int *out_variant,
                                              int *out_split);

/* 1 when orc_bcn_run_mt / orc_bcn_run_range take the explicit AVX2 BC1 path on this host. */
int orc_cpu_baseline_uses_avx2(void);


#ifdef __cplusplus
}
#endif
#endif
