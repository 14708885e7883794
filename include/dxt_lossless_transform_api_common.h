/*
 * dxt_lossless_transform_api_common.h — shared C types of the drop-in boundary.
 *
 * Replaces the cbindgen output of the reference crate dxt-lossless-transform-api-common
 * (paths relative to /root/reference/src):
 *   DltSizeEstimator + callback types   api/dxt-lossless-transform-api-common/src/c_api/size_estimation.rs:17-52
 *   YCoCgVariant (STABLE numbering)     api/dxt-lossless-transform-api-common/src/reexports/color_565.rs:65-85
 * Layouts and values are identical to the reference's; only enumerator SPELLING is prefixed with
 * the type name so that the stable and the core headers can be included together (the reference's
 * own generated headers cannot — they reuse names).
 */
#ifndef DXT_LOSSLESS_TRANSFORM_API_COMMON_H
#define DXT_LOSSLESS_TRANSFORM_API_COMMON_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 0 on success, non-zero error code on failure (size_estimation.rs:6-41). */
typedef uint32_t (*DltMaxCompressedSizeFn)(void *context, size_t len_bytes, size_t *out_size);
typedef uint32_t (*DltEstimateCompressedSizeFn)(void *context, const uint8_t *input_ptr,
                                                size_t len_bytes, uint8_t *output_ptr,
                                                size_t output_len, size_t *out_size);

/* size_estimation.rs:44-52 (repr(C)). */
typedef struct DltSizeEstimator {
  void *context;
  DltMaxCompressedSizeFn max_compressed_size;
  DltEstimateCompressedSizeFn estimate_compressed_size;
} DltSizeEstimator;

/* Stable API numbering, repr(u8) (reexports/color_565.rs:65-85). */
enum {
  YCoCgVariant_Variant1 = 0,
  YCoCgVariant_Variant2 = 1,
  YCoCgVariant_Variant3 = 2,
  YCoCgVariant_None = 3,
};
typedef uint8_t YCoCgVariant;

/* Internal numbering used by the CORE crates' C ABI, repr(u8)
 * (core/dxt-lossless-transform-common/src/color_565/decorrelate.rs:72-84). */
enum {
  DltCoreYCoCgVariant_None = 0,
  DltCoreYCoCgVariant_Variant1 = 1,
  DltCoreYCoCgVariant_Variant2 = 2,
  DltCoreYCoCgVariant_Variant3 = 3,
};
typedef uint8_t DltCoreYCoCgVariant;

#ifdef __cplusplus
}
#endif
#endif
