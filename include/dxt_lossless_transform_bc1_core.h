/*
 * dxt_lossless_transform_bc1_core.h — the "unstable" core C ABI for BC1 (drop-in for the cbindgen
 * header of crate dxt-lossless-transform-bc1, feature c-exports).  file:line relative to
 * /root/reference/src/core/dxt-lossless-transform-bc1/src/c_api.
 *
 * NOTE the reference reuses the type names Dltbc1Result / Dltbc1ErrorCode / Dltbc1TransformSettings
 * for DIFFERENT layouts and values in this crate and in the -api crate.  Here the core types carry a
 * "Core" infix so both headers can be included; the layouts are the core crate's.
 */
#ifndef DXT_LOSSLESS_TRANSFORM_BC1_CORE_H
#define DXT_LOSSLESS_TRANSFORM_BC1_CORE_H

#include "dxt_lossless_transform_api_common.h"

#ifdef __cplusplus
extern "C" {
#endif

/* transform_auto.rs:37-58 */
typedef enum Dltbc1CoreErrorCode {
  Dltbc1CoreErrorCode_Success = 0,
  Dltbc1CoreErrorCode_NullDataPointer = 1,
  Dltbc1CoreErrorCode_NullOutputBufferPointer = 2,
  Dltbc1CoreErrorCode_NullEstimatorPointer = 3,
  Dltbc1CoreErrorCode_NullTransformSettingsPointer = 4,
  Dltbc1CoreErrorCode_InvalidDataLength = 5,
  Dltbc1CoreErrorCode_OutputBufferTooSmall = 6,
  Dltbc1CoreErrorCode_SizeEstimationError = 7,
  Dltbc1CoreErrorCode_TransformationError = 8, /* also: any CUDA failure */
} Dltbc1CoreErrorCode;

/* transform_auto.rs:61-66 */
typedef struct Dltbc1CoreResult {
  Dltbc1CoreErrorCode error_code;
} Dltbc1CoreResult;

/* transform_auto.rs:27-34, transform_with_settings.rs:14-21: bool first, then the INTERNAL variant. */
typedef struct Dltbc1CoreTransformSettings {
  bool split_colour_endpoints;
  DltCoreYCoCgVariant decorrelation_mode;
} Dltbc1CoreTransformSettings;
typedef Dltbc1CoreTransformSettings Dltbc1CoreUntransformSettings;

/* transform_auto.rs:15-24 */
typedef struct Dltbc1CoreAutoTransformSettings {
  bool use_all_modes;
} Dltbc1CoreAutoTransformSettings;

/* transform_with_settings.rs:73 — input null (1), output null (2), len % 8 (5), output_len < input_len (6). */
Dltbc1CoreResult dltbc1core_transform(const uint8_t *input, size_t input_len, uint8_t *output,
                                      size_t output_len, Dltbc1CoreTransformSettings details);
/* transform_with_settings.rs:119 */
Dltbc1CoreResult dltbc1core_untransform(const uint8_t *input, size_t input_len, uint8_t *output,
                                        size_t output_len, Dltbc1CoreUntransformSettings details);
/* transform_auto.rs:143-190 — data (1), output (2), estimator (3), out_details (4) null checks. */
Dltbc1CoreResult dltbc1core_transform_auto(const uint8_t *data, size_t data_len, uint8_t *output,
                                           size_t output_len, const DltSizeEstimator *estimator,
                                           Dltbc1CoreAutoTransformSettings settings,
                                           Dltbc1CoreTransformSettings *out_details);

#ifdef __cplusplus
}
#endif
#endif
