/*
 * dxt_lossless_transform_bc2_core.h — the "unstable" core C ABI for BC2 (drop-in for the cbindgen
 * header of crate dxt-lossless-transform-bc2, feature c-exports).  file:line relative to
 * /root/reference/src/core/dxt-lossless-transform-bc2/src/c_api.
 *
 * NOTE the reference reuses the type names Dltbc2Result / Dltbc2ErrorCode / Dltbc2TransformSettings
 * for DIFFERENT layouts and values in this crate and in the -api crate.  Here the core types carry a
 * "Core" infix so both headers can be included; the layouts are the core crate's.
 */
#ifndef DXT_LOSSLESS_TRANSFORM_BC2_CORE_H
#define DXT_LOSSLESS_TRANSFORM_BC2_CORE_H

#include "dxt_lossless_transform_api_common.h"

#ifdef __cplusplus
extern "C" {
#endif

/* transform_auto.rs:37-58 */
typedef enum Dltbc2CoreErrorCode {
  Dltbc2CoreErrorCode_Success = 0,
  Dltbc2CoreErrorCode_NullDataPointer = 1,
  Dltbc2CoreErrorCode_NullOutputBufferPointer = 2,
  Dltbc2CoreErrorCode_NullEstimatorPointer = 3,
  Dltbc2CoreErrorCode_NullTransformSettingsPointer = 4,
  Dltbc2CoreErrorCode_InvalidDataLength = 5,
  Dltbc2CoreErrorCode_OutputBufferTooSmall = 6,
  Dltbc2CoreErrorCode_SizeEstimationError = 7,
  Dltbc2CoreErrorCode_TransformationError = 8, /* also: any CUDA failure */
} Dltbc2CoreErrorCode;

/* transform_auto.rs:61-66 */
typedef struct Dltbc2CoreResult {
  Dltbc2CoreErrorCode error_code;
} Dltbc2CoreResult;

/* transform_auto.rs:27-34, transform_with_settings.rs:14-21: bool first, then the INTERNAL variant. */
typedef struct Dltbc2CoreTransformSettings {
  bool split_colour_endpoints;
  DltCoreYCoCgVariant decorrelation_mode;
} Dltbc2CoreTransformSettings;
typedef Dltbc2CoreTransformSettings Dltbc2CoreUntransformSettings;

/* transform_auto.rs:15-24 */
typedef struct Dltbc2CoreAutoTransformSettings {
  bool use_all_modes;
} Dltbc2CoreAutoTransformSettings;

/* transform_with_settings.rs:73 — input null (1), output null (2), len % 16 (5), output_len < input_len (6). */
Dltbc2CoreResult dltbc2core_transform(const uint8_t *input, size_t input_len, uint8_t *output,
                                      size_t output_len, Dltbc2CoreTransformSettings details);
/* transform_with_settings.rs:119 */
Dltbc2CoreResult dltbc2core_untransform(const uint8_t *input, size_t input_len, uint8_t *output,
                                        size_t output_len, Dltbc2CoreUntransformSettings details);
/* transform_auto.rs:143-190 — data (1), output (2), estimator (3), out_details (4) null checks. */
Dltbc2CoreResult dltbc2core_transform_auto(const uint8_t *data, size_t data_len, uint8_t *output,
                                           size_t output_len, const DltSizeEstimator *estimator,
                                           Dltbc2CoreAutoTransformSettings settings,
                                           Dltbc2CoreTransformSettings *out_details);

#ifdef __cplusplus
}
#endif
#endif
