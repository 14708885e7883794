/*
 * dxt_lossless_transform_bc3_core.h — ADDITIVE EXTENSION: a C ABI for BC3.
 *
 * The reference has NO C ABI for BC3 (core/dxt-lossless-transform-bc3 has no c_api/, the bc3-api crate
 * is empty); its BC3 boundary is the Rust functions
 *   transform_bc3_with_settings    core/dxt-lossless-transform-bc3/src/transform/transform_with_settings.rs:32
 *   untransform_bc3_with_settings  .../transform_with_settings.rs:162
 *   transform_bc3_auto             .../transform_auto.rs:196
 * (safe wrappers .../safe/transform_with_settings.rs:90,196).  These three symbols give those
 * functions a C ABI in the style of the BC1/BC2 core crates: same error codes and check order.
 */
#ifndef DXT_LOSSLESS_TRANSFORM_BC3_CORE_H
#define DXT_LOSSLESS_TRANSFORM_BC3_CORE_H

#include "dxt_lossless_transform_api_common.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum Dltbc3CoreErrorCode {
  Dltbc3CoreErrorCode_Success = 0,
  Dltbc3CoreErrorCode_NullDataPointer = 1,
  Dltbc3CoreErrorCode_NullOutputBufferPointer = 2,
  Dltbc3CoreErrorCode_NullEstimatorPointer = 3,
  Dltbc3CoreErrorCode_NullTransformSettingsPointer = 4,
  Dltbc3CoreErrorCode_InvalidDataLength = 5,
  Dltbc3CoreErrorCode_OutputBufferTooSmall = 6,
  Dltbc3CoreErrorCode_SizeEstimationError = 7,
  Dltbc3CoreErrorCode_TransformationError = 8,
} Dltbc3CoreErrorCode;

typedef struct Dltbc3CoreResult {
  Dltbc3CoreErrorCode error_code;
} Dltbc3CoreResult;

/* Bc3TransformSettings (settings.rs:16-31): default (Variant1, true, true). */
typedef struct Dltbc3CoreTransformSettings {
  bool split_alpha_endpoints;
  bool split_colour_endpoints;
  DltCoreYCoCgVariant decorrelation_mode;
} Dltbc3CoreTransformSettings;
typedef Dltbc3CoreTransformSettings Dltbc3CoreUntransformSettings;

typedef struct Dltbc3CoreAutoTransformSettings {
  bool use_all_modes;
} Dltbc3CoreAutoTransformSettings;

Dltbc3CoreResult dltbc3core_transform(const uint8_t *input, size_t input_len, uint8_t *output,
                                      size_t output_len, Dltbc3CoreTransformSettings details);
Dltbc3CoreResult dltbc3core_untransform(const uint8_t *input, size_t input_len, uint8_t *output,
                                        size_t output_len, Dltbc3CoreUntransformSettings details);
/* 8 candidates (fast) or 16 (all modes); estimate = alpha endpoints [0,2N) + colour endpoints
 * [len/2, len/2+4N) (transform_auto.rs:253-281). */
Dltbc3CoreResult dltbc3core_transform_auto(const uint8_t *data, size_t data_len, uint8_t *output,
                                           size_t output_len, const DltSizeEstimator *estimator,
                                           Dltbc3CoreAutoTransformSettings settings,
                                           Dltbc3CoreTransformSettings *out_details);

#ifdef __cplusplus
}
#endif
#endif
