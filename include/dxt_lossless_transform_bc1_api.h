/*
 * dxt_lossless_transform_bc1_api.h — stable C ABI for BC1 (drop-in for the reference's cbindgen
 * header of crate dxt-lossless-transform-bc1-api, feature c-exports).
 *
 * Every function below replaces the reference function of the same name; file:line are relative to
 * /root/reference/src/api/dxt-lossless-transform-bc1-api/src/c_api.  Same names, argument order,
 * struct layouts, error-code values and null-check order.  The work runs on the current CUDA
 * device (or the one chosen with dltcuda_set_device); there is no CPU fallback, a CUDA failure is
 * reported as AllocationFailed (the stable enum has no generic failure code).
 */
#ifndef DXT_LOSSLESS_TRANSFORM_BC1_API_H
#define DXT_LOSSLESS_TRANSFORM_BC1_API_H

#include "dxt_lossless_transform_api_common.h"

#ifdef __cplusplus
extern "C" {
#endif

/* error.rs:12-39 (repr(C) enum). */
typedef enum Dltbc1ErrorCode {
  Dltbc1ErrorCode_Success = 0,
  Dltbc1ErrorCode_InvalidLength = 1,
  Dltbc1ErrorCode_OutputBufferTooSmall = 2,
  Dltbc1ErrorCode_AllocationFailed = 3,
  Dltbc1ErrorCode_SizeEstimationFailed = 4,
  Dltbc1ErrorCode_NullDataPointer = 5,
  Dltbc1ErrorCode_NullEstimatorPointer = 6,
  Dltbc1ErrorCode_NullTransformSettingsPointer = 7,
  Dltbc1ErrorCode_NullInputPointer = 8,
  Dltbc1ErrorCode_NullOutputBufferPointer = 9,
  Dltbc1ErrorCode_NullManualTransformBuilderPointer = 10,
  Dltbc1ErrorCode_NullBuilderPointer = 11,
  Dltbc1ErrorCode_NullManualBuilderOutputPointer = 12,
} Dltbc1ErrorCode;

/* error.rs:42-46 */
typedef struct Dltbc1Result {
  Dltbc1ErrorCode error_code;
} Dltbc1Result;

/* mod.rs:190-208 (ABI-stable settings structs; defaults Variant1 / true, mod.rs:210-226). */
typedef struct Dltbc1TransformSettings {
  YCoCgVariant decorrelation_mode;
  bool split_colour_endpoints;
} Dltbc1TransformSettings;
typedef Dltbc1TransformSettings Dltbc1UntransformSettings;

/* Opaque builders (transform/manual_transform_builder.rs:25-44, auto_transform_builder.rs:35-38). */
typedef struct Dltbc1ManualTransformBuilder Dltbc1ManualTransformBuilder;
typedef struct Dltbc1AutoTransformBuilder Dltbc1AutoTransformBuilder;

/* transform/manual_transform_builder.rs:71 — new builder with default settings (Variant1, split). */
Dltbc1ManualTransformBuilder *dltbc1_new_ManualTransformBuilder(void);
/* :86 — null-safe. */
void dltbc1_free_ManualTransformBuilder(Dltbc1ManualTransformBuilder *builder);
/* :107 — null in, null out. */
Dltbc1ManualTransformBuilder *dltbc1_clone_ManualTransformBuilder(
    const Dltbc1ManualTransformBuilder *builder);
/* :150 — null builder is a no-op. */
void dltbc1_ManualTransformBuilder_SetDecorrelationMode(Dltbc1ManualTransformBuilder *builder,
                                                        YCoCgVariant mode);
/* :183 */
void dltbc1_ManualTransformBuilder_SetSplitColourEndpoints(Dltbc1ManualTransformBuilder *builder,
                                                           bool split);
/* :203 */
void dltbc1_ManualTransformBuilder_ResetToDefaults(Dltbc1ManualTransformBuilder *builder);
/* :256 — checks in order: input (5), output (9), builder (10), len % 8 (1), output_len < input_len (2).
 * Host pointers; blocks until `output` holds the first input_len transformed bytes. */
Dltbc1Result dltbc1_ManualTransformBuilder_Transform(const uint8_t *input, size_t input_len,
                                                     uint8_t *output, size_t output_len,
                                                     Dltbc1ManualTransformBuilder *builder);
/* :323 — exact inverse with the same builder settings. */
Dltbc1Result dltbc1_ManualTransformBuilder_Untransform(const uint8_t *input, size_t input_len,
                                                       uint8_t *output, size_t output_len,
                                                       Dltbc1ManualTransformBuilder *builder);

/* transform/auto_transform_builder.rs:63 — copies *estimator; null in, null out. */
Dltbc1AutoTransformBuilder *dltbc1_new_AutoTransformBuilder(const DltSizeEstimator *estimator);
/* :88 */
void dltbc1_free_AutoTransformBuilder(Dltbc1AutoTransformBuilder *builder);
/* :121 — null builder -> NullBuilderPointer (11). */
Dltbc1Result dltbc1_AutoTransformBuilder_SetUseAllDecorrelationModes(
    Dltbc1AutoTransformBuilder *builder, bool use_all);
/* :190-245 — checks builder (11), data (5), output (9), out_manual_builder (12).  Tries every
 * candidate in the reference's test order, keeps the strictly smallest estimate, leaves the winner's
 * transform in `output`, returns a new manual builder holding the winning settings (caller frees);
 * on failure *out_manual_builder = NULL.  With the estimator from dltltu_new_size_estimator() the
 * whole search runs on the GPU; any other estimator is called back with host memory per candidate. */
Dltbc1Result dltbc1_AutoTransformBuilder_Transform(Dltbc1AutoTransformBuilder *builder,
                                                   const uint8_t *data, size_t data_len,
                                                   uint8_t *output, size_t output_len,
                                                   Dltbc1ManualTransformBuilder **out_manual_builder);

/* error.rs:131 — static NUL-terminated message for a code. */
const char *dltbc1_error_message(Dltbc1ErrorCode error_code);

#ifdef __cplusplus
}
#endif
#endif
