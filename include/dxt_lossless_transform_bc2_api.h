/*
 * dxt_lossless_transform_bc2_api.h — stable C ABI for BC2 (drop-in for the reference's cbindgen
 * header of crate dxt-lossless-transform-bc2-api, feature c-exports).
 *
 * Every function below replaces the reference function of the same name; file:line are relative to
 * /root/reference/src/api/dxt-lossless-transform-bc2-api/src/c_api.  Same names, argument order,
 * struct layouts, error-code values and null-check order.  The work runs on the current CUDA
 * device (or the one chosen with dltcuda_set_device); there is no CPU fallback, a CUDA failure is
 * reported as AllocationFailed (the stable enum has no generic failure code).
 */
#ifndef DXT_LOSSLESS_TRANSFORM_BC2_API_H
#define DXT_LOSSLESS_TRANSFORM_BC2_API_H

#include "dxt_lossless_transform_api_common.h"

#ifdef __cplusplus
extern "C" {
#endif

/* error.rs:12-39 (repr(C) enum). */
typedef enum Dltbc2ErrorCode {
  Dltbc2ErrorCode_Success = 0,
  Dltbc2ErrorCode_InvalidLength = 1,
  Dltbc2ErrorCode_OutputBufferTooSmall = 2,
  Dltbc2ErrorCode_AllocationFailed = 3,
  Dltbc2ErrorCode_SizeEstimationFailed = 4,
  Dltbc2ErrorCode_NullDataPointer = 5,
  Dltbc2ErrorCode_NullEstimatorPointer = 6,
  Dltbc2ErrorCode_NullTransformSettingsPointer = 7,
  Dltbc2ErrorCode_NullInputPointer = 8,
  Dltbc2ErrorCode_NullOutputBufferPointer = 9,
  Dltbc2ErrorCode_NullManualTransformBuilderPointer = 10,
  Dltbc2ErrorCode_NullBuilderPointer = 11,
  Dltbc2ErrorCode_NullManualBuilderOutputPointer = 12,
} Dltbc2ErrorCode;

/* error.rs:42-46 */
typedef struct Dltbc2Result {
  Dltbc2ErrorCode error_code;
} Dltbc2Result;

/* mod.rs:190-208 (ABI-stable settings structs; defaults Variant1 / true, mod.rs:210-226). */
typedef struct Dltbc2TransformSettings {
  YCoCgVariant decorrelation_mode;
  bool split_colour_endpoints;
} Dltbc2TransformSettings;
typedef Dltbc2TransformSettings Dltbc2UntransformSettings;

/* Opaque builders (transform/manual_transform_builder.rs:25-44, auto_transform_builder.rs:35-38). */
typedef struct Dltbc2ManualTransformBuilder Dltbc2ManualTransformBuilder;
typedef struct Dltbc2AutoTransformBuilder Dltbc2AutoTransformBuilder;

/* transform/manual_transform_builder.rs:71 — new builder with default settings (Variant1, split). */
Dltbc2ManualTransformBuilder *dltbc2_new_ManualTransformBuilder(void);
/* :86 — null-safe. */
void dltbc2_free_ManualTransformBuilder(Dltbc2ManualTransformBuilder *builder);
/* :107 — null in, null out. */
Dltbc2ManualTransformBuilder *dltbc2_clone_ManualTransformBuilder(
    const Dltbc2ManualTransformBuilder *builder);
/* :150 — null builder is a no-op. */
void dltbc2_ManualTransformBuilder_SetDecorrelationMode(Dltbc2ManualTransformBuilder *builder,
                                                        YCoCgVariant mode);
/* :183 */
void dltbc2_ManualTransformBuilder_SetSplitColourEndpoints(Dltbc2ManualTransformBuilder *builder,
                                                           bool split);
/* :203 */
void dltbc2_ManualTransformBuilder_ResetToDefaults(Dltbc2ManualTransformBuilder *builder);
/* :256 — checks in order: input (5), output (9), builder (10), len % 16 (1), output_len < input_len (2).
 * Host pointers; blocks until `output` holds the first input_len transformed bytes. */
Dltbc2Result dltbc2_ManualTransformBuilder_Transform(const uint8_t *input, size_t input_len,
                                                     uint8_t *output, size_t output_len,
                                                     Dltbc2ManualTransformBuilder *builder);
/* :323 — exact inverse with the same builder settings. */
Dltbc2Result dltbc2_ManualTransformBuilder_Untransform(const uint8_t *input, size_t input_len,
                                                       uint8_t *output, size_t output_len,
                                                       Dltbc2ManualTransformBuilder *builder);

/* transform/auto_transform_builder.rs:63 — copies *estimator; null in, null out. */
Dltbc2AutoTransformBuilder *dltbc2_new_AutoTransformBuilder(const DltSizeEstimator *estimator);
/* :88 */
void dltbc2_free_AutoTransformBuilder(Dltbc2AutoTransformBuilder *builder);
/* :121 — null builder -> NullBuilderPointer (11). */
Dltbc2Result dltbc2_AutoTransformBuilder_SetUseAllDecorrelationModes(
    Dltbc2AutoTransformBuilder *builder, bool use_all);
/* :190-245 — checks builder (11), data (5), output (9), out_manual_builder (12).  Tries every
 * candidate in the reference's test order, keeps the strictly smallest estimate, leaves the winner's
 * transform in `output`, returns a new manual builder holding the winning settings (caller frees);
 * on failure *out_manual_builder = NULL.  With the estimator from dltltu_new_size_estimator() the
 * whole search runs on the GPU; any other estimator is called back with host memory per candidate. */
Dltbc2Result dltbc2_AutoTransformBuilder_Transform(Dltbc2AutoTransformBuilder *builder,
                                                   const uint8_t *data, size_t data_len,
                                                   uint8_t *output, size_t output_len,
                                                   Dltbc2ManualTransformBuilder **out_manual_builder);

/* error.rs:131 — static NUL-terminated message for a code. */
const char *dltbc2_error_message(Dltbc2ErrorCode error_code);

#ifdef __cplusplus
}
#endif
#endif
