// cabi_internal.h — types shared by the translation units that implement the C ABI
// (cabi.cu: dltbc*/dltltu/dltcuda symbols; file_formats.cu: dltff_*/dltdds_*/is_dds/parse_dds).
#pragma once
#include <cstddef>
#include <cstdint>

#include "bcn_layout.h"

// =================================================================================================
// Shared types
// =================================================================================================
extern "C" {

// api-common/src/c_api/size_estimation.rs:17-52
typedef uint32_t (*DltMaxCompressedSizeFn)(void* context, size_t len_bytes, size_t* out_size);
typedef uint32_t (*DltEstimateCompressedSizeFn)(void* context, const uint8_t* input_ptr, size_t len_bytes,
                                                uint8_t* output_ptr, size_t output_len, size_t* out_size);
struct DltSizeEstimator {
    void* context;
    DltMaxCompressedSizeFn max_compressed_size;
    DltEstimateCompressedSizeFn estimate_compressed_size;
};

struct DltResult {  // Dltbc{1,2}Result in both crates: one repr(C) enum field
    int32_t error_code;
};

// core crates: { bool split_colour_endpoints; YCoCgVariant(u8, internal numbering) }
struct DltCoreSettings {
    bool split_colour_endpoints;
    uint8_t decorrelation_mode;
};
struct DltCoreAutoSettings {
    bool use_all_modes;
};
// additive BC3 settings (core style)
struct DltCoreBc3Settings {
    bool split_alpha_endpoints;
    bool split_colour_endpoints;
    uint8_t decorrelation_mode;
};
// additive device API
struct DltcudaPayload;
struct DltcudaSettings {
    uint8_t format;              // 1, 2, 3
    uint8_t decorrelation_mode;  // internal numbering: None=0, Variant1=1, Variant2=2, Variant3=3
    bool split_alpha_endpoints;  // BC3 only
    bool split_colour_endpoints;
};

struct DltcudaPayload {  // one independent host payload of a batch
    const uint8_t* input;
    uint8_t* output;
    size_t len;
    DltcudaSettings settings;
};

}  // extern "C"


extern "C" {
struct DltcudaAutoJob {  // one payload of dltcuda_transform_auto_batch
    uint8_t format;
    const uint8_t* input;
    uint8_t* output;
    size_t len;
    DltcudaSettings out_settings;  // out: the winning settings
    int32_t status;                // out: DltcudaStatus of this job
};
}

namespace dlt {
namespace cabi {

// Stable API codes — api/dxt-lossless-transform-bc1-api/src/c_api/error.rs:12-39
enum ApiCode : int32_t {
    kApiSuccess = 0,
    kApiInvalidLength = 1,
    kApiOutputBufferTooSmall = 2,
    kApiAllocationFailed = 3,
    kApiSizeEstimationFailed = 4,
    kApiNullDataPointer = 5,
    kApiNullEstimatorPointer = 6,
    kApiNullTransformSettingsPointer = 7,
    kApiNullInputPointer = 8,
    kApiNullOutputBufferPointer = 9,
    kApiNullManualTransformBuilderPointer = 10,
    kApiNullBuilderPointer = 11,
    kApiNullManualBuilderOutputPointer = 12,
};


// What a builder holds: Bc{1,2}ManualTransformBuilder { settings } (manual_transform_builder.rs).
struct ManualBuilder {
    int format;
    int variant;  // internal numbering
    bool split_colour;
};
struct AutoBuilder {
    int format;
    DltSizeEstimator estimator;  // a COPY, as in auto_transform_builder.rs:35-38
    bool use_all;
};


// Stable YCoCgVariant numbering (api-common/src/reexports/color_565.rs:65-85):
// Variant1=0, Variant2=1, Variant3=2, None=3  <->  internal None=0, Variant1..3=1..3.
inline int stable_to_internal(uint8_t v) { return v == 3 ? kNone : v + 1; }
inline uint8_t internal_to_stable(int v) { return v == kNone ? 3 : (uint8_t)(v - 1); }

// Bodies of dltbcN_ManualTransformBuilder_{Transform,Untransform} and dltbcN_AutoTransformBuilder_Transform
// (cabi.cu); the bundle dispatch of file_formats.cu goes through the same code.
DltResult api_manual_run(int format, bool inverse, const uint8_t* input, size_t input_len, uint8_t* output,
                         size_t output_len, ManualBuilder* b);
// Like the exported call, but hands the chosen settings back by value instead of allocating a builder.
DltResult api_auto_settings(int format, const AutoBuilder* b, const uint8_t* data, size_t data_len, uint8_t* output,
                            size_t output_len, Settings* best);

// True when `e` is the estimator handed out by dltltu_new_size_estimator (the search then runs entirely on the GPU).
bool is_gpu_ltu_estimator(const DltSizeEstimator& e);
// Body of dltcuda_transform_auto_batch: returns a DltcudaStatus (0 = Ok).
int auto_batch_host(DltcudaAutoJob* jobs, size_t count, bool use_all);

}  // namespace cabi
}  // namespace dlt
