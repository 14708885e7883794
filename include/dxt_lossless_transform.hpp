// dxt_lossless_transform.hpp — header-only C++ host side over the C ABI.
//
// The reference's host language is Rust, which this image cannot compile; this is the same surface in
// C++ (the "compiled host code above the C ABI"), mirroring names, defaults and error behaviour of
//   dxt_lossless_transform_bc{1,2,3}::{transform,untransform}_bcN_with_settings[_safe]
//       core/dxt-lossless-transform-bc1/src/transform/safe/transform_with_settings.rs:88,192 (bc2, bc3 alike)
//   dxt_lossless_transform_bc{1,2,3}::transform_bcN_auto_safe          .../safe/transform_auto.rs:95
//   Bc{1,2}ManualTransformBuilder / Bc{1,2}AutoTransformBuilder
//       api/dxt-lossless-transform-bc1-api/src/transform/{manual,auto}_transform_builder.rs
//   LosslessTransformUtilsSizeEstimation      extensions/estimators/dxt-lossless-transform-ltu/src/lib.rs:49
//   ZStandardSizeEstimation                   extensions/compressors/dxt-lossless-transform-zstd/src/lib.rs:54
//   experimental::normalize_blocks            core/dxt-lossless-transform-bc1/src/experimental/normalize_blocks/
// Link with -ldxt_lossless_transform_cuda.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>

#include "dxt_lossless_transform_bc1_api.h"
#include "dxt_lossless_transform_bc1_core.h"
#include "dxt_lossless_transform_bc2_api.h"
#include "dxt_lossless_transform_bc2_core.h"
#include "dxt_lossless_transform_bc3_core.h"
#include "dxt_lossless_transform_cuda.h"
#include "dxt_lossless_transform_ltu.h"
#include "dxt_lossless_transform_zstd.h"

namespace dxt_lossless_transform {

// Internal numbering, as in dxt_lossless_transform_common::color_565::YCoCgVariant.
enum class YCoCgVariant : uint8_t { None = 0, Variant1 = 1, Variant2 = 2, Variant3 = 3 };

struct Bc1TransformSettings {
    YCoCgVariant decorrelation_mode = YCoCgVariant::Variant1;  // Default (settings.rs:35-43)
    bool split_colour_endpoints = true;
};
using Bc2TransformSettings = Bc1TransformSettings;
struct Bc3TransformSettings {
    YCoCgVariant decorrelation_mode = YCoCgVariant::Variant1;  // Default (bc3 settings.rs:39-48)
    bool split_alpha_endpoints = true;
    bool split_colour_endpoints = true;
};

// Bc{1,2,3}ValidationError / Bc{1,2,3}AutoTransformError as exceptions.
struct InvalidLength : std::invalid_argument {
    explicit InvalidLength(size_t len) : std::invalid_argument("Invalid input length: " + std::to_string(len)), length(len) {}
    size_t length;
};
struct OutputBufferTooSmall : std::invalid_argument {
    OutputBufferTooSmall(size_t needed_, size_t actual_)
        : std::invalid_argument("Output buffer too small: needed " + std::to_string(needed_) + ", got " + std::to_string(actual_)),
          needed(needed_), actual(actual_) {}
    size_t needed, actual;
};
struct SizeEstimationError : std::runtime_error {
    SizeEstimationError() : std::runtime_error("Size estimation failed") {}
};
struct DeviceError : std::runtime_error {  // no CPU fallback: a CUDA failure is an error
    explicit DeviceError(int code) : std::runtime_error(std::string("CUDA failure: ") + dltcuda_last_error()), core_code(code) {}
    int core_code;
};

namespace detail {
inline void check_core(int code, size_t in_len, size_t out_len) {
    switch (code) {
        case 0: return;
        case 5: throw InvalidLength(in_len);
        case 6: throw OutputBufferTooSmall(in_len, out_len);
        case 7: throw SizeEstimationError();
        default: throw DeviceError(code);
    }
}
inline Dltbc1CoreTransformSettings core1(const Bc1TransformSettings& s) {
    return {s.split_colour_endpoints, static_cast<DltCoreYCoCgVariant>(s.decorrelation_mode)};
}
inline Dltbc2CoreTransformSettings core2(const Bc2TransformSettings& s) {
    return {s.split_colour_endpoints, static_cast<DltCoreYCoCgVariant>(s.decorrelation_mode)};
}
inline Dltbc3CoreTransformSettings core3(const Bc3TransformSettings& s) {
    return {s.split_alpha_endpoints, s.split_colour_endpoints, static_cast<DltCoreYCoCgVariant>(s.decorrelation_mode)};
}
}  // namespace detail

// ---- with-settings (host buffers; synchronous; only the first in_len bytes of `out` are written) ----
inline void transform_bc1_with_settings(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, Bc1TransformSettings s = {}) {
    detail::check_core(dltbc1core_transform(in, in_len, out, out_len, detail::core1(s)).error_code, in_len, out_len);
}
inline void untransform_bc1_with_settings(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, Bc1TransformSettings s = {}) {
    detail::check_core(dltbc1core_untransform(in, in_len, out, out_len, detail::core1(s)).error_code, in_len, out_len);
}
inline void transform_bc2_with_settings(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, Bc2TransformSettings s = {}) {
    detail::check_core(dltbc2core_transform(in, in_len, out, out_len, detail::core2(s)).error_code, in_len, out_len);
}
inline void untransform_bc2_with_settings(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, Bc2TransformSettings s = {}) {
    detail::check_core(dltbc2core_untransform(in, in_len, out, out_len, detail::core2(s)).error_code, in_len, out_len);
}
inline void transform_bc3_with_settings(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, Bc3TransformSettings s = {}) {
    detail::check_core(dltbc3core_transform(in, in_len, out, out_len, detail::core3(s)).error_code, in_len, out_len);
}
inline void untransform_bc3_with_settings(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, Bc3TransformSettings s = {}) {
    detail::check_core(dltbc3core_untransform(in, in_len, out, out_len, detail::core3(s)).error_code, in_len, out_len);
}

// ---- estimators -----------------------------------------------------------------------------------
// LosslessTransformUtilsSizeEstimation: its callbacks are the library's own, so transform_bcN_auto keeps
// the whole search on the GPU.
class LosslessTransformUtilsSizeEstimation {
public:
    LosslessTransformUtilsSizeEstimation() : e_(dltltu_new_size_estimator()) {
        if (!e_) throw std::bad_alloc();
    }
    ~LosslessTransformUtilsSizeEstimation() { dltltu_free_size_estimator(e_); }
    LosslessTransformUtilsSizeEstimation(const LosslessTransformUtilsSizeEstimation&) = delete;
    LosslessTransformUtilsSizeEstimation& operator=(const LosslessTransformUtilsSizeEstimation&) = delete;
    const DltSizeEstimator* c_estimator() const { return e_; }
    size_t estimate_compressed_size(const uint8_t* data, size_t len) const {
        size_t out = 0;
        if (e_->estimate_compressed_size(e_->context, data, len, nullptr, 0, &out) != 0) throw SizeEstimationError();
        return out;
    }

private:
    DltSizeEstimator* e_;
};

// ZStandardSizeEstimation (dxt-lossless-transform-zstd/src/lib.rs:54-140): real zstd sizes with the reference's
// parameters; inside transform_bcN_auto the candidates are transformed on the GPU and compressed concurrently.
struct InvalidLevel : std::invalid_argument {
    explicit InvalidLevel(int level) : std::invalid_argument("Invalid compression level: " + std::to_string(level)) {}
};
class ZStandardSizeEstimation {
public:
    explicit ZStandardSizeEstimation(int compression_level) : e_(nullptr) {
        if (compression_level < 1 || compression_level > 22) throw InvalidLevel(compression_level);   // lib.rs:62-66
        e_ = dltzstd_new_size_estimator(compression_level);
        if (!e_) throw std::runtime_error("no libzstd could be loaded (set DLTCUDA_LIBZSTD)");
    }
    static ZStandardSizeEstimation new_fast() { return ZStandardSizeEstimation(1); }
    static ZStandardSizeEstimation new_default() { return ZStandardSizeEstimation(3); }
    static ZStandardSizeEstimation new_best() { return ZStandardSizeEstimation(22); }
    ~ZStandardSizeEstimation() { dltzstd_free_size_estimator(e_); }
    ZStandardSizeEstimation(ZStandardSizeEstimation&& o) noexcept : e_(std::exchange(o.e_, nullptr)) {}
    ZStandardSizeEstimation(const ZStandardSizeEstimation&) = delete;
    ZStandardSizeEstimation& operator=(const ZStandardSizeEstimation&) = delete;
    const DltSizeEstimator* c_estimator() const { return e_; }
    size_t max_compressed_size(size_t len) const {
        size_t out = 0;
        if (e_->max_compressed_size(e_->context, len, &out) != 0) throw SizeEstimationError();
        return out;
    }
    static unsigned library_version() { return dltzstd_version_number(); }

private:
    DltSizeEstimator* e_;
};

// ---- transform_bcN_auto ------------------------------------------------------------------------------
inline Bc1TransformSettings transform_bc1_auto(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len,
                                               const DltSizeEstimator* estimator, bool use_all_decorrelation_modes = false) {
    Dltbc1CoreTransformSettings d{};
    detail::check_core(dltbc1core_transform_auto(in, in_len, out, out_len, estimator, {use_all_decorrelation_modes}, &d).error_code,
                       in_len, out_len);
    return {static_cast<YCoCgVariant>(d.decorrelation_mode), d.split_colour_endpoints};
}
inline Bc2TransformSettings transform_bc2_auto(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len,
                                               const DltSizeEstimator* estimator, bool use_all_decorrelation_modes = false) {
    Dltbc2CoreTransformSettings d{};
    detail::check_core(dltbc2core_transform_auto(in, in_len, out, out_len, estimator, {use_all_decorrelation_modes}, &d).error_code,
                       in_len, out_len);
    return {static_cast<YCoCgVariant>(d.decorrelation_mode), d.split_colour_endpoints};
}
inline Bc3TransformSettings transform_bc3_auto(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len,
                                               const DltSizeEstimator* estimator, bool use_all_decorrelation_modes = false) {
    Dltbc3CoreTransformSettings d{};
    detail::check_core(dltbc3core_transform_auto(in, in_len, out, out_len, estimator, {use_all_decorrelation_modes}, &d).error_code,
                       in_len, out_len);
    return {static_cast<YCoCgVariant>(d.decorrelation_mode), d.split_alpha_endpoints, d.split_colour_endpoints};
}

// ---- experimental::normalize_blocks (BC1) ----------------------------------------------------------------
namespace experimental {
enum class ColorNormalizationMode : int { None = 0, Color0Only = 1, ReplicateColor = 2 };   // normalize.rs:487-500
struct Bc1TransformDetailsWithNormalization {   // normalize_blocks/mod.rs:98-111
    ColorNormalizationMode color_normalization_mode = ColorNormalizationMode::None;
    YCoCgVariant decorrelation_mode = YCoCgVariant::Variant1;
    bool split_colour_endpoints = true;
};
inline void check_len(size_t in_len, size_t out_len) {
    if (in_len % 8) throw InvalidLength(in_len);
    if (out_len < in_len) throw OutputBufferTooSmall(in_len, out_len);
}
// normalize_blocks (normalize.rs:38); `out` may be `in`.
inline void normalize_blocks(const uint8_t* in, uint8_t* out, size_t len, ColorNormalizationMode mode) {
    check_len(len, len);
    if (const int rc = dltcuda_bc1_normalize_blocks(in, out, len, static_cast<int>(mode))) throw DeviceError(rc);
}
// transform_bc1_with_normalize_blocks (transform.rs:65): one fused pass on the GPU.
inline void transform_bc1_with_normalize_blocks(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len,
                                                Bc1TransformDetailsWithNormalization d = {}) {
    check_len(in_len, out_len);
    if (const int rc = dltcuda_bc1_transform_with_normalize_blocks(in, out, in_len, static_cast<int>(d.color_normalization_mode),
                                                                   static_cast<DltCoreYCoCgVariant>(d.decorrelation_mode),
                                                                   d.split_colour_endpoints))
        throw DeviceError(rc);
}
// transform_bc1_auto_with_normalization (transform.rs:222) with the LTU-semantics estimator on the GPU.
inline Bc1TransformDetailsWithNormalization transform_bc1_auto_with_normalization(const uint8_t* in, size_t in_len, uint8_t* out,
                                                                                  size_t out_len, bool use_all_decorrelation_modes = false) {
    check_len(in_len, out_len);
    int norm = 0;
    DltCoreYCoCgVariant var{};
    bool split = false;
    if (const int rc = dltcuda_bc1_transform_auto_with_normalization(in, out, in_len, use_all_decorrelation_modes, &norm, &var, &split, nullptr))
        throw DeviceError(rc);
    return {static_cast<ColorNormalizationMode>(norm), static_cast<YCoCgVariant>(var), split};
}
}  // namespace experimental

// ---- page-locked buffers (full host-link speed for the host-pointer entry points) -----------------
class PinnedBuffer {
public:
    explicit PinnedBuffer(size_t bytes) : p_(static_cast<uint8_t*>(dltcuda_alloc_pinned(bytes))), n_(bytes) {
        if (!p_) throw std::bad_alloc();
    }
    ~PinnedBuffer() { dltcuda_free_pinned(p_); }
    PinnedBuffer(const PinnedBuffer&) = delete;
    PinnedBuffer& operator=(const PinnedBuffer&) = delete;
    uint8_t* data() { return p_; }
    size_t size() const { return n_; }

private:
    uint8_t* p_;
    size_t n_;
};

}  // namespace dxt_lossless_transform
