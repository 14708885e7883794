// cabi.cu — the C ABI of libdxt_lossless_transform_cuda.so.
//
// Three groups of symbols (declared in include/*.h, each with the reference file:line it replaces):
//   1. the reference's cbindgen surface, same names / structs / enum values / check order:
//        dltbc1_* , dltbc2_*          (api/dxt-lossless-transform-bc{1,2}-api/src/c_api/**)
//        dltbc1core_*, dltbc2core_*   (core/dxt-lossless-transform-bc{1,2}/src/c_api/**)
//        dltltu_*                     (extensions/estimators/dxt-lossless-transform-ltu/src/c_api.rs)
//   2. dltbc3core_*: BC3 has no C ABI in the reference (only Rust functions); these mirror the core
//      style for BC3 and are an additive extension.
//   3. dltcuda_*: additive device-resident / sharding / pinned-memory entry points.
// Nothing in here falls back to the CPU: without a usable CUDA device the calls return an error.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "auto_search.h"
#include "cabi_internal.h"
#include "bcn_kernels.h"
#include "bcn_layout.h"
#include "estimator.h"
#include "host_pipeline.h"
#include "zstd_estimator.h"

using namespace dlt;
using namespace dlt::cabi;

#define DLT_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

// Core codes — core/dxt-lossless-transform-bc1/src/c_api/transform_auto.rs:37-58
enum CoreCode : int32_t {
    kCoreSuccess = 0,
    kCoreNullDataPointer = 1,
    kCoreNullOutputBufferPointer = 2,
    kCoreNullEstimatorPointer = 3,
    kCoreNullTransformSettingsPointer = 4,
    kCoreInvalidDataLength = 5,
    kCoreOutputBufferTooSmall = 6,
    kCoreSizeEstimationError = 7,
    kCoreTransformationError = 8,
};

enum class Outcome { kOk, kInvalidLength, kTooSmall, kDevice, kOutOfMemory, kEstimator, kHostAlloc };

Outcome validate(size_t in_len, size_t out_len, int format) {
    // safe/transform_with_settings.rs:88-105: length check first, then the output size.
    if (in_len % (size_t)block_bytes(format) != 0) return Outcome::kInvalidLength;
    if (out_len < in_len) return Outcome::kTooSmall;
    return Outcome::kOk;
}

Outcome from_status(Status s) {
    return s == Status::kOk ? Outcome::kOk : s == Status::kOutOfMemory ? Outcome::kOutOfMemory : Outcome::kDevice;
}

Outcome transform_host(const Settings& st, bool inverse, const uint8_t* in, size_t in_len, uint8_t* out,
                       size_t out_len) {
    Outcome v = validate(in_len, out_len, st.format);
    if (v != Outcome::kOk) return v;
    if (st.variant < kNone || st.variant > kVariant3) return Outcome::kDevice;
    return from_status(run_host(st, inverse, in, out, in_len, -1));
}

int32_t api_code(Outcome o) {
    switch (o) {
        case Outcome::kOk: return kApiSuccess;
        case Outcome::kInvalidLength: return kApiInvalidLength;
        case Outcome::kTooSmall: return kApiOutputBufferTooSmall;
        case Outcome::kEstimator: return kApiSizeEstimationFailed;
        // The stable enum has no generic failure code; a CUDA failure reports as AllocationFailed.
        default: return kApiAllocationFailed;
    }
}
int32_t core_code(Outcome o) {
    switch (o) {
        case Outcome::kOk: return kCoreSuccess;
        case Outcome::kInvalidLength: return kCoreInvalidDataLength;
        case Outcome::kTooSmall: return kCoreOutputBufferTooSmall;
        case Outcome::kEstimator: return kCoreSizeEstimationError;
        default: return kCoreTransformationError;
    }
}

// ---- the LTU estimator callbacks (host pointers in, GPU estimator underneath) -------------------
uint32_t ltu_max_compressed_size(void* context, size_t, size_t* out_size) {
    if (!context || !out_size) return 1;  // ltu/src/c_api.rs:113-115
    *out_size = 0;                        // lib.rs:90-94: no scratch buffer needed
    return 0;
}

uint32_t ltu_estimate_compressed_size(void* context, const uint8_t* input, size_t len, uint8_t*, size_t,
                                      size_t* out_size) {
    if (!context || !out_size) return 1;  // ltu/src/c_api.rs:137-139
    if (!input || len == 0) {             // lib.rs:103-109
        *out_size = 0;
        return 0;
    }
    Status st;
    Context* ctx = acquire_context(-1, &st);
    if (!ctx) return 3;
    uint32_t rc = 3;
    if (ensure_device_buffers(ctx, len) == Status::kOk &&
        cudaMemcpyAsync(ctx->d_in, input, len, cudaMemcpyHostToDevice, ctx->stream[0]) == cudaSuccess) {
        LtuSegment seg{ctx->d_in, len};
        uint64_t m = 0;
        if (ltu_matches_device(ctx, &seg, 1, &m, ctx->stream[0]) == Status::kOk) {
            *out_size = ltu_estimate_from_matches(len, m);
            rc = 0;
        }
    }
    release_context_synced(ctx);
    return rc;
}

bool is_gpu_ltu(const DltSizeEstimator& e) { return e.estimate_compressed_size == &ltu_estimate_compressed_size; }

// ---- transform_bcN_auto on host pointers -----------------------------------------------------------
// C++ exceptions must not cross the C ABI (std::vector / std::thread allocate): every exported body that can throw runs
// inside guarded(), which maps bad_alloc to the caller's out-of-memory code and anything else to its failure code.
template <class F, class R>
R guarded(F&& f, R oom, R other) noexcept {
    try {
        return f();
    } catch (const std::bad_alloc&) {
        return oom;
    } catch (...) {
        return other;
    }
}

Outcome auto_host_impl(int format, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len,
                       const DltSizeEstimator& est, bool use_all, Settings* best);
Outcome auto_host(int format, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len,
                  const DltSizeEstimator& est, bool use_all, Settings* best) {
    return guarded([&] { return auto_host_impl(format, in, in_len, out, out_len, est, use_all, best); }, Outcome::kHostAlloc,
                   Outcome::kDevice);
}

Outcome auto_host_impl(int format, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len,
                       const DltSizeEstimator& est, bool use_all, Settings* best) {
    Outcome v = validate(in_len, out_len, format);
    if (v != Outcome::kOk) return v;
    const size_t len = in_len;

    // transform_auto.rs:215-227: scratch for the estimator sized by max_compressed_size(len/2 | len/4).
    size_t max_comp = 0;
    if (est.max_compressed_size(est.context, format == 1 ? len / 2 : len / 4, &max_comp) != 0) return Outcome::kEstimator;
    uint8_t* comp = nullptr;
    if (max_comp) {
        if (posix_memalign(reinterpret_cast<void**>(&comp), 64, max_comp) != 0) return Outcome::kHostAlloc;
    }
    struct FreeComp {
        uint8_t* p;
        ~FreeComp() { std::free(p); }
    } free_comp{comp};

    Settings order[kMaxCandidates];
    const int k = candidate_order(format, use_all, order);
    EstimateRange ranges[2];
    const int nr = estimate_ranges(format, len, ranges);

    if (len == 0) {
        // The reference still walks the candidates (estimating empty slices); first minimum wins.
        Settings best_s = default_settings(format);
        size_t best_size = SIZE_MAX;
        for (int i = 0; i < k; i++) {
            size_t total = 0;
            for (int r = 0; r < nr; r++) {
                size_t sz = 0;
                if (est.estimate_compressed_size(est.context, out, 0, comp, max_comp, &sz) != 0) return Outcome::kEstimator;
                total += sz;
            }
            if (total < best_size) best_size = total, best_s = order[i];
        }
        *best = best_s;
        return Outcome::kOk;
    }

    Status st;
    Context* ctx = acquire_context(-1, &st);
    if (!ctx) return from_status(st);
    struct Releaser {
        Context* c;
        ~Releaser() { release_context_synced(c); }
    } releaser{ctx};
    if ((st = ensure_device_buffers(ctx, len)) != Status::kOk) return from_status(st);
    cudaStream_t s = ctx->stream[0];
    auto cuda_fail = [](cudaError_t e) {
        note_cuda_error(e);
        return e == cudaErrorMemoryAllocation ? Outcome::kOutOfMemory : Outcome::kDevice;
    };
    cudaError_t e = cudaMemcpyAsync(ctx->d_in, in, len, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return cuda_fail(e);
    const size_t n = len / block_bytes(format);

    Settings best_s = default_settings(format);
    if (is_gpu_ltu(est)) {
        // Whole search on the device; one D2H of the winner.
        if ((st = auto_ltu_device(ctx, format, ctx->d_in, ctx->d_out, len, use_all, &best_s, nullptr, s)) != Status::kOk)
            return from_status(st);
    } else if (is_zstd_estimator(est)) {
        // dltzstd_new_size_estimator (zstd_estimator.cu): the callback is thread-safe and known, so every candidate is
        // compressed by its own host thread as soon as its endpoint streams have arrived — the reference's loop
        // (transform_auto.rs:230-262) compresses them one after another.  Same estimates, same strict-'<' selection.
        // Only DISTINCT ranges are compressed (BC3: the alpha range depends on split_alpha alone, the colour range on
        // variant + split_colour: 2 + 4 / 8 zstd passes instead of 16 / 32); zstd is deterministic, the per-candidate
        // totals are the reference's.
        const int level = zstd_estimator_level(est);
        const DistinctPlan plan = plan_distinct(format, order, k, nr);
        const int nseg = (int)plan.segs.size();
        size_t largest = 0;
        for (int r = 0; r < nr; r++) largest = std::max(largest, ranges[r].len);
        const size_t per = (largest + 255) / 256 * 256 + (max_comp + 255) / 256 * 256;   // one range + its compression buffer
        constexpr size_t kHostBudget = (size_t)4 << 30;   // pinned ranges + compression buffers in flight
        // one host thread per range in flight: bounded by the pinned budget AND by the cores of the box (several caller
        // threads may be searching at the same time; zstd is CPU-bound, more threads than cores only thrash)
        const size_t cores = std::max(1u, std::thread::hardware_concurrency());
        const int wave = (int)std::min({(size_t)nseg, std::max<size_t>(1, kHostBudget / per), cores});
        if ((st = ensure_host_scratch(ctx, (size_t)wave * per)) != Status::kOk) return from_status(st);
        std::vector<size_t> seg_size((size_t)nseg, 0);
        std::vector<uint32_t> seg_rc((size_t)nseg, 0);
        std::vector<cudaEvent_t> arrived((size_t)nseg, nullptr);
        Outcome failed = Outcome::kOk;
        int transformed = -1;   // candidate whose image currently sits in ctx->d_out
        for (int q0 = 0; q0 < nseg && failed == Outcome::kOk; q0 += wave) {
            const int wb = std::min(wave, nseg - q0);
            struct Joiner {   // a thread that cannot be started must not leave running ones un-joined (std::terminate)
                std::vector<std::thread> t;
                ~Joiner() {
                    for (auto& w : t)
                        if (w.joinable()) w.join();
                }
            } joiner;
            std::vector<std::thread>& workers = joiner.t;
            workers.reserve(wb);
            for (int c = 0; c < wb; c++) {
                const int q = q0 + c;
                const DistinctPlan::Seg sg = plan.segs[q];   // segments are in candidate order: one transform per candidate
                uint8_t* buf = ctx->h_scratch + (size_t)c * per;
                e = cudaSuccess;
                if (transformed != sg.cand) {
                    e = launch_transform(order[sg.cand], ctx->d_in, reference_layout(ctx->d_out, n, 0, order[sg.cand]), n, s);
                    transformed = sg.cand;
                }
                const size_t rlen = ranges[sg.range].len;
                if (e == cudaSuccess) e = cudaMemcpyAsync(buf, ctx->d_out + ranges[sg.range].offset, rlen, cudaMemcpyDeviceToHost, s);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&arrived[q], cudaEventDisableTiming);
                if (e == cudaSuccess) e = cudaEventRecord(arrived[q], s);
                if (e != cudaSuccess) {
                    failed = cuda_fail(e);
                    break;
                }
                cudaEvent_t ev = arrived[q];
                const int dev = ctx->device;
                uint8_t* comp_buf = buf + (largest + 255) / 256 * 256;
                workers.emplace_back([=, &seg_size, &seg_rc] {
                    if (cudaSetDevice(dev) != cudaSuccess || cudaEventSynchronize(ev) != cudaSuccess) {
                        seg_rc[q] = 3;
                        return;
                    }
                    size_t sz = 0;
                    seg_rc[q] = zstd_compressed_size(level, buf, rlen, comp_buf, max_comp, &sz);
                    seg_size[q] = sz;
                });
            }
            for (auto& w : workers) w.join();
            if (cudaStreamSynchronize(s) != cudaSuccess && failed == Outcome::kOk) failed = Outcome::kDevice;
        }
        for (cudaEvent_t ev : arrived)
            if (ev) cudaEventDestroy(ev);
        if (failed != Outcome::kOk) return failed;
        size_t best_size = SIZE_MAX;
        for (int i = 0; i < k; i++) {
            size_t total = 0;
            for (int r = 0; r < nr; r++) {
                const int q = plan.seg_of[(size_t)i * nr + r];
                if (seg_rc[q] != 0) return Outcome::kEstimator;
                total += seg_size[q];
            }
            if (total < best_size) best_size = total, best_s = order[i];
        }
        e = launch_transform(best_s, ctx->d_in, reference_layout(ctx->d_out, n, 0, best_s), n, s);
        if (e != cudaSuccess) return cuda_fail(e);
    } else {
        // Caller-supplied estimator: it sees host memory, exactly as in the reference — each
        // candidate is transformed on the GPU and only the estimated ranges travel back.
        size_t best_size = SIZE_MAX;
        int best_i = -1;
        for (int i = 0; i < k; i++) {
            e = launch_transform(order[i], ctx->d_in, reference_layout(ctx->d_out, n, 0, order[i]), n, s);
            for (int r = 0; r < nr && e == cudaSuccess; r++)
                e = cudaMemcpyAsync(out + ranges[r].offset, ctx->d_out + ranges[r].offset, ranges[r].len,
                                    cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) return cuda_fail(e);
            size_t total = 0;
            for (int r = 0; r < nr; r++) {
                size_t sz = 0;
                if (est.estimate_compressed_size(est.context, out + ranges[r].offset, ranges[r].len, comp, max_comp,
                                                 &sz) != 0)
                    return Outcome::kEstimator;
                total += sz;
            }
            if (total < best_size) best_size = total, best_s = order[i], best_i = i;
        }
        if (best_i != k - 1) {
            e = launch_transform(best_s, ctx->d_in, reference_layout(ctx->d_out, n, 0, best_s), n, s);
            if (e != cudaSuccess) return cuda_fail(e);
        }
    }
    e = cudaMemcpyAsync(out, ctx->d_out, len, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return cuda_fail(e);
    *best = best_s;
    return Outcome::kOk;
}

// ---- stable API bodies, shared by BC1 and BC2 ----------------------------------------------------
ManualBuilder* new_manual(int format) {
    // Bc1ManualTransformBuilder::new() = default settings (Variant1, split) — settings.rs:35-43
    return new (std::nothrow) ManualBuilder{format, kVariant1, true};
}

DltResult manual_run_impl(int format, bool inverse, const uint8_t* input, size_t input_len, uint8_t* output,
                         size_t output_len, ManualBuilder* b) {
    // manual_transform_builder.rs:256-287: input, output, builder null checks in this order.
    if (!input) return {kApiNullDataPointer};
    if (!output) return {kApiNullOutputBufferPointer};
    if (!b) return {kApiNullManualTransformBuilderPointer};
    const Settings st{format, b->variant, false, b->split_colour};
    return {api_code(transform_host(st, inverse, input, input_len, output, output_len))};
}

DltResult auto_run_impl(int format, AutoBuilder* b, const uint8_t* data, size_t data_len, uint8_t* output,
                       size_t output_len, ManualBuilder** out_manual) {
    // auto_transform_builder.rs:190-245: builder, data, output, out_manual_builder.
    if (!b) return {kApiNullBuilderPointer};
    if (!data) return {kApiNullDataPointer};
    if (!output) return {kApiNullOutputBufferPointer};
    if (!out_manual) return {kApiNullManualBuilderOutputPointer};
    Settings best{};
    const Outcome o = auto_host(format, data, data_len, output, output_len, b->estimator, b->use_all, &best);
    if (o != Outcome::kOk) {
        *out_manual = nullptr;
        return {api_code(o)};
    }
    ManualBuilder* m = new (std::nothrow) ManualBuilder{format, best.variant, best.split_colour};
    *out_manual = m;
    return {m ? kApiSuccess : kApiAllocationFailed};
}

const char* api_error_message(int32_t code, int format) {
    // error.rs:131-171 (BC2: "16 (BC2 block size)", Dltbc2* type names)
    const bool b1 = format == 1;
    switch (code) {
        case kApiSuccess: return "Success";
        case kApiInvalidLength:
            return b1 ? "Invalid input length: Length must be divisible by 8 (BC1 block size)"
                      : "Invalid input length: Length must be divisible by 16 (BC2 block size)";
        case kApiOutputBufferTooSmall: return "Output buffer too small for the operation";
        case kApiAllocationFailed: return "Memory allocation failed";
        case kApiSizeEstimationFailed: return "Size estimation failed during transform optimization";
        case kApiNullDataPointer: return "Null pointer provided for data parameter";
        case kApiNullEstimatorPointer: return "Null pointer provided for DltSizeEstimator parameter";
        case kApiNullTransformSettingsPointer:
            return b1 ? "Null pointer provided for Dltbc1TransformSettings parameter"
                      : "Null pointer provided for Dltbc2TransformSettings parameter";
        case kApiNullInputPointer: return "Null pointer provided for input parameter";
        case kApiNullOutputBufferPointer: return "Null pointer provided for output parameter";
        case kApiNullManualTransformBuilderPointer:
            return b1 ? "Null pointer provided for Dltbc1ManualTransformBuilder parameter"
                      : "Null pointer provided for Dltbc2ManualTransformBuilder parameter";
        case kApiNullBuilderPointer:
            return b1 ? "Null pointer provided for Dltbc1EstimateSettingsBuilder parameter"
                      : "Null pointer provided for Dltbc2EstimateSettingsBuilder parameter";
        case kApiNullManualBuilderOutputPointer: return "Null pointer provided for manual builder output parameter";
        default: return "Unknown error";
    }
}

// ---- core API bodies -------------------------------------------------------------------------------
DltResult core_run(int format, bool inverse, const uint8_t* input, size_t input_len, uint8_t* output,
                   size_t output_len, int variant, bool split_alpha, bool split_colour) {
    // core c_api/transform_with_settings.rs:73-100
    if (!input) return {kCoreNullDataPointer};
    if (!output) return {kCoreNullOutputBufferPointer};
    const Settings st{format, variant, split_alpha, split_colour};
    return {core_code(transform_host(st, inverse, input, input_len, output, output_len))};
}

DltResult core_auto(int format, const uint8_t* data, size_t data_len, uint8_t* output, size_t output_len,
                    const DltSizeEstimator* estimator, bool use_all, void* out_details, Settings* best) {
    // core c_api/transform_auto.rs:143-190
    if (!data) return {kCoreNullDataPointer};
    if (!output) return {kCoreNullOutputBufferPointer};
    if (!estimator) return {kCoreNullEstimatorPointer};
    if (!out_details) return {kCoreNullTransformSettingsPointer};
    return {core_code(auto_host(format, data, data_len, output, output_len, *estimator, use_all, best))};
}

bool to_settings(const DltcudaSettings& s, Settings* out) {
    if (s.format < 1 || s.format > 3 || s.decorrelation_mode > 3) return false;
    *out = Settings{s.format, s.decorrelation_mode, s.format == 3 && s.split_alpha_endpoints, s.split_colour_endpoints};
    return true;
}

}  // namespace

namespace dlt {
namespace cabi {
DltResult api_manual_run(int format, bool inverse, const uint8_t* input, size_t input_len, uint8_t* output,
                         size_t output_len, ManualBuilder* b) {
    return manual_run_impl(format, inverse, input, input_len, output, output_len, b);
}
DltResult api_auto_settings(int format, const AutoBuilder* b, const uint8_t* data, size_t data_len, uint8_t* output,
                            size_t output_len, Settings* best) {
    if (!b) return {kApiNullBuilderPointer};
    if (!data) return {kApiNullDataPointer};
    if (!output) return {kApiNullOutputBufferPointer};
    return {api_code(auto_host(format, data, data_len, output, output_len, b->estimator, b->use_all, best))};
}
}  // namespace cabi
}  // namespace dlt

// =================================================================================================
// 1a. Stable API: dltbc1_* / dltbc2_*
// =================================================================================================
#define DLT_DEFINE_STABLE_API(N)                                                                                      \
    DLT_EXPORT void* dltbc##N##_new_ManualTransformBuilder(void) { return new_manual(N); }                            \
    DLT_EXPORT void dltbc##N##_free_ManualTransformBuilder(void* b) { delete static_cast<ManualBuilder*>(b); }        \
    DLT_EXPORT void* dltbc##N##_clone_ManualTransformBuilder(const void* b) {                                         \
        if (!b) return nullptr;                                                                                       \
        return new (std::nothrow) ManualBuilder(*static_cast<const ManualBuilder*>(b));                               \
    }                                                                                                                 \
    DLT_EXPORT void dltbc##N##_ManualTransformBuilder_SetDecorrelationMode(void* b, uint8_t mode) {                   \
        if (!b || mode > 3) return;                                                                                   \
        static_cast<ManualBuilder*>(b)->variant = stable_to_internal(mode);                                           \
    }                                                                                                                 \
    DLT_EXPORT void dltbc##N##_ManualTransformBuilder_SetSplitColourEndpoints(void* b, bool split) {                  \
        if (!b) return;                                                                                               \
        static_cast<ManualBuilder*>(b)->split_colour = split;                                                         \
    }                                                                                                                 \
    DLT_EXPORT void dltbc##N##_ManualTransformBuilder_ResetToDefaults(void* b) {                                      \
        if (!b) return;                                                                                               \
        *static_cast<ManualBuilder*>(b) = ManualBuilder{N, kVariant1, true};                                          \
    }                                                                                                                 \
    DLT_EXPORT DltResult dltbc##N##_ManualTransformBuilder_Transform(const uint8_t* input, size_t input_len,          \
                                                                     uint8_t* output, size_t output_len, void* b) {   \
        return manual_run_impl(N, false, input, input_len, output, output_len, static_cast<ManualBuilder*>(b));        \
    }                                                                                                                 \
    DLT_EXPORT DltResult dltbc##N##_ManualTransformBuilder_Untransform(const uint8_t* input, size_t input_len,        \
                                                                       uint8_t* output, size_t output_len, void* b) { \
        return manual_run_impl(N, true, input, input_len, output, output_len, static_cast<ManualBuilder*>(b));         \
    }                                                                                                                 \
    DLT_EXPORT void* dltbc##N##_new_AutoTransformBuilder(const DltSizeEstimator* estimator) {                         \
        if (!estimator) return nullptr;                                                                               \
        return new (std::nothrow) AutoBuilder{N, *estimator, false};                                                  \
    }                                                                                                                 \
    DLT_EXPORT void dltbc##N##_free_AutoTransformBuilder(void* b) { delete static_cast<AutoBuilder*>(b); }            \
    DLT_EXPORT DltResult dltbc##N##_AutoTransformBuilder_SetUseAllDecorrelationModes(void* b, bool use_all) {         \
        if (!b) return {kApiNullBuilderPointer};                                                                      \
        static_cast<AutoBuilder*>(b)->use_all = use_all;                                                              \
        return {kApiSuccess};                                                                                         \
    }                                                                                                                 \
    DLT_EXPORT DltResult dltbc##N##_AutoTransformBuilder_Transform(void* b, const uint8_t* data, size_t data_len,     \
                                                                   uint8_t* output, size_t output_len,                \
                                                                   void** out_manual_builder) {                       \
        return auto_run_impl(N, static_cast<AutoBuilder*>(b), data, data_len, output, output_len,                      \
                            reinterpret_cast<ManualBuilder**>(out_manual_builder));                                   \
    }                                                                                                                 \
    DLT_EXPORT const char* dltbc##N##_error_message(int32_t code) { return api_error_message(code, N); }

DLT_DEFINE_STABLE_API(1)
DLT_DEFINE_STABLE_API(2)

// Introspection of a manual builder (additive; the reference exposes get_settings() only to Rust,
// manual_transform_builder.rs `get_settings`).  mode uses the STABLE numbering.
DLT_EXPORT int dltcuda_ManualTransformBuilder_GetSettings(const void* b, uint8_t* out_mode, bool* out_split) {
    if (!b) return 1;
    const ManualBuilder* m = static_cast<const ManualBuilder*>(b);
    if (out_mode) *out_mode = internal_to_stable(m->variant);
    if (out_split) *out_split = m->split_colour;
    return 0;
}

// =================================================================================================
// 1b. Core API: dltbc1core_* / dltbc2core_*   (+ 2. additive dltbc3core_*)
// =================================================================================================
#define DLT_DEFINE_CORE_API(N)                                                                                        \
    DLT_EXPORT DltResult dltbc##N##core_transform(const uint8_t* input, size_t input_len, uint8_t* output,            \
                                                  size_t output_len, DltCoreSettings details) {                       \
        return core_run(N, false, input, input_len, output, output_len, details.decorrelation_mode, false,            \
                        details.split_colour_endpoints);                                                              \
    }                                                                                                                 \
    DLT_EXPORT DltResult dltbc##N##core_untransform(const uint8_t* input, size_t input_len, uint8_t* output,          \
                                                    size_t output_len, DltCoreSettings details) {                     \
        return core_run(N, true, input, input_len, output, output_len, details.decorrelation_mode, false,             \
                        details.split_colour_endpoints);                                                              \
    }                                                                                                                 \
    DLT_EXPORT DltResult dltbc##N##core_transform_auto(const uint8_t* data, size_t data_len, uint8_t* output,         \
                                                       size_t output_len, const DltSizeEstimator* estimator,          \
                                                       DltCoreAutoSettings settings, DltCoreSettings* out_details) {  \
        Settings best{};                                                                                              \
        DltResult r = core_auto(N, data, data_len, output, output_len, estimator, settings.use_all_modes,             \
                                out_details, &best);                                                                  \
        if (r.error_code == kCoreSuccess) *out_details = DltCoreSettings{best.split_colour, (uint8_t)best.variant};   \
        return r;                                                                                                     \
    }

DLT_DEFINE_CORE_API(1)
DLT_DEFINE_CORE_API(2)

DLT_EXPORT DltResult dltbc3core_transform(const uint8_t* input, size_t input_len, uint8_t* output, size_t output_len,
                                          DltCoreBc3Settings d) {
    return core_run(3, false, input, input_len, output, output_len, d.decorrelation_mode, d.split_alpha_endpoints,
                    d.split_colour_endpoints);
}
DLT_EXPORT DltResult dltbc3core_untransform(const uint8_t* input, size_t input_len, uint8_t* output,
                                            size_t output_len, DltCoreBc3Settings d) {
    return core_run(3, true, input, input_len, output, output_len, d.decorrelation_mode, d.split_alpha_endpoints,
                    d.split_colour_endpoints);
}
DLT_EXPORT DltResult dltbc3core_transform_auto(const uint8_t* data, size_t data_len, uint8_t* output,
                                               size_t output_len, const DltSizeEstimator* estimator,
                                               DltCoreAutoSettings settings, DltCoreBc3Settings* out_details) {
    Settings best{};
    DltResult r = core_auto(3, data, data_len, output, output_len, estimator, settings.use_all_modes, out_details, &best);
    if (r.error_code == kCoreSuccess)
        *out_details = DltCoreBc3Settings{best.split_alpha, best.split_colour, (uint8_t)best.variant};
    return r;
}

// =================================================================================================
// 1c. LTU estimator factory (extensions/estimators/dxt-lossless-transform-ltu/src/c_api.rs:74-103)
// =================================================================================================
DLT_EXPORT DltSizeEstimator* dltltu_new_size_estimator(void) {
    static char ltu_instance;  // LosslessTransformUtilsSizeEstimation is a stateless unit struct
    DltSizeEstimator* e = new (std::nothrow) DltSizeEstimator;
    if (!e) return nullptr;
    e->context = &ltu_instance;
    e->max_compressed_size = &ltu_max_compressed_size;
    e->estimate_compressed_size = &ltu_estimate_compressed_size;
    return e;
}
DLT_EXPORT void dltltu_free_size_estimator(DltSizeEstimator* e) { delete e; }

// =================================================================================================
// 3. Additive device API: dltcuda_*
// =================================================================================================
enum DltcudaStatus : int {
    kDltcudaOk = 0,
    kDltcudaInvalidLength = 1,
    kDltcudaInvalidSettings = 2,
    kDltcudaCudaError = 3,
    kDltcudaNullPointer = 4,
    kDltcudaOutOfMemory = 5,
};

static int dltcuda_status(cudaError_t e) {
    if (e == cudaSuccess) return kDltcudaOk;
    note_cuda_error(e);
    return e == cudaErrorMemoryAllocation ? kDltcudaOutOfMemory : kDltcudaCudaError;
}
static int dltcuda_status(Status s) {
    return s == Status::kOk ? kDltcudaOk : s == Status::kOutOfMemory ? kDltcudaOutOfMemory : kDltcudaCudaError;
}

DLT_EXPORT int dltcuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}
DLT_EXPORT void dltcuda_set_device(int device) { set_thread_device(device); }
DLT_EXPORT const char* dltcuda_last_error(void) { return last_error_string(); }
DLT_EXPORT uint64_t dltcuda_kernel_launch_count(void) { return kernel_launch_count() + estimator_launch_count(); }

DLT_EXPORT size_t dltcuda_release_cached_memory(void) { return release_cached_memory(); }

DLT_EXPORT void* dltcuda_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    const int dev = thread_device();
    if (dev >= 0 && cudaSetDevice(dev) != cudaSuccess) return nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return p;
}
DLT_EXPORT void dltcuda_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

// Whole payload, device-resident, reference single-buffer layout.  Asynchronous on `stream`.
DLT_EXPORT int dltcuda_transform_device(const uint8_t* d_input, uint8_t* d_output, size_t len, DltcudaSettings s,
                                        void* stream) {
    Settings st;
    if (!to_settings(s, &st)) return kDltcudaInvalidSettings;
    if (len % (size_t)block_bytes(st.format)) return kDltcudaInvalidLength;
    if (len == 0) return kDltcudaOk;
    if (!d_input || !d_output) return kDltcudaNullPointer;
    const size_t n = len / block_bytes(st.format);
    return dltcuda_status(launch_transform(st, d_input, reference_layout(d_output, n, 0, st), n, (cudaStream_t)stream));
}
DLT_EXPORT int dltcuda_untransform_device(const uint8_t* d_input, uint8_t* d_output, size_t len, DltcudaSettings s,
                                          void* stream) {
    Settings st;
    if (!to_settings(s, &st)) return kDltcudaInvalidSettings;
    if (len % (size_t)block_bytes(st.format)) return kDltcudaInvalidLength;
    if (len == 0) return kDltcudaOk;
    if (!d_input || !d_output) return kDltcudaNullPointer;
    const size_t n = len / block_bytes(st.format);
    return dltcuda_status(launch_untransform(st, reference_layout(const_cast<uint8_t*>(d_input), n, 0, st), d_output, n,
                                             (cudaStream_t)stream));
}

// Block-range shard of a payload of `total_blocks` blocks: blocks [first_block, first_block+num_blocks).
//   transform  : d_blocks points at the shard's first block; d_streams_base is the base of the FULL
//                transformed image (reference layout for total_blocks); only this shard's slice of
//                every stream is written.
//   untransform: the mirror.
// No data crosses shards; the only shared quantities are total_blocks and first_block.
DLT_EXPORT int dltcuda_transform_device_range(const uint8_t* d_blocks, uint8_t* d_streams_base, size_t total_blocks,
                                              size_t first_block, size_t num_blocks, DltcudaSettings s, void* stream) {
    Settings st;
    if (!to_settings(s, &st)) return kDltcudaInvalidSettings;
    if (first_block > total_blocks || num_blocks > total_blocks - first_block) return kDltcudaInvalidLength;
    if (num_blocks == 0) return kDltcudaOk;
    if (!d_blocks || !d_streams_base) return kDltcudaNullPointer;
    return dltcuda_status(launch_transform(st, d_blocks, reference_layout(d_streams_base, total_blocks, first_block, st),
                                           num_blocks, (cudaStream_t)stream));
}
DLT_EXPORT int dltcuda_untransform_device_range(const uint8_t* d_streams_base, uint8_t* d_blocks, size_t total_blocks,
                                                size_t first_block, size_t num_blocks, DltcudaSettings s,
                                                void* stream) {
    Settings st;
    if (!to_settings(s, &st)) return kDltcudaInvalidSettings;
    if (first_block > total_blocks || num_blocks > total_blocks - first_block) return kDltcudaInvalidLength;
    if (num_blocks == 0) return kDltcudaOk;
    if (!d_blocks || !d_streams_base) return kDltcudaNullPointer;
    return dltcuda_status(launch_untransform(
        st, reference_layout(const_cast<uint8_t*>(d_streams_base), total_blocks, first_block, st), d_blocks, num_blocks,
        (cudaStream_t)stream));
}

// Explicit per-stream pointers (bcn_layout.h stream order; 2..6 streams): what a rank uses when it
// holds only its own shard, or when a caller wants every stream in its own allocation.
DLT_EXPORT int dltcuda_transform_device_streams(const uint8_t* d_blocks, uint8_t* const* d_streams, size_t num_blocks,
                                                DltcudaSettings s, void* stream) {
    Settings st;
    if (!to_settings(s, &st)) return kDltcudaInvalidSettings;
    if (num_blocks == 0) return kDltcudaOk;
    if (!d_blocks || !d_streams) return kDltcudaNullPointer;
    StreamPtrs sp{};
    for (int k = 0; k < num_streams(st.format, st.split_alpha, st.split_colour); k++) {
        if (!d_streams[k]) return kDltcudaNullPointer;
        sp.p[k] = d_streams[k];
    }
    return dltcuda_status(launch_transform(st, d_blocks, sp, num_blocks, (cudaStream_t)stream));
}
DLT_EXPORT int dltcuda_untransform_device_streams(const uint8_t* const* d_streams, uint8_t* d_blocks,
                                                  size_t num_blocks, DltcudaSettings s, void* stream) {
    Settings st;
    if (!to_settings(s, &st)) return kDltcudaInvalidSettings;
    if (num_blocks == 0) return kDltcudaOk;
    if (!d_blocks || !d_streams) return kDltcudaNullPointer;
    StreamPtrs sp{};
    for (int k = 0; k < num_streams(st.format, st.split_alpha, st.split_colour); k++) {
        if (!d_streams[k]) return kDltcudaNullPointer;
        sp.p[k] = const_cast<uint8_t*>(d_streams[k]);
    }
    return dltcuda_status(launch_untransform(st, sp, d_blocks, num_blocks, (cudaStream_t)stream));
}
// Number of streams / element width of stream k for a settings combination (layout introspection).
DLT_EXPORT int dltcuda_stream_count(DltcudaSettings s) {
    Settings st;
    return to_settings(s, &st) ? num_streams(st.format, st.split_alpha, st.split_colour) : 0;
}
DLT_EXPORT int dltcuda_stream_width(DltcudaSettings s, int k) {
    Settings st;
    if (!to_settings(s, &st) || k < 0 || k >= num_streams(st.format, st.split_alpha, st.split_colour)) return 0;
    return stream_width(st.format, st.split_alpha, st.split_colour, k);
}

// First block of shard `shard` of `num_shards` (shard == num_shards gives total_blocks): contiguous
// block ranges whose boundaries are multiples of the kernel tile, so every per-stream slice keeps
// the alignment of its stream base.  This host-side prefix is the whole multi-GPU "exchange".
DLT_EXPORT size_t dltcuda_shard_first_block(int format, size_t total_blocks, int shard, int num_shards) {
    if (num_shards <= 0 || shard <= 0) return 0;
    if (shard >= num_shards) return total_blocks;
    const size_t tile = (size_t)tile_blocks(format == 1 ? 1 : 2);
    const size_t tiles = (total_blocks + tile - 1) / tile;
    const size_t first = tiles * (size_t)shard / (size_t)num_shards * tile;
    return first < total_blocks ? first : total_blocks;
}

// split_color_endpoints on device-resident colour pairs (len_bytes % 4 == 0).  Asynchronous on `stream`.
DLT_EXPORT int dltcuda_split_color_endpoints_device(const uint8_t* d_colors, uint8_t* d_colors_out, size_t len_bytes,
                                                    void* stream) {
    if (len_bytes % 4) return kDltcudaInvalidLength;
    if (len_bytes == 0) return kDltcudaOk;
    if (!d_colors || !d_colors_out) return kDltcudaNullPointer;
    return dltcuda_status(launch_split_color_endpoints(d_colors, d_colors_out, len_bytes, (cudaStream_t)stream));
}
// The same on host buffers (synchronous).
DLT_EXPORT int dltcuda_split_color_endpoints(const uint8_t* colors, uint8_t* colors_out, size_t len_bytes) {
    if (len_bytes % 4) return kDltcudaInvalidLength;
    if (len_bytes == 0) return kDltcudaOk;
    if (!colors || !colors_out) return kDltcudaNullPointer;
    Status st;
    Context* ctx = acquire_context(-1, &st);
    if (!ctx) return dltcuda_status(st);
    st = ensure_device_buffers(ctx, len_bytes);
    cudaError_t e = cudaSuccess;
    if (st == Status::kOk) {
        cudaStream_t s = ctx->stream[0];
        e = cudaMemcpyAsync(ctx->d_in, colors, len_bytes, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = launch_split_color_endpoints(ctx->d_in, ctx->d_out, len_bytes, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(colors_out, ctx->d_out, len_bytes, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    release_context_synced(ctx);
    return st != Status::kOk ? dltcuda_status(st) : dltcuda_status(e);
}

// A batch of independent host payloads (mixed formats / settings), pipelined through the device(s):
// chunks of consecutive payloads overlap and there is one wait at the end.  With several devices the
// payloads are dealt out whole, largest-remaining-capacity first (payload-granular sharding: nothing
// is exchanged between devices); one host thread drives each device.
static int batch_impl_unguarded(const DltcudaPayload* payloads, size_t count, bool untransform, const int* devices,
                                int num_devices);
static int batch_impl(const DltcudaPayload* payloads, size_t count, bool untransform, const int* devices, int num_devices) {
    return guarded([&] { return batch_impl_unguarded(payloads, count, untransform, devices, num_devices); },
                              (int)kDltcudaOutOfMemory, (int)kDltcudaCudaError);
}
static int batch_impl_unguarded(const DltcudaPayload* payloads, size_t count, bool untransform, const int* devices,
                                int num_devices) {
    if (count == 0) return kDltcudaOk;
    if (!payloads || num_devices < 1) return kDltcudaNullPointer;
    std::vector<std::vector<HostJob>> jobs((size_t)num_devices);
    std::vector<size_t> load((size_t)num_devices, 0);
    for (size_t i = 0; i < count; i++) {
        Settings st;
        if (!to_settings(payloads[i].settings, &st)) return kDltcudaInvalidSettings;
        if (payloads[i].len % (size_t)block_bytes(st.format)) return kDltcudaInvalidLength;
        if (payloads[i].len && (!payloads[i].input || !payloads[i].output)) return kDltcudaNullPointer;
        size_t d = 0;
        for (size_t k = 1; k < load.size(); k++)
            if (load[k] < load[d]) d = k;
        load[d] += payloads[i].len;
        jobs[d].push_back(HostJob{st, untransform, payloads[i].input, payloads[i].output, payloads[i].len});
    }
    if (num_devices == 1)
        return dltcuda_status(run_host_batch(jobs[0].data(), jobs[0].size(), devices ? devices[0] : -1));
    std::vector<Status> results((size_t)num_devices, Status::kOk);
    std::vector<std::thread> threads;
    for (int d = 0; d < num_devices; d++)
        threads.emplace_back([&, d] { results[d] = run_host_batch(jobs[d].data(), jobs[d].size(), devices[d]); });
    for (auto& t : threads) t.join();
    for (Status r : results)
        if (r != Status::kOk) return dltcuda_status(r);
    return kDltcudaOk;
}

DLT_EXPORT int dltcuda_transform_batch(const DltcudaPayload* payloads, size_t count, bool untransform) {
    return batch_impl(payloads, count, untransform, nullptr, 1);
}
DLT_EXPORT int dltcuda_transform_batch_multi_gpu(const DltcudaPayload* payloads, size_t count, bool untransform,
                                                 const int* devices, int num_devices) {
    if (!devices) return kDltcudaNullPointer;
    return batch_impl(payloads, count, untransform, devices, num_devices);
}

// LTU-semantics estimate of a device-resident byte range.  Synchronous.
DLT_EXPORT int dltcuda_ltu_estimate_device(const uint8_t* d_data, size_t len, size_t* out_size) {
    if (!out_size) return kDltcudaNullPointer;
    if (!d_data || len == 0) {
        *out_size = 0;
        return kDltcudaOk;
    }
    Status st;
    Context* ctx = acquire_context(-1, &st);
    if (!ctx) return dltcuda_status(st);
    LtuSegment seg{d_data, len};
    uint64_t m = 0;
    st = ltu_matches_device(ctx, &seg, 1, &m, ctx->stream[0]);
    release_context_synced(ctx);
    if (st == Status::kOk) *out_size = ltu_estimate_from_matches(len, m);
    return dltcuda_status(st);
}

// The unverified parts of the restated LTU algorithm as run-time parameters (estimator.h LtuParams).
DLT_EXPORT int dltcuda_ltu_set_params(int hash_bits, bool index_from_top_bits, int group) {
    LtuParams p;
    p.hash_bits = hash_bits, p.index_top = index_from_top_bits, p.group = group;
    return ltu_set_params(p) ? kDltcudaOk : kDltcudaInvalidSettings;
}
DLT_EXPORT void dltcuda_ltu_get_params(int* hash_bits, bool* index_from_top_bits, int* group) {
    const LtuParams p = ltu_params();
    if (hash_bits) *hash_bits = p.hash_bits;
    if (index_from_top_bits) *index_from_top_bits = p.index_top;
    if (group) *group = p.group;
}

// transform_bcN_auto with the GPU LTU estimator on device-resident buffers.  out_estimates (optional)
// receives the per-candidate estimates in the reference's test order (4/8 for BC1,BC2; 8/16 for BC3).
// Synchronous; d_output holds the winner's transform on return.
DLT_EXPORT int dltcuda_transform_auto_device(int format, const uint8_t* d_input, uint8_t* d_output, size_t len,
                                             bool use_all_modes, DltcudaSettings* out_settings, size_t* out_estimates) {
    if (format < 1 || format > 3) return kDltcudaInvalidSettings;
    if (len % (size_t)block_bytes(format)) return kDltcudaInvalidLength;
    if (!out_settings || (len && (!d_input || !d_output))) return kDltcudaNullPointer;
    Status st;
    Context* ctx = acquire_context(-1, &st);
    if (!ctx) return dltcuda_status(st);
    Settings best{};
    st = auto_ltu_device(ctx, format, d_input, d_output, len, use_all_modes, &best, out_estimates, ctx->stream[0]);
    if (st == Status::kOk) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream[0]);
        if (e != cudaSuccess) st = Status::kCudaError, note_cuda_error(e);
    }
    release_context_synced(ctx);
    if (st == Status::kOk)
        *out_settings = DltcudaSettings{(uint8_t)format, (uint8_t)best.variant, best.split_alpha, best.split_colour};
    return dltcuda_status(st);
}

// transform_bcN_auto with the GPU LTU estimator for a BATCH of independent host payloads (a directory of textures):
// one upload per payload, ONE set of estimator launches for all candidates of all payloads, one download per payload.
// jobs[i].status receives that job's DltcudaStatus; the return value is the first failure (or Ok).
namespace dlt {
namespace cabi {
bool is_gpu_ltu_estimator(const DltSizeEstimator& e) { return is_gpu_ltu(e); }

int auto_batch_host(DltcudaAutoJob* jobs, size_t count, bool use_all) {
    if (count == 0) return kDltcudaOk;
    if (!jobs) return kDltcudaNullPointer;
    int first_error = kDltcudaOk;
    auto note = [&first_error](DltcudaAutoJob& j, int st) {
        j.status = st;
        if (st != kDltcudaOk && first_error == kDltcudaOk) first_error = st;
    };
    Status st;
    Context* ctx = acquire_context(-1, &st);
    if (!ctx) return dltcuda_status(st);
    struct Releaser {
        Context* c;
        ~Releaser() { release_context_synced(c); }
    } releaser{ctx};
    // Three queues: uploads, the search (transform candidates + estimator + winners), downloads.  The payloads are cut
    // into ROUNDS; round r+1 is uploaded and round r-1 downloaded while round r is searched (two device slots).
    cudaStream_t s_search = ctx->stream[0], s_in = ctx->stream[1], s_out = ctx->stream[2];
    struct Round {
        std::vector<size_t> idx, offset;
        size_t total = 0;
    };
    size_t all_bytes = 0;
    for (size_t i = 0; i < count; i++) all_bytes += jobs[i].len;
    // enough rounds to overlap the copies with the search, large enough to keep the estimator's launches busy
    const size_t round_bytes = std::min<size_t>((size_t)256 << 20, std::max<size_t>((size_t)16 << 20, all_bytes / 8));
    std::vector<Round> rounds;
    for (size_t i = 0; i < count; i++) {
        DltcudaAutoJob& j = jobs[i];
        j.status = kDltcudaOk;
        if (j.format < 1 || j.format > 3) { note(j, kDltcudaInvalidSettings); continue; }
        if (j.len % (size_t)block_bytes(j.format)) { note(j, kDltcudaInvalidLength); continue; }
        if (j.len && (!j.input || !j.output)) { note(j, kDltcudaNullPointer); continue; }
        const size_t padded = (j.len + 255) / 256 * 256;
        if (rounds.empty() || (!rounds.back().idx.empty() && rounds.back().total + padded > round_bytes)) rounds.emplace_back();
        Round& r = rounds.back();
        r.idx.push_back(i);
        r.offset.push_back(r.total);
        r.total += padded;
    }
    if (rounds.empty()) return first_error;
    size_t slot_bytes = 256;
    for (const Round& r : rounds) slot_bytes = std::max(slot_bytes, r.total);
    const int nslots = rounds.size() > 1 ? 2 : 1;
    if ((st = ensure_device_buffers(ctx, slot_bytes * nslots)) != Status::kOk) return dltcuda_status(st);

    struct InFlight {   // nothing may touch caller memory after the call returns, whatever the exit path
        cudaStream_t s[3];
        cudaEvent_t uploaded[2] = {}, downloaded[2] = {};
        ~InFlight() {
            for (cudaStream_t q : s) (void)cudaStreamSynchronize(q);
            for (int i = 0; i < 2; i++) {
                if (uploaded[i]) cudaEventDestroy(uploaded[i]);
                if (downloaded[i]) cudaEventDestroy(downloaded[i]);
            }
        }
    } fl{{s_search, s_in, s_out}};
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < nslots && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&fl.uploaded[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fl.downloaded[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return dltcuda_status(e);

    // Small payloads in page-locked memory (a directory of textures read into a pinned pool) do not pay a copy each:
    // their inputs are GATHERED by one kernel over the mapped host memory, and the winners' transforms write straight
    // into the mapped outputs.  Everything else goes through cudaMemcpyAsync.
    PinnedRanges pinned;
    // Up to 256 KiB per payload.  Measured, 64 payloads, BC1 fast, ms with / without: 64 KiB 0.93 / 1.38, 128 KiB 1.17 / 1.60,
    // 256 KiB 1.66 / 2.14, 512 KiB 2.79 / 2.43, 1 MiB 5.0 / 3.3 — once a payload is large the copy engines beat mapped reads
    // and writes, and their work overlaps with the search of the neighbouring rounds.
    static const size_t kMappedMax = [] {
        const char* v = std::getenv("DLTCUDA_MAPPED_SEARCH_MAX_KIB");
        const long kib = v ? std::atol(v) : 256;
        return kib >= 0 ? (size_t)kib << 10 : (size_t)256 << 10;
    }();
    std::vector<uint8_t*> in_dev(count, nullptr), out_dev(count, nullptr);
    size_t max_gather = 0;
    for (const Round& rd : rounds) {
        size_t g = 0;
        for (size_t i : rd.idx) {
            const DltcudaAutoJob& j = jobs[i];
            if (j.len == 0 || j.len > kMappedMax) continue;
            if ((reinterpret_cast<uintptr_t>(j.input) & 7) == 0 && pinned.covers(j.input, j.len, &in_dev[i])) g++;
            else in_dev[i] = nullptr;
            // (a 16-byte aligned base keeps every stream of the reference layout naturally aligned: the tiled kernels apply)
            if ((reinterpret_cast<uintptr_t>(j.output) & 15) != 0 || !pinned.covers(j.output, j.len, &out_dev[i])) out_dev[i] = nullptr;
        }
        max_gather = std::max(max_gather, g);
    }
    if (max_gather && (st = ensure_desc(ctx, 2 * max_gather * sizeof(CopyBatchItem))) != Status::kOk) return dltcuda_status(st);
    std::vector<CopyBatchItem> gather;
    auto upload = [&](size_t r) {
        const int slot = (int)(r % nslots);
        cudaError_t err = cudaSuccess;
        gather.clear();
        uint64_t gather_max = 0;
        for (size_t k = 0; k < rounds[r].idx.size() && err == cudaSuccess; k++) {
            const size_t i = rounds[r].idx[k];
            const DltcudaAutoJob& j = jobs[i];
            if (!j.len) continue;
            uint8_t* dst = ctx->d_in + slot * slot_bytes + rounds[r].offset[k];
            if (in_dev[i]) gather.push_back(CopyBatchItem{in_dev[i], dst, j.len}), gather_max = std::max<uint64_t>(gather_max, j.len);
            else err = cudaMemcpyAsync(dst, j.input, j.len, cudaMemcpyHostToDevice, s_in);
        }
        if (err == cudaSuccess && !gather.empty()) {
            // two descriptor areas, alternating with the slot: round r+1 is queued while round r's gather may still run
            CopyBatchItem* d = reinterpret_cast<CopyBatchItem*>(ctx->d_desc) + slot * max_gather;
            err = cudaMemcpyAsync(d, gather.data(), gather.size() * sizeof(CopyBatchItem), cudaMemcpyHostToDevice, s_in);
            if (err == cudaSuccess) err = launch_copy_batch(d, (int)gather.size(), gather_max, s_in);
        }
        if (err == cudaSuccess) err = cudaEventRecord(fl.uploaded[slot], s_in);
        return err;
    };
    if ((e = upload(0)) != cudaSuccess) return dltcuda_status(e);
    for (size_t r = 0; r < rounds.size(); r++) {
        const int slot = (int)(r % nslots);
        const Round& rd = rounds[r];
        // the other slot's input was last read by round r-1's search, which has completed (the search ends with a wait)
        if (r + 1 < rounds.size() && (e = upload(r + 1)) != cudaSuccess) return dltcuda_status(e);
        e = cudaStreamWaitEvent(s_search, fl.uploaded[slot], 0);
        // this slot's output may still be on its way to the host (round r-2)
        if (e == cudaSuccess && r >= 2) e = cudaStreamWaitEvent(s_search, fl.downloaded[slot], 0);
        if (e != cudaSuccess) return dltcuda_status(e);
        std::vector<AutoJob> aj(rd.idx.size());
        for (size_t k = 0; k < rd.idx.size(); k++) {
            const DltcudaAutoJob& j = jobs[rd.idx[k]];
            uint8_t* winner = out_dev[rd.idx[k]] ? out_dev[rd.idx[k]] : ctx->d_out + slot * slot_bytes + rd.offset[k];
            aj[k] = AutoJob{j.format, ctx->d_in + slot * slot_bytes + rd.offset[k], winner, j.len, Settings{}, {}};
        }
        if ((st = auto_ltu_device_batch(ctx, aj.data(), (int)aj.size(), use_all, s_search)) != Status::kOk) return dltcuda_status(st);
        for (size_t k = 0; k < rd.idx.size() && e == cudaSuccess; k++) {
            DltcudaAutoJob& j = jobs[rd.idx[k]];
            j.out_settings = DltcudaSettings{j.format, (uint8_t)aj[k].best.variant, aj[k].best.split_alpha, aj[k].best.split_colour};
            if (j.len && !out_dev[rd.idx[k]])
                e = cudaMemcpyAsync(j.output, ctx->d_out + slot * slot_bytes + rd.offset[k], j.len, cudaMemcpyDeviceToHost, s_out);
        }
        if (e == cudaSuccess) e = cudaEventRecord(fl.downloaded[slot], s_out);
        if (e != cudaSuccess) return dltcuda_status(e);
    }
    e = cudaStreamSynchronize(s_out);
    if (e != cudaSuccess) return dltcuda_status(e);
    return first_error;
}
}  // namespace cabi
}  // namespace dlt

DLT_EXPORT int dltcuda_transform_auto_batch(DltcudaAutoJob* jobs, size_t count, bool use_all_modes) {
    return guarded([&] { return dlt::cabi::auto_batch_host(jobs, count, use_all_modes); }, (int)kDltcudaOutOfMemory,
                              (int)kDltcudaCudaError);
}

// The same over several GPUs of one box: whole payloads are dealt out (least-loaded device first), one host thread
// per device runs its share through dltcuda_transform_auto_batch; nothing is exchanged between devices (an estimate
// is a property of a whole payload, so payloads - never one payload's streams - are what gets sharded).
static int auto_batch_multi_gpu_unguarded(DltcudaAutoJob* jobs, size_t count, bool use_all_modes, const int* devices,
                                          int num_devices);
DLT_EXPORT int dltcuda_transform_auto_batch_multi_gpu(DltcudaAutoJob* jobs, size_t count, bool use_all_modes,
                                                      const int* devices, int num_devices) {
    return guarded([&] { return auto_batch_multi_gpu_unguarded(jobs, count, use_all_modes, devices, num_devices); },
                   (int)kDltcudaOutOfMemory, (int)kDltcudaCudaError);
}
static int auto_batch_multi_gpu_unguarded(DltcudaAutoJob* jobs, size_t count, bool use_all_modes, const int* devices,
                                          int num_devices) {
    if (count == 0) return kDltcudaOk;
    if (!jobs || !devices || num_devices < 1) return kDltcudaNullPointer;
    std::vector<std::vector<size_t>> share((size_t)num_devices);
    std::vector<size_t> load((size_t)num_devices, 0);
    for (size_t i = 0; i < count; i++) {
        size_t d = 0;
        for (size_t k = 1; k < load.size(); k++)
            if (load[k] < load[d]) d = k;
        load[d] += jobs[i].len + 1;
        share[d].push_back(i);
    }
    std::vector<int> rc((size_t)num_devices, kDltcudaOk);
    std::vector<std::thread> threads;
    for (int d = 0; d < num_devices; d++)
        threads.emplace_back([&, d] {
            std::vector<DltcudaAutoJob> mine;
            for (size_t i : share[d]) mine.push_back(jobs[i]);
            set_thread_device(devices[d]);
            rc[d] = dlt::cabi::auto_batch_host(mine.data(), mine.size(), use_all_modes);
            for (size_t k = 0; k < mine.size(); k++) jobs[share[d][k]] = mine[k];
        });
    for (auto& th : threads) th.join();
    for (int r : rc)
        if (r != kDltcudaOk) return r;
    return kDltcudaOk;
}

// =================================================================================================
// experimental::normalize_blocks (BC1) — additive, the reference has no C ABI for its experimental module
// (core/dxt-lossless-transform-bc1/src/experimental/normalize_blocks/{normalize.rs,transform.rs}).
// mode: ColorNormalizationMode, None = 0, Color0Only = 1, ReplicateColor = 2.
// =================================================================================================
namespace {

// Host buffers through the device for the stand-alone normalization passes (small helper: one upload, one launch,
// one download per output; these passes exist for completeness — transform_with_normalize_blocks fuses them).
int normalize_host(const uint8_t* input, uint8_t* const outs[3], size_t len, bool* any_normalized) {
    if (len % 8) return kDltcudaInvalidLength;
    if (any_normalized) *any_normalized = false;
    if (len == 0) return kDltcudaOk;
    if (!input) return kDltcudaNullPointer;
    int nout = 0;
    for (int m = 0; m < 3; m++) nout += outs[m] != nullptr;
    Status st;
    Context* ctx = acquire_context(-1, &st);
    if (!ctx) return dltcuda_status(st);
    struct Releaser {
        Context* c;
        ~Releaser() { release_context_synced(c); }
    } releaser{ctx};
    if ((st = ensure_device_buffers(ctx, len)) != Status::kOk) return dltcuda_status(st);
    if ((st = ensure_scratch(ctx, 3 * ((len + 255) / 256 * 256) + 256)) != Status::kOk) return dltcuda_status(st);
    cudaStream_t s = ctx->stream[0];
    const size_t img = (len + 255) / 256 * 256;
    unsigned int* d_any = reinterpret_cast<unsigned int*>(ctx->d_scratch + 3 * img);
    uint8_t* d_out[3];
    for (int m = 0; m < 3; m++) d_out[m] = outs[m] ? ctx->d_scratch + (size_t)m * img : nullptr;
    unsigned int any = 0;
    cudaError_t e = cudaMemcpyAsync(ctx->d_in, input, len, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_any, 0, sizeof(unsigned int), s);
    if (e == cudaSuccess) e = launch_normalize_blocks(ctx->d_in, d_out[0], d_out[1], d_out[2], len / 8, d_any, s);
    for (int m = 0; m < 3 && e == cudaSuccess; m++)
        if (outs[m]) e = cudaMemcpyAsync(outs[m], d_out[m], len, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&any, d_any, sizeof(any), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return dltcuda_status(e);
    if (any_normalized) *any_normalized = any != 0;
    (void)nout;
    return kDltcudaOk;
}

}  // namespace

// normalize_blocks (normalize.rs:38): output may equal input (in place).
DLT_EXPORT int dltcuda_bc1_normalize_blocks(const uint8_t* input, uint8_t* output, size_t len, int mode) {
    if (mode < 0 || mode > 2) return kDltcudaInvalidSettings;
    if (len % 8) return kDltcudaInvalidLength;
    if (len && (!input || !output)) return kDltcudaNullPointer;
    if (mode == kNormNone) {   // normalize.rs:56-66: a plain copy (nothing at all when in place)
        if (len && input != output) std::memmove(output, input, len);
        return kDltcudaOk;
    }
    uint8_t* outs[3] = {nullptr, nullptr, nullptr};
    outs[mode] = output;
    return normalize_host(input, outs, len, nullptr);
}
// normalize_blocks_all_modes (normalize.rs:417): one pass, three outputs; *any_normalized = its return value.
DLT_EXPORT int dltcuda_bc1_normalize_blocks_all_modes(const uint8_t* input, uint8_t* out_none, uint8_t* out_color0_only,
                                                      uint8_t* out_replicate_color, size_t len, bool* any_normalized) {
    if (len && (!out_none || !out_color0_only || !out_replicate_color)) return kDltcudaNullPointer;
    uint8_t* outs[3] = {out_none, out_color0_only, out_replicate_color};
    return normalize_host(input, outs, len, any_normalized);
}
// normalize_split_blocks_in_place (normalize.rs:286): colours ([c0 c1] per block) and indices in separate arrays.
DLT_EXPORT int dltcuda_bc1_normalize_split_blocks_in_place(uint8_t* colors, uint8_t* indices, size_t num_blocks, int mode) {
    if (mode < 0 || mode > 2) return kDltcudaInvalidSettings;
    if (num_blocks == 0 || mode == 0) return kDltcudaOk;
    if (!colors || !indices) return kDltcudaNullPointer;
    Status st;
    Context* ctx = acquire_context(-1, &st);
    if (!ctx) return dltcuda_status(st);
    struct Releaser {
        Context* c;
        ~Releaser() { release_context_synced(c); }
    } releaser{ctx};
    const size_t bytes = num_blocks * 4;
    if ((st = ensure_device_buffers(ctx, bytes)) != Status::kOk) return dltcuda_status(st);
    cudaStream_t s = ctx->stream[0];
    cudaError_t e = cudaMemcpyAsync(ctx->d_in, colors, bytes, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_out, indices, bytes, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = launch_normalize_split_blocks(ctx->d_in, ctx->d_out, num_blocks, mode, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(colors, ctx->d_in, bytes, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(indices, ctx->d_out, bytes, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    return dltcuda_status(e);
}
// Device-resident normalize_blocks (asynchronous on `stream`; d_output may equal d_input).
DLT_EXPORT int dltcuda_bc1_normalize_blocks_device(const uint8_t* d_input, uint8_t* d_output, size_t len, int mode, void* stream) {
    if (mode < 0 || mode > 2) return kDltcudaInvalidSettings;
    if (len % 8) return kDltcudaInvalidLength;
    if (len == 0) return kDltcudaOk;
    if (!d_input || !d_output) return kDltcudaNullPointer;
    if (mode == kNormNone)
        return d_input == d_output ? kDltcudaOk
                                   : dltcuda_status(cudaMemcpyAsync(d_output, d_input, len, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    uint8_t* outs[3] = {nullptr, nullptr, nullptr};
    outs[mode] = d_output;
    return dltcuda_status(launch_normalize_blocks(d_input, outs[0], outs[1], outs[2], len / 8, nullptr, (cudaStream_t)stream));
}

// transform_bc1_with_normalize_blocks (transform.rs:65): normalization fused into the transform kernel — one pass.
// Host pointers (synchronous) and device pointers (asynchronous).  Untransform with dltbc1core_untransform /
// dltcuda_untransform_device and the same decorrelation mode + split flag (normalization is not undone).
DLT_EXPORT int dltcuda_bc1_transform_with_normalize_blocks(const uint8_t* input, uint8_t* output, size_t len,
                                                           int normalization_mode, uint8_t decorrelation_mode,
                                                           bool split_colour_endpoints) {
    if (normalization_mode < 0 || normalization_mode > 2 || decorrelation_mode > 3) return kDltcudaInvalidSettings;
    if (len % 8) return kDltcudaInvalidLength;
    if (len == 0) return kDltcudaOk;
    if (!input || !output) return kDltcudaNullPointer;
    Settings st{1, decorrelation_mode, false, split_colour_endpoints};
    st.normalize = normalization_mode;
    return dltcuda_status(run_host(st, false, input, output, len, -1));
}
DLT_EXPORT int dltcuda_bc1_transform_with_normalize_blocks_device(const uint8_t* d_input, uint8_t* d_output, size_t len,
                                                                  int normalization_mode, uint8_t decorrelation_mode,
                                                                  bool split_colour_endpoints, void* stream) {
    if (normalization_mode < 0 || normalization_mode > 2 || decorrelation_mode > 3) return kDltcudaInvalidSettings;
    if (len % 8) return kDltcudaInvalidLength;
    if (len == 0) return kDltcudaOk;
    if (!d_input || !d_output) return kDltcudaNullPointer;
    Settings st{1, decorrelation_mode, false, split_colour_endpoints};
    st.normalize = normalization_mode;
    return dltcuda_status(launch_transform(st, d_input, reference_layout(d_output, len / 8, 0, st), len / 8, (cudaStream_t)stream));
}

// transform_bc1_auto_with_normalization (transform.rs:222) with the GPU LTU estimator, host pointers, synchronous.
// out_estimates (optional, >= 24 entries): per-candidate estimates, normalization mode outermost (only when a block was
// normalizable; otherwise the plain search's 4 / 8 entries).
DLT_EXPORT int dltcuda_bc1_transform_auto_with_normalization(const uint8_t* input, uint8_t* output, size_t len,
                                                             bool use_all_modes, int* out_normalization_mode,
                                                             uint8_t* out_decorrelation_mode, bool* out_split_colour_endpoints,
                                                             size_t* out_estimates) {
    if (len % 8) return kDltcudaInvalidLength;
    if (!out_normalization_mode || !out_decorrelation_mode || !out_split_colour_endpoints) return kDltcudaNullPointer;
    if (len && (!input || !output)) return kDltcudaNullPointer;
    Settings best{};
    if (len == 0) {   // no block is normalizable: the plain search on an empty payload picks its first candidate
        Settings order[kMaxCandidates];
        candidate_order(1, use_all_modes, order);
        best = order[0];
    } else {
        Status st;
        Context* ctx = acquire_context(-1, &st);
        if (!ctx) return dltcuda_status(st);
        struct Releaser {
            Context* c;
            ~Releaser() { release_context_synced(c); }
        } releaser{ctx};
        if ((st = ensure_device_buffers(ctx, len)) != Status::kOk) return dltcuda_status(st);
        cudaStream_t s = ctx->stream[0];
        cudaError_t e = cudaMemcpyAsync(ctx->d_in, input, len, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return dltcuda_status(e);
        if ((st = auto_ltu_norm_device(ctx, ctx->d_in, ctx->d_out, len, use_all_modes, &best, out_estimates, s)) != Status::kOk)
            return dltcuda_status(st);
        e = cudaMemcpyAsync(output, ctx->d_out, len, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return dltcuda_status(e);
    }
    *out_normalization_mode = best.normalize;
    *out_decorrelation_mode = (uint8_t)best.variant;
    *out_split_colour_endpoints = best.split_colour;
    return kDltcudaOk;
}

// Candidate order of the search, for callers that want to label out_estimates.  Returns the count.
DLT_EXPORT int dltcuda_auto_candidates(int format, bool use_all_modes, DltcudaSettings* out /* >= 16 entries */) {
    if (format < 1 || format > 3 || !out) return 0;
    Settings order[kMaxCandidates];
    const int k = candidate_order(format, use_all_modes, order);
    for (int i = 0; i < k; i++)
        out[i] = DltcudaSettings{(uint8_t)format, (uint8_t)order[i].variant, order[i].split_alpha, order[i].split_colour};
    return k;
}
