// estimator.h — GPU size estimator with the semantics of the reference's LTU estimator.
//
// Reference boundary: LosslessTransformUtilsSizeEstimation::estimate_compressed_size
// (extensions/estimators/dxt-lossless-transform-ltu/src/lib.rs:67-119):
//     estimate(data) = data.len().saturating_sub(estimate_num_lz_matches_fast(data)),  0 if empty
// where estimate_num_lz_matches_fast lives in the third-party crate lossless-transform-utils 0.1.3
// (not under /root/reference; restated — PARITY UNPINNED, see DESIGN.md and the LTU_* constants).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "host_pipeline.h"

namespace dlt {

// The five constants of the restated algorithm (kept in one place; oracle/ltu_params.h mirrors them).
constexpr int kLtuHashBits = 16;
constexpr uint32_t kLtuGoldenRatio = 0x9E3779B1u;
constexpr uint32_t kLtuKeyMask = 0x00FFFFFFu;
constexpr int kLtuGroup = 4;      // positions compared, then written, per loop iteration
constexpr int kLtuTailGuard = 7;  // loop runs while i < len.saturating_sub(7)

struct LtuSegment {
    const uint8_t* d_ptr;  // device pointer
    size_t len;
};

// Device scratch the call below needs for these segments (sort buffers: ~8.2 bytes per position).
size_t ltu_scratch_bytes(const LtuSegment* segs, int nseg);

// The same figure kept incrementally while a caller appends segments (launch sets are formed in input order, so
// appending a segment only changes the last, still open set): O(1) per segment instead of re-planning all of them.
// bytes() == ltu_scratch_bytes() of the segments added so far.  Cheap to copy (try a job, keep or drop the copy).
class LtuScratchMeter {
public:
    void add(size_t len);
    size_t bytes() const;

private:
    static constexpr int kSet = 256;   // == kMaxSegs of estimator.cu (static_assert there)
    size_t nseg_ = 0, closed_bytes_ = 0, open_len_[kSet] = {};
    int open_ = 0;
};

// Number of LZ matches of each device-resident segment, written to host `matches[0..nseg)`.
// `scratch` is device memory of at least ltu_scratch_bytes().  Synchronises `stream` before returning.
Status ltu_matches_device(const LtuSegment* segs, int nseg, uint64_t* matches, cudaStream_t stream, uint8_t* scratch,
                          size_t scratch_bytes);

// Convenience: takes the scratch from the context (grows ctx->d_scratch as needed).
inline Status ltu_matches_device(Context* ctx, const LtuSegment* segs, int nseg, uint64_t* matches,
                                 cudaStream_t stream) {
    const size_t need = ltu_scratch_bytes(segs, nseg);
    const Status st = ensure_scratch(ctx, need);
    if (st != Status::kOk) return st;
    return ltu_matches_device(segs, nseg, matches, stream, ctx->d_scratch, ctx->d_scratch_cap);
}

inline size_t ltu_estimate_from_matches(size_t len, uint64_t matches) {
    return len == 0 ? 0 : (matches >= len ? 0 : len - (size_t)matches);
}

uint64_t estimator_launch_count();

}  // namespace dlt
