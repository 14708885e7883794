// estimator.h — GPU size estimator with the semantics of the reference's LTU estimator.
//
// Reference boundary: LosslessTransformUtilsSizeEstimation::estimate_compressed_size
// (extensions/estimators/dxt-lossless-transform-ltu/src/lib.rs:67-119):
//     estimate(data) = data.len().saturating_sub(estimate_num_lz_matches_fast(data)),  0 if empty
// where estimate_num_lz_matches_fast lives in the third-party crate lossless-transform-utils 0.1.3
// (not under /root/reference; restated — PARITY UNPINNED, see DESIGN.md and LtuParams below).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "host_pipeline.h"

namespace dlt {

// Fixed parts of the restated algorithm (oracle/ltu_params.h mirrors them).
constexpr uint32_t kLtuGoldenRatio = 0x9E3779B1u;
constexpr uint32_t kLtuKeyMask = 0x00FFFFFFu;
constexpr int kLtuTailGuard = 7;  // loop runs while i < len.saturating_sub(7)

// The parts of the restatement that could not be checked against the crate's source are RUN-TIME parameters of
// the estimator (and of the oracle, oracle/bcn_oracle.c orc_ltu_num_lz_matches_params): the day the crate's
// source is at hand, parity is a call to ltu_set_params(), not a redesign.
//   hash_bits  : log2 of the table size, 12..17
//   index_top  : true  -> index = (key * GOLDEN) >> (32 - hash_bits)
//                false -> index = (key * GOLDEN) & ((1 << hash_bits) - 1)
//   group      : positions per loop iteration (all compares of a group see the table as it was before the group,
//                then the group's updates are applied in order): 4 or 1
struct LtuParams {
    int hash_bits = 16;
    bool index_top = true;
    int group = 4;
};
bool ltu_params_supported(const LtuParams& p);
bool ltu_set_params(const LtuParams& p);   // process-wide; false (and no change) if unsupported
LtuParams ltu_params();

struct LtuSegment {
    const uint8_t* d_ptr;  // device pointer
    size_t len;
};

// Device scratch the call below needs for these segments (chunk descriptors and, for segments that are cut into
// several chunks, 640 KiB of hand-over state per chunk: far below one byte per position).
size_t ltu_scratch_bytes(const LtuSegment* segs, int nseg);

// An upper bound of the same figure kept incrementally while a caller appends segments: O(1) per segment.
// bytes() >= ltu_scratch_bytes() of the segments added so far.  Cheap to copy (try a job, keep or drop the copy).
class LtuScratchMeter {
public:
    void add(size_t len);
    size_t bytes() const;

private:
    size_t nseg_ = 0, batches_ = 0;
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
