// auto_search.h — transform_bcN_auto: brute-force search for the best transform settings.
//
// Reference: core/dxt-lossless-transform-bc1/src/transform/transform_auto.rs:200-270,
//            core/dxt-lossless-transform-bc2/src/transform/transform_auto.rs:196-272,
//            core/dxt-lossless-transform-bc3/src/transform/transform_auto.rs:196-293.
// Candidates are tried in the reference's test order; the estimate of a candidate is taken over the
// endpoint streams only (BC1: out[0,len/2); BC2: out[len/2, len/2+len/4); BC3: out[0,2N) plus
// out[len/2, len/2+4N)); a candidate replaces the best only on a strictly smaller estimate, so the
// first candidate in order wins ties; the output finally holds the winner's transform.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <vector>

#include "bcn_layout.h"
#include "host_pipeline.h"

namespace dlt {

constexpr int kMaxCandidates = 16;

// Test orders (bc1/bc2 settings.rs:81-98, bc3 settings.rs:91-121).  Returns the candidate count.
int candidate_order(int format, bool use_all_modes, Settings out[kMaxCandidates]);

// Bc{1,2,3}TransformSettings::default() (bc1 settings.rs:35-43, bc3 settings.rs:39-48).
Settings default_settings(int format);

// The byte ranges of the transformed payload an estimate covers (1 for BC1/BC2, 2 for BC3).
struct EstimateRange {
    size_t offset, len;
};
int estimate_ranges(int format, size_t len, EstimateRange out[2]);

// Which candidates really differ in an estimated range (see auto_search.cu): segs = the distinct ranges, each held by
// the image of candidate `cand`; seg_of[cand * nr + range] = its index in segs; images = the candidates to transform.
struct DistinctPlan {
    struct Seg {
        int cand, range;
    };
    std::vector<Seg> segs;
    std::vector<int> seg_of;
    std::vector<int> images;
};
DistinctPlan plan_distinct(int format, const Settings* order, int k, int nr);

// Device-resident search with the GPU LTU estimator.  d_in: len bytes of blocks; d_out: len bytes,
// holds the winner's transform on return.  `sizes` (optional) receives the per-candidate estimates
// in test order.  Synchronises `stream`.
Status auto_ltu_device(Context* ctx, int format, const uint8_t* d_in, uint8_t* d_out, size_t len, bool use_all_modes,
                       Settings* best, size_t* sizes, cudaStream_t stream);

// experimental::transform_bc1_auto_with_normalization (core/dxt-lossless-transform-bc1/src/experimental/
// normalize_blocks/transform.rs:222-340) with the GPU LTU estimator.  best->normalize receives the winning
// ColorNormalizationMode; `sizes` (optional, 3 * kMaxCandidates entries) the estimates, normalization mode outermost
// (only filled when at least one block is normalizable; otherwise the plain search's k entries).
Status auto_ltu_norm_device(Context* ctx, const uint8_t* d_in, uint8_t* d_out, size_t len, bool use_all_modes,
                            Settings* best, size_t* sizes, cudaStream_t stream);

// The same search for MANY independent device-resident payloads at once (a directory of textures): the candidates of
// all payloads are transformed into scratch images and every endpoint stream of every candidate becomes one segment
// of a single estimator call, so the launch count does not grow with the number of payloads.  Jobs that do not fit
// the scratch budget together are processed in consecutive groups; a job too large for a group of its own goes
// through auto_ltu_device.  d_out of every job holds its winner's transform on return.  Synchronises `stream`.
struct AutoJob {
    int format;
    const uint8_t* d_in;
    uint8_t* d_out;
    size_t len;
    Settings best;                 // out
    size_t sizes[kMaxCandidates];  // out: per-candidate estimates in test order
};
Status auto_ltu_device_batch(Context* ctx, AutoJob* jobs, int njobs, bool use_all_modes, cudaStream_t stream);

}  // namespace dlt
