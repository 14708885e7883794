// estimator.cu — LTU-semantics LZ match estimator on the GPU (see estimator.h).
//
// The restated CPU algorithm is a sequential scan with a 2^16-entry "last 3-byte key seen in this
// bucket" table, processed 4 positions at a time (4 compares against the table as it was before
// the group, then 4 updates).  Position p is a match iff the most recent earlier position q of the
// same bucket that lies in an EARLIER group of four holds the same key (an untouched bucket holds 0).
//
// Buckets never interact, so the scan parallelises over BUCKET GROUPS: a warp owns a contiguous
// range of buckets, sees the positions that hash into its range in stream order, 32 at a time, and
// resolves the order inside a step with __match_any_sync.  Per bucket it keeps two words:
//   last = key of the most recent position of the bucket,
//   base = what the group of four containing that position compared against (the table state
//          before the group started).
// A position that has `nskip` same-bucket predecessors inside its own group of four compares
// against the (nskip+1)-th most recent position of the bucket, i.e. a lower lane of the step, or
// `last`, or — when its group started in an earlier step — `base`.
#include "estimator.h"

#include <atomic>

namespace dlt {
namespace {

std::atomic<uint64_t> g_est_launches{0};

constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxSegsPerLaunch = 8;

struct SegBatch {
    LtuSegment s[kMaxSegsPerLaunch];
};

// Number of positions the reference loop visits: groups of 4 starting at i = 0,4,.. while i < len-7.
__host__ __device__ inline size_t ltu_positions(size_t len) {
    const size_t end = len > (size_t)kLtuTailGuard ? len - kLtuTailGuard : 0;
    return (end + kLtuGroup - 1) / kLtuGroup * kLtuGroup;
}

__device__ __forceinline__ uint32_t ltu_bucket(uint32_t key) { return (key * kLtuGoldenRatio) >> (32 - kLtuHashBits); }

// One warp step over up to 32 positions of this warp's bucket range, in stream order by lane.
// Returns the number of matches in the step (same value in every lane).
__device__ __forceinline__ int consume_step(bool valid, uint32_t b, uint32_t key, int nskip, uint32_t* t_last,
                                            uint32_t* t_base) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned mask = __match_any_sync(kFull, valid ? b : (0x80000000u | lane));
    const unsigned lower = mask & ((1u << lane) - 1u);
    const int r = __popc(lower);
    unsigned m = lower;
#pragma unroll
    for (int i = 0; i < kLtuGroup - 1; i++)
        if (i < nskip && m) m &= ~(0x80000000u >> __clz(m));
    const int src = m ? 31 - __clz(m) : (int)lane;
    const uint32_t from_lane = __shfl_sync(kFull, key, src);
    uint32_t cmp = 0;
    if (valid) cmp = nskip < r ? from_lane : (nskip == r ? t_last[b] : t_base[b]);
    const bool match = valid && key == cmp;
    __syncwarp();
    if (valid && (mask >> lane) == 1u) {  // highest lane of this bucket in the step
        t_last[b] = key;
        t_base[b] = cmp;
    }
    __syncwarp();
    return __popc(__ballot_sync(kFull, match));
}

// ---- v0: every warp scans the whole segment and keeps the positions of its bucket range ---------
constexpr int kScanGroups = 64;
constexpr int kScanBuckets = (1 << kLtuHashBits) / kScanGroups;

__global__ void __launch_bounds__(32) ltu_scan_filter_kernel(const SegBatch batch, unsigned long long* matches) {
    __shared__ uint32_t t_last[kScanBuckets], t_base[kScanBuckets];
    const unsigned lane = threadIdx.x;
    const uint32_t g = blockIdx.x;
    const LtuSegment seg = batch.s[blockIdx.y];
    for (int i = lane; i < kScanBuckets; i += 32) t_last[i] = t_base[i] = 0u;
    __syncwarp();
    const size_t npos = ltu_positions(seg.len);
    const uint8_t* d = seg.d_ptr;
    unsigned long long count = 0;
    for (size_t base = 0; base < npos; base += 32) {
        const size_t p = base + lane;
        const bool inb = p < npos;
        uint32_t key = 0;
        if (inb) key = (uint32_t)d[p] | ((uint32_t)d[p + 1] << 8) | ((uint32_t)d[p + 2] << 16);
        const uint32_t bucket = ltu_bucket(key);
        int nskip = 0;
#pragma unroll
        for (int k = 1; k < kLtuGroup; k++) {
            const uint32_t bk = __shfl_up_sync(kFull, bucket, k);
            if ((int)(lane & (kLtuGroup - 1)) >= k && bk == bucket) nskip++;
        }
        const bool mine = inb && (bucket / kScanBuckets) == g;
        count += consume_step(mine, bucket % kScanBuckets, key, nskip, t_last, t_base);
    }
    if (lane == 0 && count) atomicAdd(&matches[blockIdx.y], count);
}

}  // namespace

uint64_t estimator_launch_count() { return g_est_launches.load(std::memory_order_relaxed); }

Status ltu_matches_device(Context* ctx, const LtuSegment* segs, int nseg, uint64_t* matches, cudaStream_t stream) {
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
    Status st = ensure_scratch(ctx, 4096);
    if (st != Status::kOk) return st;
    unsigned long long* d_matches = reinterpret_cast<unsigned long long*>(ctx->d_scratch);
    for (int first = 0; first < nseg; first += kMaxSegsPerLaunch) {
        const int cnt = nseg - first < kMaxSegsPerLaunch ? nseg - first : kMaxSegsPerLaunch;
        SegBatch batch{};
        for (int i = 0; i < cnt; i++) batch.s[i] = segs[first + i];
        cudaError_t e = cudaMemsetAsync(d_matches, 0, sizeof(unsigned long long) * kMaxSegsPerLaunch, stream);
        if (e == cudaSuccess) {
            ltu_scan_filter_kernel<<<dim3(kScanGroups, cnt), 32, 0, stream>>>(batch, d_matches);
            g_est_launches.fetch_add(1, std::memory_order_relaxed);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(matches + first, d_matches, sizeof(uint64_t) * cnt, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
            note_cuda_error(e);
            return Status::kCudaError;
        }
    }
    return Status::kOk;
}

}  // namespace dlt
