// estimator.cu — LTU-semantics LZ match estimator on the GPU (see estimator.h).
//
// The restated CPU algorithm is a sequential scan with a 2^16-entry "last 3-byte key seen in this
// bucket" table, processed 4 positions at a time (4 compares against the table as it was before
// the group, then 4 updates).  Hence: position p is a match iff the most recent earlier position of
// the SAME BUCKET that lies in an EARLIER group of four holds the same key (an untouched bucket
// holds 0, so key 0 matches it).
//
// Buckets never interact, so the sequential table disappears once the positions of every bucket sit
// next to each other in stream order:
//   record(p) = key(p) | nskip(p) << 24,   nskip = same-bucket positions before p inside p's group of 4
//   stable sort of the records by bucket (LSD radix, 2 x 8 bits, hand-written: per-tile shared-memory
//   histograms, stable warp ranks from ballot-built match masks, one exclusive scan per pass)
//   match(i) <=> key(rec[i]) == (bucket(rec[i-1-nskip]) == bucket(rec[i]) ? key(rec[i-1-nskip]) : 0)
// which is embarrassingly parallel and has no data-dependent skew (a flat texture puts every
// position in one bucket; a per-bucket sequential consumer would serialise on it).
//
// Tiny inputs use a single-launch kernel instead (every warp owns a bucket range, scans the whole
// segment and keeps a (last, base) pair per bucket); it also serves as an independent second
// implementation in the tests.
#include "estimator.h"

#include <algorithm>
#include <atomic>

namespace dlt {
namespace {

std::atomic<uint64_t> g_est_launches{0};

constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kRecKeyMask = 0x00FFFFFFu;

// Number of positions the reference loop visits: groups of 4 starting at i = 0,4,.. while i < len-7.
__host__ __device__ inline size_t ltu_positions(size_t len) {
    const size_t end = len > (size_t)kLtuTailGuard ? len - kLtuTailGuard : 0;
    return (end + kLtuGroup - 1) / kLtuGroup * kLtuGroup;
}

__device__ __forceinline__ uint32_t ltu_bucket(uint32_t key) { return (key * kLtuGoldenRatio) >> (32 - kLtuHashBits); }

// Lanes that are valid and hold the same BITS-bit value as the calling lane.  Built from BITS ballots:
// constant cost, whereas MATCH.ANY slows down with the number of distinct values in the warp (measured:
// ~400 cycles per call on random 8-bit digits, which made the first version of the scatter 15x slower).
template <int BITS>
__device__ __forceinline__ unsigned match_bits(uint32_t v, bool valid) {
    unsigned mask = __ballot_sync(kFull, valid);
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        const bool bit = (v >> b) & 1u;
        const unsigned bal = __ballot_sync(kFull, bit);
        mask &= bit ? bal : ~bal;
    }
    return mask;
}

// nskip for the 32 consecutive positions held by a warp (groups of 4 are lane-aligned).
__device__ __forceinline__ uint32_t group_nskip(uint32_t bucket, unsigned lane) {
    uint32_t nskip = 0;
#pragma unroll
    for (int k = 1; k < kLtuGroup; k++) {
        const uint32_t bk = __shfl_up_sync(kFull, bucket, k);
        if ((int)(lane & (kLtuGroup - 1)) >= k && bk == bucket) nskip++;
    }
    return nskip;
}

// =================================================================================================
// Small-input path: one launch, warps own bucket ranges
// =================================================================================================
constexpr int kMaxSegsSmall = 32;
struct SmallBatch {
    LtuSegment s[kMaxSegsSmall];
};

constexpr int kScanGroups = 64;
constexpr int kScanBucketBits = kLtuHashBits - 6;
constexpr int kScanBuckets = 1 << kScanBucketBits;
static_assert(kScanGroups * kScanBuckets == (1 << kLtuHashBits), "bucket groups");

// One warp step over up to 32 positions of this warp's bucket range, in stream order by lane.
__device__ __forceinline__ int consume_step(bool valid, uint32_t b, uint32_t key, int nskip, uint32_t* t_last,
                                            uint32_t* t_base) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned mask = match_bits<kScanBucketBits>(b, valid);
    const unsigned lower = mask & ((1u << lane) - 1u);
    const int r = __popc(lower);
    unsigned m = lower;
#pragma unroll
    for (int i = 0; i < kLtuGroup - 1; i++)
        if (i < nskip && m) m &= ~(0x80000000u >> __clz(m));
    const int src = m ? 31 - __clz(m) : (int)lane;
    const uint32_t from_lane = __shfl_sync(kFull, key, src);
    uint32_t cmp = 0;
    // nskip < r: the predecessor is a lower lane; == r: it is the bucket's `last`; > r: this
    // position's group of four began in an earlier step, compare against what that group saw.
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


__global__ void __launch_bounds__(32) ltu_scan_filter_kernel(const SmallBatch batch, unsigned long long* matches) {
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
        const int nskip = (int)group_nskip(bucket, lane);
        const bool mine = inb && (bucket / kScanBuckets) == g;
        count += consume_step(mine, bucket % kScanBuckets, key, nskip, t_last, t_base);
    }
    if (lane == 0 && count) atomicAdd(&matches[blockIdx.y], count);
}

// =================================================================================================
// Large-input path: LSD radix sort of the records by bucket, then a parallel compare
// =================================================================================================
constexpr int kMaxSegs = 16;
constexpr int kTile = 8192;              // records per CTA tile
constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kPerWarp = kTile / kSortWarps;     // contiguous records per warp (scatter: warp-striped)
constexpr int kSteps = kPerWarp / 32;            // records per thread
constexpr int kRadix = 256;
constexpr int kScanBlockElems = 4096;            // elements per block of the offset scan
static_assert(kSteps == 16 && kTile / kSortThreads == 16, "16 records per thread");

struct SortBatch {
    LtuSegment seg[kMaxSegs];
    uint32_t* rec_a[kMaxSegs];   // records after pass 0
    uint32_t* rec_b[kMaxSegs];   // records after pass 1 (sorted by bucket, stable)
    uint32_t* cnt[kMaxSegs];     // [kRadix][ntiles] tile histograms -> exclusive offsets
    uint32_t* blk[kMaxSegs];     // block sums of the scan
    uint32_t npos[kMaxSegs];
    uint32_t ntiles[kMaxSegs];
};

template <int PASS>
__device__ __forceinline__ uint32_t digit_of(uint32_t rec) {
    const uint32_t b = ltu_bucket(rec & kRecKeyMask);
    return PASS == 0 ? (b & 0xFFu) : (b >> 8);
}

// Stage `need` bytes starting at `src` (any alignment) so that stage[(src & 15) + i] == src[i].
// Interior as 128-bit loads, ragged edges bytewise; never reads outside [src, src + need).
__device__ __forceinline__ int stage_bytes(const uint8_t* src, int need, uint8_t* stage) {
    const int sh = (int)(reinterpret_cast<uintptr_t>(src) & 15);
    const uint8_t* al = src - sh;
    const int nch = (sh + need + 15) >> 4;
    for (int k = threadIdx.x; k < nch; k += blockDim.x) {
        const int lo = k << 4;
        if (lo >= sh && lo + 16 <= sh + need) {
            *reinterpret_cast<uint4*>(stage + lo) = __ldg(reinterpret_cast<const uint4*>(al + lo));
        } else {
            const int a = lo > sh ? lo : sh, b = lo + 16 < sh + need ? lo + 16 : sh + need;
            for (int i = a; i < b; i++) stage[i] = al[i];
        }
    }
    return sh;
}

// ---- tile histograms: 16 CONSECUTIVE records per thread, run-length aggregated shared atomics ------
// (no ordering needed here, so no match_any; a flat texture gives one atomic per thread, not 16)
template <int PASS>
__global__ void __launch_bounds__(kSortThreads) ltu_hist_kernel(const SortBatch b) {
    const int seg = blockIdx.y;
    const uint32_t tile = blockIdx.x;
    if (tile >= b.ntiles[seg]) return;
    __shared__ uint32_t hist[kRadix];
    __shared__ __align__(16) uint8_t stage[PASS == 0 ? kTile + 48 : 16];
    if (threadIdx.x < kRadix) hist[threadIdx.x] = 0;
    const uint32_t left = b.npos[seg] - tile * kTile;
    const int nvalid = left < (uint32_t)kTile ? (int)left : kTile;
    const int first = threadIdx.x * 16;
    uint32_t digit[16];
    if constexpr (PASS == 0) {
        const int sh = stage_bytes(b.seg[seg].d_ptr + (size_t)tile * kTile, nvalid + 2, stage);
        __syncthreads();
        // bytes [first, first + 18) of the tile -> five byte-aligned words -> 16 three-byte keys
        const int a = sh + first;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(stage + (a & ~3));
        uint32_t raw[6], w[5];
#pragma unroll
        for (int k = 0; k < 6; k++) raw[k] = wp[k];
#pragma unroll
        for (int k = 0; k < 5; k++) w[k] = __funnelshift_r(raw[k], raw[k + 1], 8 * (a & 3));
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const uint32_t key = __funnelshift_r(w[j >> 2], w[(j >> 2) + 1], 8 * (j & 3)) & kRecKeyMask;
            digit[j] = ltu_bucket(key) & 0xFFu;
        }
    } else {
        const uint4* src = reinterpret_cast<const uint4*>(b.rec_a[seg] + (size_t)tile * kTile + first);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (first + 4 * q < nvalid) v = __ldg(src + q);  // nvalid is a multiple of 4
            digit[4 * q + 0] = digit_of<1>(v.x);
            digit[4 * q + 1] = digit_of<1>(v.y);
            digit[4 * q + 2] = digit_of<1>(v.z);
            digit[4 * q + 3] = digit_of<1>(v.w);
        }
    }
    uint32_t run = 0, cur = digit[0];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if (first + j < nvalid) {
            if (digit[j] != cur) {
                atomicAdd(&hist[cur], run);
                cur = digit[j], run = 0;
            }
            run++;
        }
    }
    if (run) atomicAdd(&hist[cur], run);
    __syncthreads();
    if (threadIdx.x < kRadix) b.cnt[seg][(size_t)threadIdx.x * b.ntiles[seg] + tile] = hist[threadIdx.x];
}

// ---- exclusive scan of cnt[seg][0 .. kRadix*ntiles) in three small launches ----------------------
__global__ void __launch_bounds__(256) ltu_scan_sums_kernel(const SortBatch b) {
    const int seg = blockIdx.y;
    const uint32_t n = kRadix * b.ntiles[seg];
    const uint32_t base = blockIdx.x * kScanBlockElems;
    if (base >= n) return;
    uint32_t s = 0;
    for (uint32_t i = base + threadIdx.x; i < base + kScanBlockElems && i < n; i += 256) s += b.cnt[seg][i];
    __shared__ uint32_t ws[8];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; i++) t += ws[i];
        b.blk[seg][blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(32) ltu_scan_blocks_kernel(const SortBatch b) {
    const int seg = blockIdx.x;
    const uint32_t n = kRadix * b.ntiles[seg];
    const uint32_t nblk = (n + kScanBlockElems - 1) / kScanBlockElems;
    const unsigned lane = threadIdx.x;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nblk; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t v = i < nblk ? b.blk[seg][i] : 0;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(kFull, inc, o);
            if ((int)lane >= o) inc += up;
        }
        if (i < nblk) b.blk[seg][i] = carry + inc - v;
        carry += __shfl_sync(kFull, inc, 31);
    }
}

__global__ void __launch_bounds__(256) ltu_scan_apply_kernel(const SortBatch b) {
    const int seg = blockIdx.y;
    const uint32_t n = kRadix * b.ntiles[seg];
    const uint32_t base = blockIdx.x * kScanBlockElems;
    if (base >= n) return;
    constexpr int kPer = kScanBlockElems / 256;  // consecutive elements per thread
    uint32_t v[kPer], sum = 0;
    const uint32_t first = base + threadIdx.x * kPer;
#pragma unroll
    for (int k = 0; k < kPer; k++) {
        v[k] = first + k < n ? b.cnt[seg][first + k] : 0;
        sum += v[k];
    }
    // block-wide exclusive scan of the per-thread sums
    __shared__ uint32_t ws[8];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(kFull, inc, o);
        if ((int)lane >= o) inc += up;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    uint32_t off = b.blk[seg][blockIdx.x] + inc - sum;
    for (int w = 0; w < (int)warp; w++) off += ws[w];
#pragma unroll
    for (int k = 0; k < kPer; k++) {
        if (first + k < n) b.cnt[seg][first + k] = off;
        off += v[k];
    }
}

// ---- stable scatter of one tile by the pass's digit -------------------------------------------------
// Warp-striped: warp w owns tile records [w*kPerWarp, (w+1)*kPerWarp), lane l of step t holds record
// w*kPerWarp + t*32 + l, so lane order inside a step is stream order and the match-mask ranks are stable.
template <int PASS>
__global__ void __launch_bounds__(kSortThreads, 2) ltu_scatter_kernel(const SortBatch b) {
    const int seg = blockIdx.y;
    const uint32_t tile = blockIdx.x;
    if (tile >= b.ntiles[seg]) return;
    // the byte staging of pass 0 is dead once the records are in registers: it shares `sorted`
    __shared__ __align__(16) uint32_t sorted[kTile];     // the tile in digit order
    __shared__ uint16_t warp_cnt[kSortWarps][kRadix];    // per-warp digit counts -> per-warp bases
    __shared__ uint32_t bin_start[kRadix], gofs[kRadix];
    __shared__ uint32_t wsum[kRadix / 32];
    uint8_t* stage = reinterpret_cast<uint8_t*>(sorted);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * kRadix / 2; i += kSortThreads)
        reinterpret_cast<uint32_t*>(&warp_cnt[0][0])[i] = 0;
    const uint32_t left = b.npos[seg] - tile * kTile;
    const int nvalid = left < (uint32_t)kTile ? (int)left : kTile;

    uint32_t rec[kSteps];
    if constexpr (PASS == 0) {
        const int sh = stage_bytes(b.seg[seg].d_ptr + (size_t)tile * kTile, nvalid + 2, stage);  // key(p) = bytes p..p+2
        __syncthreads();
#pragma unroll
        for (int t = 0; t < kSteps; t++) {
            const int i = warp * kPerWarp + t * 32 + lane;
            uint32_t key = 0;
            if (i < nvalid) {
                const int a = sh + i;
                const uint32_t w0 = *reinterpret_cast<const uint32_t*>(stage + (a & ~3));
                const uint32_t w1 = *reinterpret_cast<const uint32_t*>(stage + (a & ~3) + 4);
                key = __funnelshift_r(w0, w1, 8 * (a & 3)) & kRecKeyMask;
            }
            rec[t] = key | (group_nskip(ltu_bucket(key), lane) << 24);  // nvalid is a multiple of 4
        }
        __syncthreads();  // everyone is done reading `stage` before `sorted` is written
    } else {
        const uint32_t* src = b.rec_a[seg] + (size_t)tile * kTile;
#pragma unroll
        for (int t = 0; t < kSteps; t++) {
            const int i = warp * kPerWarp + t * 32 + lane;
            rec[t] = i < nvalid ? __ldg(src + i) : 0u;
        }
        __syncthreads();
    }

    // rank of every record among the records of its warp with the same digit (stream order);
    // meta = digit | rank << 8
    uint32_t meta[kSteps];
#pragma unroll
    for (int t = 0; t < kSteps; t++) {
        const int i = warp * kPerWarp + t * 32 + lane;
        const bool valid = i < nvalid;
        const uint32_t d = digit_of<PASS>(rec[t]);
        const unsigned mask = match_bits<8>(d, valid);
        const uint32_t before = valid ? warp_cnt[warp][d] : 0;
        meta[t] = d | ((before + __popc(mask & ((1u << lane) - 1u))) << 8);
        __syncwarp();
        if (valid && (mask >> lane) == 1u) warp_cnt[warp][d] = (uint16_t)(before + __popc(mask));
        __syncwarp();
    }
    __syncthreads();

    // thread d < 256: per-warp bases of digit d, the tile's digit histogram and its exclusive scan
    uint32_t tot = 0, inc = 0;
    if (threadIdx.x < kRadix) {
        const int d = threadIdx.x;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            const uint32_t c = warp_cnt[w][d];
            warp_cnt[w][d] = (uint16_t)tot;
            tot += c;
        }
        inc = tot;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(kFull, inc, o);
            if ((int)lane >= o) inc += up;
        }
        if (lane == 31) wsum[warp] = inc;
    }
    __syncthreads();
    if (threadIdx.x < kRadix) {
        uint32_t off = inc - tot;
        for (int w = 0; w < (int)warp; w++) off += wsum[w];
        bin_start[threadIdx.x] = off;
        gofs[threadIdx.x] = b.cnt[seg][(size_t)threadIdx.x * b.ntiles[seg] + tile];
    }
    __syncthreads();

#pragma unroll
    for (int t = 0; t < kSteps; t++) {
        const int i = warp * kPerWarp + t * 32 + lane;
        if (i < nvalid) {
            const uint32_t d = meta[t] & 0xFFu;
            sorted[bin_start[d] + warp_cnt[warp][d] + (meta[t] >> 8)] = rec[t];
        }
    }
    __syncthreads();

    // write-out: consecutive threads write consecutive records of a digit's run
    uint32_t* out = PASS == 0 ? b.rec_a[seg] : b.rec_b[seg];
    for (int i = threadIdx.x; i < nvalid; i += kSortThreads) {
        const uint32_t r = sorted[i];
        const uint32_t d = digit_of<PASS>(r);
        out[gofs[d] + (uint32_t)i - bin_start[d]] = r;
    }
}

// ---- compare every record with its bucket predecessor --------------------------------------------
// Four consecutive records per thread (one 128-bit load); the predecessor of a record is at most 4
// records back, i.e. in the thread's own vector or in the previous lane's (shuffle; lane 0 reloads).
__global__ void __launch_bounds__(256) ltu_compare_kernel(const SortBatch b, unsigned long long* matches) {
    const int seg = blockIdx.y;
    const uint32_t n = b.npos[seg];  // a multiple of 4
    const uint4* rec4 = reinterpret_cast<const uint4*>(b.rec_b[seg]);
    const uint32_t nvec = n / 4;
    const unsigned lane = threadIdx.x & 31;
    uint32_t count = 0;
    for (uint32_t vbase = blockIdx.x * 256u + (threadIdx.x & ~31u); vbase < nvec; vbase += gridDim.x * 256u) {
        const uint32_t v = vbase + lane;
        const bool in = v < nvec;
        const uint4 cur = in ? __ldg(rec4 + v) : make_uint4(0, 0, 0, 0);
        uint4 prev;
        prev.x = __shfl_up_sync(kFull, cur.x, 1);
        prev.y = __shfl_up_sync(kFull, cur.y, 1);
        prev.z = __shfl_up_sync(kFull, cur.z, 1);
        prev.w = __shfl_up_sync(kFull, cur.w, 1);
        if (lane == 0 && in && v > 0) prev = __ldg(rec4 + v - 1);
        const uint32_t w[8] = {prev.x, prev.y, prev.z, prev.w, cur.x, cur.y, cur.z, cur.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t r = w[4 + k];
            const uint32_t key = r & kRecKeyMask, back = (r >> 24) + 1u;  // 1..4
            const uint32_t q = (back == 1 ? w[3 + k] : back == 2 ? w[2 + k] : back == 3 ? w[1 + k] : w[k]) & kRecKeyMask;
            const bool have = (uint64_t)v * 4 + k >= back;  // a predecessor index exists at all
            uint32_t cmp = 0;                               // an untouched bucket holds 0
            if (have && ltu_bucket(q) == ltu_bucket(key)) cmp = q;
            count += in && key == cmp;
        }
    }
    for (int o = 16; o; o >>= 1) count += __shfl_xor_sync(kFull, count, o);
    __shared__ uint32_t ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = count;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; i++) t += ws[i];
        if (t) atomicAdd(&matches[seg], t);
    }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct SegPlan {
    size_t npos, ntiles, rec_bytes, cnt_bytes, blk_bytes;
};
SegPlan plan_segment(size_t len) {
    SegPlan p{};
    p.npos = ltu_positions(len);
    p.ntiles = (p.npos + kTile - 1) / kTile;
    p.rec_bytes = align_up(p.npos * 4, 256);
    p.cnt_bytes = align_up(p.ntiles * kRadix * 4, 256);
    p.blk_bytes = align_up((p.ntiles * kRadix + kScanBlockElems - 1) / kScanBlockElems * 4, 256);
    return p;
}

constexpr size_t kSmallPositions = 4096;   // at or below: the single-launch kernel
constexpr size_t kResultBytes = 256;       // matches[] at the front of the scratch

}  // namespace

uint64_t estimator_launch_count() { return g_est_launches.load(std::memory_order_relaxed); }

size_t ltu_scratch_bytes(const LtuSegment* segs, int nseg) {
    size_t total = kResultBytes;
    for (int i = 0; i < nseg; i++) {
        const SegPlan p = plan_segment(segs[i].len);
        if (p.npos > kSmallPositions) total += 2 * p.rec_bytes + p.cnt_bytes + p.blk_bytes;
    }
    return total;
}

Status ltu_matches_device(const LtuSegment* segs, int nseg, uint64_t* matches, cudaStream_t stream, uint8_t* scratch,
                          size_t scratch_bytes) {
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
    static_assert(kMaxSegs * sizeof(unsigned long long) <= kResultBytes && kMaxSegsSmall * 8 <= kResultBytes, "");
    if (scratch_bytes < ltu_scratch_bytes(segs, nseg)) return Status::kOutOfMemory;
    unsigned long long* d_matches = reinterpret_cast<unsigned long long*>(scratch);
    auto fail = [](cudaError_t e) {
        note_cuda_error(e);
        return Status::kCudaError;
    };

    // split into the two paths, keeping the caller's order in `matches`
    int small_idx[kMaxSegsSmall], large_idx[kMaxSegs];
    int done = 0;
    while (done < nseg) {
        int ns = 0, nl = 0, i = done;
        for (; i < nseg; i++) {
            const size_t npos = ltu_positions(segs[i].len);
            if (npos > 0xFFFFFFFFull) return Status::kCudaError;  // record indices are 32-bit
            if (npos <= kSmallPositions) {
                if (ns == kMaxSegsSmall) break;
                small_idx[ns++] = i;
            } else {
                if (nl == kMaxSegs) break;
                large_idx[nl++] = i;
            }
        }
        cudaError_t e = cudaMemsetAsync(d_matches, 0, kResultBytes, stream);
        if (e != cudaSuccess) return fail(e);
        uint64_t host[kMaxSegsSmall];

        if (ns) {
            SmallBatch sb{};
            for (int k = 0; k < ns; k++) sb.s[k] = segs[small_idx[k]];
            ltu_scan_filter_kernel<<<dim3(kScanGroups, ns), 32, 0, stream>>>(sb, d_matches);
            g_est_launches.fetch_add(1, std::memory_order_relaxed);
            if ((e = cudaGetLastError()) != cudaSuccess) return fail(e);
            if ((e = cudaMemcpyAsync(host, d_matches, sizeof(uint64_t) * ns, cudaMemcpyDeviceToHost, stream)) != cudaSuccess ||
                (e = cudaStreamSynchronize(stream)) != cudaSuccess)
                return fail(e);
            for (int k = 0; k < ns; k++) matches[small_idx[k]] = host[k];
            if (nl && (e = cudaMemsetAsync(d_matches, 0, kResultBytes, stream)) != cudaSuccess) return fail(e);
        }

        if (nl) {
            SortBatch b{};
            uint8_t* p = scratch + kResultBytes;
            uint32_t max_tiles = 0, max_scan_blocks = 0, max_cmp_blocks = 0;
            for (int k = 0; k < nl; k++) {
                const SegPlan pl = plan_segment(segs[large_idx[k]].len);
                b.seg[k] = segs[large_idx[k]];
                b.npos[k] = (uint32_t)pl.npos;
                b.ntiles[k] = (uint32_t)pl.ntiles;
                b.rec_a[k] = reinterpret_cast<uint32_t*>(p), p += pl.rec_bytes;
                b.rec_b[k] = reinterpret_cast<uint32_t*>(p), p += pl.rec_bytes;
                b.cnt[k] = reinterpret_cast<uint32_t*>(p), p += pl.cnt_bytes;
                b.blk[k] = reinterpret_cast<uint32_t*>(p), p += pl.blk_bytes;
                const uint32_t sblk = (uint32_t)((pl.ntiles * kRadix + kScanBlockElems - 1) / kScanBlockElems);
                const uint32_t cblk = (uint32_t)std::min<size_t>((pl.npos / 4 + 255) / 256, 148 * 8);  // grid-stride, 4 records per thread
                max_tiles = pl.ntiles > max_tiles ? (uint32_t)pl.ntiles : max_tiles;
                max_scan_blocks = sblk > max_scan_blocks ? sblk : max_scan_blocks;
                max_cmp_blocks = cblk > max_cmp_blocks ? cblk : max_cmp_blocks;
            }
            const dim3 tiles(max_tiles, nl), scan_grid(max_scan_blocks, nl);
            auto scan = [&]() {
                ltu_scan_sums_kernel<<<scan_grid, 256, 0, stream>>>(b);
                ltu_scan_blocks_kernel<<<nl, 32, 0, stream>>>(b);
                ltu_scan_apply_kernel<<<scan_grid, 256, 0, stream>>>(b);
            };
            ltu_hist_kernel<0><<<tiles, kSortThreads, 0, stream>>>(b);
            scan();
            ltu_scatter_kernel<0><<<tiles, kSortThreads, 0, stream>>>(b);
            ltu_hist_kernel<1><<<tiles, kSortThreads, 0, stream>>>(b);
            scan();
            ltu_scatter_kernel<1><<<tiles, kSortThreads, 0, stream>>>(b);
            ltu_compare_kernel<<<dim3(max_cmp_blocks, nl), 256, 0, stream>>>(b, d_matches);
            g_est_launches.fetch_add(11, std::memory_order_relaxed);
            if ((e = cudaGetLastError()) != cudaSuccess) return fail(e);
            if ((e = cudaMemcpyAsync(host, d_matches, sizeof(uint64_t) * nl, cudaMemcpyDeviceToHost, stream)) != cudaSuccess ||
                (e = cudaStreamSynchronize(stream)) != cudaSuccess)
                return fail(e);
            for (int k = 0; k < nl; k++) matches[large_idx[k]] = host[k];
        }
        done = i;
    }
    return Status::kOk;
}

}  // namespace dlt
