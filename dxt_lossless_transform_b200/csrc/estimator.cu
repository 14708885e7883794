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
#include <vector>

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
    int slot[kMaxSegsSmall];   // index of the segment's result in matches[]
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
    if (lane == 0 && count) atomicAdd(&matches[batch.slot[blockIdx.y]], count);
}

// =================================================================================================
// Large-input path: ONE stable partition pass by the low kPartBits bucket bits, then per-piece sequential compare
// =================================================================================================
//   1. hist + scan + scatter : stable partition of the records by bucket & (kParts - 1) (hand-written radix pass:
//      per-tile shared-memory histograms, ballot-built stable ranks, tile staged in shared memory so every
//      partition's run leaves as one contiguous write).  Inside a partition the records are in stream order and a
//      bucket is identified by its CLASS = bucket >> kPartBits (kClasses of them).
//   2. plan    : the partitioned array is cut into pieces of <= L records (L-aligned slots, never across a partition
//      boundary); one small kernel numbers them.
//   3. runs    : ONE THREAD owns a piece and walks it sequentially with a private kClasses-entry "last key of this
//      class" table in shared memory (table[class][lane]: conflict-free); record loads are double-buffered 128-byte
//      vectors, the interior is processed four records at a time without branches.  No cross-lane traffic: ~30
//      instructions per record instead of the ~165 of a second radix pass + neighbour compare.  A record whose
//      table entry is still unknown (first touch of the class in the piece) cannot be decided locally: its key is
//      parked in dkey[] and the piece publishes its final table (state[]).
//      kPartBits = 10 keeps the table at 256 bytes per thread, so ~20 warps of piece-threads are resident per SM
//      (the kernel is a chain of dependent shared-memory round trips: it lives off warp-level parallelism).
//   4. resolve : chunks of 32 pieces; per class the parked keys are compared against the last key written by an
//      earlier piece of the same partition (or 0 at the partition start).  The in-state of a chunk is fetched lazily
//      from per-chunk summaries, so a flat texture (everything in one partition, one class) costs one step and
//      nothing is sequential across the whole array.
// Same-group semantics (4 positions compared before the 4 updates): every record carries nskip; a record with
// nskip > 0 ("follower") sees what the first same-bucket record of its group ("leader") saw, which the thread keeps
// in a three-record history.  A follower belongs to the piece that owns its leader, so pieces hand over cleanly: a
// piece skips leading followers of a foreign leader and processes up to three trailing followers of its own.
constexpr int kMaxSegs = 256;              // segments per launch set (the batch descriptor is a ~28 KiB kernel parameter; limit 32,764 B)
constexpr int kTile = 8192;              // records per CTA tile of the partition pass
constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kPerWarp = kTile / kSortWarps;     // contiguous records per warp (scatter: warp-striped)
constexpr int kSteps = kPerWarp / 32;            // records per thread
constexpr int kPartBits = 10;
constexpr int kParts = 1 << kPartBits;           // partitions = radix of the scatter pass
constexpr int kClassBits = kLtuHashBits - kPartBits;
constexpr int kClasses = 1 << kClassBits;        // buckets per partition
constexpr int kColChunk = 64;                    // tiles per chunk of the column scan
constexpr int kChunkPieces = 32;                 // pieces per resolve chunk
constexpr uint32_t kMinRunLen = 128;             // shortest piece (records); always a multiple of 32
constexpr int kRunsWarpsPerSm = 20;              // resident piece-warps per SM (register cap 102)
constexpr uint32_t kTargetPieces = 148 * kRunsWarpsPerSm * 32 * 9 / 10;  // one resident wave of piece-threads, 10 % slack
constexpr uint32_t kUnknown = 0x80000000u;       // table / seen: no record of this class yet in this piece
constexpr uint32_t kForeign = 0x40000000u;       // seen: the record belongs to a neighbouring piece
constexpr uint32_t kParkedMask = 0x07000000u;    // state word bits 24-26: number of parked keys (0..4)
static_assert(kSteps == 16 && kTile / kSortThreads == 16, "16 records per thread");
static_assert(kClasses % 4 == 0 && kClasses >= 32, "class table layout");

struct SortBatch {
    LtuSegment seg[kMaxSegs];
    uint32_t* rec[kMaxSegs];        // records, partitioned by bucket & (kParts - 1) (stable)
    uint32_t* cnt[kMaxSegs];        // [ntiles + 1][kParts] tile histograms -> exclusive offsets (digit-major order);
                                    // row ntiles = where every digit's partition ends
    uint32_t* blk[kMaxSegs];        // [chunks][kParts] column sums of kColChunk tiles -> offset of the chunk's first tile
    uint32_t* part_off[kMaxSegs];   // [kParts + 1]: first record of every partition
    uint32_t* piece_base[kMaxSegs]; // [kParts + 1]: first piece of every partition; [kParts] = number of pieces
    uint16_t* part[kMaxSegs];       // [piece]: partition of the piece
    uint32_t* state[kMaxSegs];      // [piece][class]: last key | parked count << 24 | kUnknown
    uint32_t* dkey[kMaxSegs];       // [piece][4][class]: parked keys
    uint32_t* sum_word[kMaxSegs];   // [chunk][class]: state word of the last piece of the chunk that touched the class
    uint16_t* sum_part[kMaxSegs];   // [chunk][class]: partition of that piece
    uint32_t npos[kMaxSegs];
    uint32_t ntiles[kMaxSegs];
    uint32_t slot[kMaxSegs];        // index of the segment's result in matches[]
    uint32_t run_len;               // L: records per piece slot (multiple of 32)
};

__device__ __forceinline__ uint32_t digit_of_bucket(uint32_t bucket) { return bucket & (kParts - 1); }
__device__ __forceinline__ uint32_t digit_of(uint32_t rec) { return digit_of_bucket(ltu_bucket(rec & kRecKeyMask)); }
__device__ __forceinline__ uint32_t class_of(uint32_t rec) { return ltu_bucket(rec & kRecKeyMask) >> kPartBits; }

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
// (no ordering needed here, so no match masks; a flat texture gives one atomic per thread, not 16)
__global__ void __launch_bounds__(kSortThreads) ltu_hist_kernel(const SortBatch b) {
    const int seg = blockIdx.y;
    const uint32_t tile = blockIdx.x;
    if (tile >= b.ntiles[seg]) return;
    __shared__ uint32_t hist[kParts];
    __shared__ __align__(16) uint8_t stage[kTile + 48];
    for (int d = threadIdx.x; d < kParts; d += kSortThreads) hist[d] = 0;
    const uint32_t left = b.npos[seg] - tile * kTile;
    const int nvalid = left < (uint32_t)kTile ? (int)left : kTile;
    const int first = threadIdx.x * 16;
    uint32_t digit[16];
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
        digit[j] = digit_of_bucket(ltu_bucket(key));
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
    for (int d = threadIdx.x; d < kParts; d += kSortThreads) b.cnt[seg][(size_t)tile * kParts + d] = hist[d];   // coalesced row
}

// ---- exclusive scan of the [tile][digit] matrix in digit-major order (all tiles of digit 0, then digit 1, ...) -------
// Three small launches, every access a coalesced row of kParts counters: column sums per chunk of kColChunk tiles,
// one CTA per segment that turns them into chunk offsets (+ the digit bases), then the running sums inside each chunk.
__global__ void __launch_bounds__(kParts) ltu_colsum_kernel(const SortBatch b) {
    const int seg = blockIdx.y;
    const uint32_t nt = b.ntiles[seg];
    const uint32_t t0 = blockIdx.x * kColChunk;
    if (t0 >= nt) return;
    const uint32_t t1 = min(nt, t0 + kColChunk);
    const uint32_t d = threadIdx.x;
    const uint32_t* m = b.cnt[seg];
    uint32_t s = 0;
#pragma unroll 8
    for (uint32_t t = t0; t < t1; t++) s += __ldg(m + (size_t)t * kParts + d);
    b.blk[seg][(size_t)blockIdx.x * kParts + d] = s;
}

__global__ void __launch_bounds__(kParts) ltu_colbase_kernel(const SortBatch b) {
    const int seg = blockIdx.x;
    const uint32_t nt = b.ntiles[seg];
    const uint32_t nchunks = (nt + kColChunk - 1) / kColChunk;
    const uint32_t d = threadIdx.x;
    uint32_t* blk = b.blk[seg];
    uint32_t total = 0;
    for (uint32_t c = 0; c < nchunks; c += 8) {   // column total (8 independent loads in flight)
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = c + k < nchunks ? blk[(size_t)(c + k) * kParts + d] : 0;
#pragma unroll
        for (int k = 0; k < 8; k++) total += v[k];
    }
    // digit bases: exclusive scan of the column totals over the digits
    __shared__ uint32_t ws[kParts / 32];
    const unsigned lane = d & 31, warp = d >> 5;
    uint32_t inc = total;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(kFull, inc, o);
        if ((int)lane >= o) inc += up;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    uint32_t base = inc - total;
    for (unsigned w = 0; w < warp; w++) base += ws[w];
    uint32_t run = base;   // exclusive running sum down the column of chunk sums, starting at the digit's base
    for (uint32_t c = 0; c < nchunks; c += 8) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = c + k < nchunks ? blk[(size_t)(c + k) * kParts + d] : 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (c + k < nchunks) blk[(size_t)(c + k) * kParts + d] = run;
            run += v[k];
        }
    }
    b.cnt[seg][(size_t)nt * kParts + d] = base + total;   // the extra row: end of the digit's partition
}

__global__ void __launch_bounds__(kParts) ltu_colapply_kernel(const SortBatch b) {
    const int seg = blockIdx.y;
    const uint32_t nt = b.ntiles[seg];
    const uint32_t t0 = blockIdx.x * kColChunk;
    if (t0 >= nt) return;
    const uint32_t t1 = min(nt, t0 + kColChunk);
    const uint32_t d = threadIdx.x;
    uint32_t* m = b.cnt[seg];
    uint32_t run = b.blk[seg][(size_t)blockIdx.x * kParts + d];
    for (uint32_t t = t0; t < t1; t += 8) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = t + k < t1 ? m[(size_t)(t + k) * kParts + d] : 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (t + k < t1) m[(size_t)(t + k) * kParts + d] = run;
            run += v[k];
        }
    }
}

// ---- stable scatter of one tile by bucket & (kParts - 1) ---------------------------------------------
// Tile-local LSD sort in shared memory, two rounds of kSubBits, entirely THREAD-SEQUENTIAL: a thread owns 32
// consecutive records and a private row of 32 counters per round (cnt[bin][thread]); count -> block-wide exclusive
// scan in (bin, thread) order -> the counters become write cursors.  No ballots, no shuffles in the ranking: ~35
// instructions per record for the 10-bit digit instead of ~75 with warp match masks, and flat data costs the same
// as random data.  Records of a digit then leave the tile as one contiguous, coalesced run.
//   smem: sorted[] and cnt[] are both indexed through pad(): 4 words of padding per 32, so that a thread reading its
//   32 consecutive words with 128-bit loads is conflict-free (lane stride 36 words) and cnt[bin][thread] stays
//   conflict-free too.
constexpr int kPThreads = 256;
constexpr int kPPer = kTile / kPThreads;         // consecutive records per thread
constexpr int kSubBits = kPartBits / 2;
constexpr int kSubBins = 1 << kSubBits;
constexpr int kPaddedTile = kTile + kTile / 8;
constexpr int kScatterSmemBytes = 2 * kPaddedTile * 4 + 128;
static_assert(kPPer == 32 && kSubBits * 2 == kPartBits && kSubBins * kPThreads == kTile, "scatter geometry");
static_assert(kParts == 4 * kPThreads, "four digits per thread in the offset phase");
static_assert(kTile + 48 <= kPaddedTile * 4 && 2 * kParts * 4 <= kPaddedTile * 4, "aliases of the counter area");

__device__ __forceinline__ int pad(int i) { return i + ((i >> 5) << 2); }

// While a record sits in the tile sort its spare bits 26-30 carry the 5-bit digit of the current round, so the ranking
// loops need one shift per record instead of a hash.  Padding records of a partial tile are 0x7FFFFFFF: digit 31 in
// both rounds and last in stream order, hence last in the sorted tile.
constexpr uint32_t kRecMask26 = 0x03FFFFFFu;
constexpr uint32_t kPadRecord = 0x7FFFFFFFu;

template <bool FULL>
__device__ __forceinline__ void scatter_tile(const SortBatch& b, const int seg, const uint32_t tile, const int nvalid,
                                             uint8_t* smem) {
    uint32_t* sorted = reinterpret_cast<uint32_t*>(smem);                 // padded, kPaddedTile words
    uint32_t* cnt = sorted + kPaddedTile;                                  // padded [kSubBins][kPThreads]
    uint8_t* stage = reinterpret_cast<uint8_t*>(cnt);                      // byte staging (dead before round A)
    uint32_t* bin_start = cnt;                                             // [kParts]  (after round B)
    uint32_t* gofs = cnt + kParts;                                         // [kParts]
    uint32_t* wsum = sorted + 2 * kPaddedTile;                             // [8] + [8] + [8]
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31, warp = tid >> 5;

    // global offsets of this tile's four digits per thread and the digit counts (difference to the same digits of the
    // next tile; the row after the last tile holds the partition ends): two coalesced 128-bit loads, used at the end
    uint32_t g_off[4], g_cnt[4];
    {
        const uint4 o = __ldg(reinterpret_cast<const uint4*>(b.cnt[seg] + (size_t)tile * kParts) + tid);
        const uint4 n = __ldg(reinterpret_cast<const uint4*>(b.cnt[seg] + (size_t)(tile + 1) * kParts) + tid);
        g_off[0] = o.x, g_off[1] = o.y, g_off[2] = o.z, g_off[3] = o.w;
        g_cnt[0] = n.x - o.x, g_cnt[1] = n.y - o.y, g_cnt[2] = n.z - o.z, g_cnt[3] = n.w - o.w;
    }

    // ---- keys: strided extraction (conflict-free), one word per position into sorted[]
    const int sh = stage_bytes(b.seg[seg].d_ptr + (size_t)tile * kTile, nvalid + 2, stage);  // key(p) = bytes p..p+2
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kPPer; k++) {
        const int i = k * kPThreads + tid;
        uint32_t key = kPadRecord;
        if (FULL || i < nvalid) {
            const int a = sh + i;
            const uint32_t w0 = *reinterpret_cast<const uint32_t*>(stage + (a & ~3));
            const uint32_t w1 = *reinterpret_cast<const uint32_t*>(stage + (a & ~3) + 4);
            key = __funnelshift_r(w0, w1, 8 * (a & 3)) & kRecKeyMask;
        }
        sorted[pad(i)] = key;
    }
    __syncthreads();   // stage is dead from here on

    // ---- records: 32 consecutive positions per thread, nskip inside each group of four, round-A digit in bits 26-30
    uint32_t rec[kPPer];
    {
        const uint4* src = reinterpret_cast<const uint4*>(sorted + pad(tid * kPPer));
#pragma unroll
        for (int q = 0; q < kPPer / 4; q++) {
            const uint4 v = src[q];
            const uint32_t k0 = v.x, k1 = v.y, k2 = v.z, k3 = v.w;
            // nvalid is a multiple of 4: a group is entirely valid or entirely padding
            if (!FULL && k0 == kPadRecord) {
                rec[4 * q + 0] = rec[4 * q + 1] = rec[4 * q + 2] = rec[4 * q + 3] = kPadRecord;
                continue;
            }
            const uint32_t b0 = ltu_bucket(k0), b1 = ltu_bucket(k1), b2 = ltu_bucket(k2), b3 = ltu_bucket(k3);
            const uint32_t n1 = b1 == b0, n2 = (uint32_t)(b2 == b0) + (b2 == b1), n3 = (uint32_t)(b3 == b0) + (b3 == b1) + (b3 == b2);
            rec[4 * q + 0] = k0 | ((b0 & (kSubBins - 1)) << 26);
            rec[4 * q + 1] = k1 | (n1 << 24) | ((b1 & (kSubBins - 1)) << 26);
            rec[4 * q + 2] = k2 | (n2 << 24) | ((b2 & (kSubBins - 1)) << 26);
            rec[4 * q + 3] = k3 | (n3 << 24) | ((b3 & (kSubBins - 1)) << 26);
        }
    }

    // ---- two LSD rounds.  cnt[bin][thread] lives at word bin*288 + row (pad() folded into the constants).
    const int row = tid + ((tid >> 5) << 2);
#pragma unroll
    for (int round = 0; round < 2; round++) {
        // zero the counters (the padding words too; they are never read)
        for (int i = tid; i < kPaddedTile / 4; i += kPThreads) reinterpret_cast<uint4*>(cnt)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();   // also: every thread holds its records in registers before sorted[] is overwritten
        // count, four records at a time: the four counters are loaded together (no smem round trip between records);
        // a record whose digit repeats an earlier one of the batch adds to that record's value, and the stores go out
        // in order so the last one per counter wins
#pragma unroll
        for (int j = 0; j < kPPer; j += 4) {
            uint32_t d[4], c[4];
            uint32_t* ptr[4];
#pragma unroll
            for (int i = 0; i < 4; i++) d[i] = rec[j + i] >> 26, ptr[i] = cnt + d[i] * 288 + row;
#pragma unroll
            for (int i = 0; i < 4; i++) c[i] = *ptr[i];
            c[1] += d[1] == d[0];
            c[2] += (uint32_t)(d[2] == d[0]) + (d[2] == d[1]);
            c[3] += (uint32_t)(d[3] == d[0]) + (d[3] == d[1]) + (d[3] == d[2]);
#pragma unroll
            for (int i = 0; i < 4; i++) *ptr[i] = c[i] + 1;
        }
        __syncthreads();
        // exclusive scan in (bin, thread) order: thread t owns the 32 consecutive counters [32t, 32t+32)
        {
            uint4* rowp = reinterpret_cast<uint4*>(cnt + pad(tid * 32));
            uint32_t run = 0;   // pass 1: the row total (the row is re-read in pass 2: cheaper than 32 live registers)
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint4 v = rowp[q];
                run += v.x + v.y + v.z + v.w;
            }
            uint32_t inc = run;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(kFull, inc, o);
                if ((int)lane >= o) inc += up;
            }
            if (lane == 31) wsum[round * 8 + warp] = inc;
            __syncthreads();
            uint32_t base = inc - run;
            for (unsigned w = 0; w < warp; w++) base += wsum[round * 8 + w];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint4 v = rowp[q];
                uint4 o;
                o.x = base, base += v.x;
                o.y = base, base += v.y;
                o.z = base, base += v.z;
                o.w = base, base += v.w;
                rowp[q] = o;
            }
        }
        __syncthreads();
        // the counters are cursors now: stable because a thread walks its records in stream order
#pragma unroll
        for (int j = 0; j < kPPer; j += 4) {
            uint32_t d[4], c[4];
            uint32_t* ptr[4];
#pragma unroll
            for (int i = 0; i < 4; i++) d[i] = rec[j + i] >> 26, ptr[i] = cnt + d[i] * 288 + row;
#pragma unroll
            for (int i = 0; i < 4; i++) c[i] = *ptr[i];
            c[1] += d[1] == d[0];
            c[2] += (uint32_t)(d[2] == d[0]) + (d[2] == d[1]);
            c[3] += (uint32_t)(d[3] == d[0]) + (d[3] == d[1]) + (d[3] == d[2]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                *ptr[i] = c[i] + 1;
                sorted[pad((int)c[i])] = rec[j + i];
            }
        }
        __syncthreads();
        if (round == 0) {
            // reload in round-A order and switch the spare bits to the round-B digit
            const uint4* src = reinterpret_cast<const uint4*>(sorted + pad(tid * kPPer));
#pragma unroll
            for (int q = 0; q < kPPer / 4; q++) {
                const uint4 v = src[q];
                const uint32_t r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t hi = (ltu_bucket(r[i] & kRecKeyMask) >> kSubBits) & (kSubBins - 1);
                    rec[4 * q + i] = (!FULL && r[i] == kPadRecord) ? kPadRecord : (r[i] & kRecMask26) | (hi << 26);
                }
            }
        }
    }

    // ---- where each digit's run starts inside the tile (exclusive scan of the four counts per thread) and globally
    {
        const uint32_t sum = g_cnt[0] + g_cnt[1] + g_cnt[2] + g_cnt[3];
        uint32_t inc = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(kFull, inc, o);
            if ((int)lane >= o) inc += up;
        }
        if (lane == 31) wsum[16 + warp] = inc;   // cnt[] (aliased by bin_start / gofs) is dead: all cursors were consumed
        __syncthreads();
        uint32_t off = inc - sum;
        for (unsigned w = 0; w < warp; w++) off += wsum[16 + w];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            bin_start[tid * 4 + k] = off;
            gofs[tid * 4 + k] = g_off[k];
            off += g_cnt[k];
        }
    }
    __syncthreads();

    // ---- write-out: consecutive threads write consecutive records of a digit's run
    uint32_t* out = b.rec[seg];
#pragma unroll 4
    for (int k = 0; k < kPPer; k++) {
        const int i = k * kPThreads + tid;
        if (FULL || i < nvalid) {
            const uint32_t r = sorted[pad(i)] & kRecMask26;
            const uint32_t d = digit_of(r);
            out[gofs[d] + (uint32_t)i - bin_start[d]] = r;
        }
    }
}

#ifndef DLT_SCATTER_CTAS
#define DLT_SCATTER_CTAS 2
#endif
__global__ void __launch_bounds__(kPThreads, DLT_SCATTER_CTAS) ltu_scatter_kernel(const SortBatch b) {
    const int seg = blockIdx.y;
    const uint32_t tile = blockIdx.x;
    if (tile >= b.ntiles[seg]) return;
    extern __shared__ __align__(16) uint8_t scatter_smem[];
    const uint32_t left = b.npos[seg] - tile * kTile;
    if (left >= (uint32_t)kTile) scatter_tile<true>(b, seg, tile, kTile, scatter_smem);
    else scatter_tile<false>(b, seg, tile, (int)left, scatter_smem);
}

// ---- plan: partition boundaries -> pieces ------------------------------------------------------------
// Partition p holds records [part_off[p], part_off[p+1]); it is cut at multiples of L, so it owns the slots
// part_off[p] / L .. (part_off[p+1] - 1) / L.  piece_base[] numbers the pieces partition after partition.
__device__ __forceinline__ void plan_pieces(const SortBatch& b, const int seg, const uint32_t o0, const uint32_t o1) {
    const uint32_t p = threadIdx.x;
    const uint32_t npos = b.npos[seg], L = b.run_len;
    const uint32_t pieces = o1 > o0 ? (o1 - 1) / L - o0 / L + 1 : 0;
    __shared__ uint32_t ws[kParts / 32];
    const unsigned lane = p & 31, warp = p >> 5;
    uint32_t inc = pieces;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(kFull, inc, o);
        if ((int)lane >= o) inc += up;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    uint32_t base = inc - pieces;
    for (unsigned w = 0; w < warp; w++) base += ws[w];
    b.part_off[seg][p] = o0;
    b.piece_base[seg][p] = base;
    if (p == kParts - 1) b.part_off[seg][kParts] = npos, b.piece_base[seg][kParts] = base + pieces;
    for (uint32_t k = 0; k < pieces; k++) b.part[seg][base + k] = (uint16_t)p;
}

__global__ void __launch_bounds__(kParts) ltu_plan_kernel(const SortBatch b) {
    const int seg = blockIdx.x;
    const uint32_t p = threadIdx.x;
    const uint32_t o0 = b.cnt[seg][p];   // row 0 of the scanned matrix: first record of every partition
    const uint32_t o1 = p == kParts - 1 ? b.npos[seg] : b.cnt[seg][p + 1];
    plan_pieces(b, seg, o0, o1);
}

// ---- few tiles (every segment of the launch set has <= kColChunk tiles): the three column-scan launches and the plan
// in ONE launch, one CTA per segment — what a typical texture (<= 512 KiB per endpoint stream) goes through.
__global__ void __launch_bounds__(kParts) ltu_colscan_small_kernel(const SortBatch b) {
    const int seg = blockIdx.x;
    const uint32_t nt = b.ntiles[seg];
    const uint32_t d = threadIdx.x;
    uint32_t* m = b.cnt[seg];
    uint32_t total = 0;
    for (uint32_t t = 0; t < nt; t += 8) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = t + k < nt ? m[(size_t)(t + k) * kParts + d] : 0;
#pragma unroll
        for (int k = 0; k < 8; k++) total += v[k];
    }
    __shared__ uint32_t ws[kParts / 32];
    __shared__ uint32_t s_base[kParts + 1];
    const unsigned lane = d & 31, warp = d >> 5;
    uint32_t inc = total;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(kFull, inc, o);
        if ((int)lane >= o) inc += up;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    uint32_t base = inc - total;
    for (unsigned w = 0; w < warp; w++) base += ws[w];
    s_base[d] = base;
    if (d == kParts - 1) s_base[kParts] = base + total;
    uint32_t run = base;
    for (uint32_t t = 0; t < nt; t += 8) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = t + k < nt ? m[(size_t)(t + k) * kParts + d] : 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (t + k < nt) m[(size_t)(t + k) * kParts + d] = run;
            run += v[k];
        }
    }
    m[(size_t)nt * kParts + d] = base + total;   // the extra row: end of the digit's partition
    __syncthreads();   // s_base complete; ws is reused by plan_pieces after its own barrier
    plan_pieces(b, seg, s_base[d], s_base[d + 1]);
}

// ---- runs: one thread per piece, private class table in shared memory ---------------------------------

__device__ __forceinline__ void ldg_nc32(const void* p, uint4& a, uint4& b) {   // one 256-bit load: a whole sector per lane
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}

// Every piece-thread streams its own piece: tens of thousands of concurrent 128-byte reads at unrelated addresses
// (DRAM row misses).  Each thread therefore pulls the next kPrefetchRecords of its piece into L2 with ONE bulk
// prefetch, so DRAM sees kilobyte-sized contiguous reads and the record loads hit L2.
#ifndef DLT_PREFETCH_RECORDS
#define DLT_PREFETCH_RECORDS 512
#endif
constexpr uint32_t kPrefetchRecords = DLT_PREFETCH_RECORDS;   // 512 records = 2 KiB
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(32, kRunsWarpsPerSm) ltu_runs_kernel(const SortBatch b, unsigned long long* matches) {
    __shared__ uint32_t table[kClasses * 32];      // [class][lane]
    const int seg = blockIdx.y;
    const unsigned lane = threadIdx.x;
    const uint32_t L = b.run_len;
    const uint32_t total = b.piece_base[seg][kParts];
    if (blockIdx.x * 32u >= total) return;

    const uint32_t g = blockIdx.x * 32u + lane;   // piece id
    uint32_t start = 0, end = 0, part_start = 0, part_end = 0;
    if (g < total) {
        const uint32_t p = b.part[seg][g];
        part_start = b.part_off[seg][p], part_end = b.part_off[seg][p + 1];
        const uint32_t slot = part_start / L + (g - b.piece_base[seg][p]);
        start = max(part_start, slot * L);
        end = min(part_end, (slot + 1) * L);
    }
    // A piece that starts its partition knows its in-state: every bucket still holds 0 (the reference's table is
    // zero-initialised), so nothing is parked for the resolve pass.  With many small segments (a directory of
    // textures) every piece is such a piece.
    const uint32_t initial = (g < total && start == part_start) ? 0u : kUnknown;
#pragma unroll 8
    for (int c = 0; c < kClasses; c++) table[c * 32 + lane] = initial;

    const uint32_t* rec = b.rec[seg];
    uint32_t* dk = b.dkey[seg] + (size_t)g * 4 * kClasses;
    uint32_t count = 0;
    // history of the last three records: class and what the record saw (a key, kUnknown or kForeign)
    uint32_t hc1 = 0xFFFFFFFFu, hc2 = 0xFFFFFFFFu, hc3 = 0xFFFFFFFFu, hs1 = kForeign, hs2 = kForeign, hs3 = kForeign;
    auto push = [&](uint32_t cls, uint32_t seen) {
        hc3 = hc2, hs3 = hs2, hc2 = hc1, hs2 = hs1, hc1 = cls, hs1 = seen;
    };
    // General (scalar) step: piece edges, where records may belong to the neighbouring piece.
    auto process = [&](uint32_t r, bool in_range) {
        const uint32_t key = r & kRecKeyMask, nskip = r >> 24;
        const uint32_t cls = ltu_bucket(key) >> kPartBits;
        uint32_t* slot = table + cls * 32 + lane;
        const uint32_t w = *slot;
        uint32_t seen;
        if (nskip) seen = hc1 == cls ? hs1 : hc2 == cls ? hs2 : hs3;   // follower: what its leader saw
        else seen = in_range ? (w & (kUnknown | kRecKeyMask)) : kForeign;  // leader past `end`: the next piece's
        if (!(seen & kForeign)) {
            uint32_t nw = (w & kParkedMask) | key;
            if (seen & kUnknown) {
                dk[((w >> 24) & 7u) * kClasses + cls] = key;   // decided by the resolve pass
                nw += 0x01000000u;
            } else {
                count += seen == key;
            }
            *slot = nw;
        }
        push(cls, seen);
    };
    // Interior step, four records at a time, branch-free: no record can be foreign once three records of the piece
    // have been processed.  The four table entries are loaded together; a record that hits the class of one of the
    // (up to three) records before it in the batch takes that record's updated word instead of the loaded one.
    auto process4 = [&](const uint4 v) {
        const uint32_t r[4] = {v.x, v.y, v.z, v.w};
        uint32_t key[4], cls[4], w[4], nw[4], seen[4];
        uint32_t* slot[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            key[i] = r[i] & kRecKeyMask;
            cls[i] = ltu_bucket(key[i]) >> kPartBits;
            slot[i] = table + cls[i] * 32 + lane;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = *slot[i];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            // the three records before record i: in the batch (distance <= i) or in the history
            const uint32_t c1 = i >= 1 ? cls[i - 1] : hc1, c2 = i >= 2 ? cls[i - 2] : (i == 1 ? hc1 : hc2);
            const uint32_t c3 = i >= 3 ? cls[0] : (i == 2 ? hc1 : i == 1 ? hc2 : hc3);
            const uint32_t s1 = i >= 1 ? seen[i - 1] : hs1, s2 = i >= 2 ? seen[i - 2] : (i == 1 ? hs1 : hs2);
            const uint32_t s3 = i >= 3 ? seen[0] : (i == 2 ? hs1 : i == 1 ? hs2 : hs3);
            const bool m1 = c1 == cls[i], m2 = c2 == cls[i], m3 = c3 == cls[i];
            uint32_t W = w[i];
            if (i >= 3 && m3) W = nw[i - 3];
            if (i >= 2 && m2) W = nw[i - 2];
            if (i >= 1 && m1) W = nw[i - 1];
            const uint32_t follower_seen = m1 ? s1 : m2 ? s2 : s3;
            seen[i] = (r[i] >> 24) ? follower_seen : (W & (kUnknown | kRecKeyMask));
            const uint32_t unknown = seen[i] >> 31;
            count += seen[i] == key[i];                  // an unknown `seen` has bit 31 set: never equal to a key
            nw[i] = ((W & kParkedMask) | key[i]) + (unknown << 24);
            if (unknown) dk[((W >> 24) & 7u) * kClasses + cls[i]] = key[i];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) *slot[i] = nw[i];
        hc3 = cls[1], hs3 = seen[1], hc2 = cls[2], hs2 = seen[2], hc1 = cls[3], hs1 = seen[3];
    };

    if (g < total) {
        // the (up to three) records before the piece are foreign history
        for (uint32_t j = start >= part_start + 3 ? start - 3 : part_start; j < start; j++) push(class_of(__ldg(rec + j)), kForeign);
        uint32_t i = start;
        for (; i < end && ((i & 31u) || i < start + 3); i++) process(__ldg(rec + i), true);
        if (i + 32 <= end) {
            // prime the first prefetch window; the loop keeps one window ahead
            const uint32_t w0 = (i + kPrefetchRecords - 1) / kPrefetchRecords * kPrefetchRecords;   // first window boundary
            if (w0 + 4 <= end) prefetch_l2_bulk(rec + w0, (min(kPrefetchRecords, end - w0) & ~3u) * 4u);   // multiple of 16 bytes
            // two 16-record buffers (two 256-bit loads each): while one is processed the other one's loads are in flight
            uint4 bufa[4], bufb[4];
            ldg_nc32(rec + i, bufa[0], bufa[1]);
            ldg_nc32(rec + i + 8, bufa[2], bufa[3]);
            for (; i + 32 <= end; i += 32) {
                if ((i & (kPrefetchRecords - 1)) == 0) {
                    const uint32_t ahead = i + kPrefetchRecords;   // the window after the one being read
                    if (ahead + 4 <= end) prefetch_l2_bulk(rec + ahead, (min(kPrefetchRecords, end - ahead) & ~3u) * 4u);
                }
                ldg_nc32(rec + i + 16, bufb[0], bufb[1]);
                ldg_nc32(rec + i + 24, bufb[2], bufb[3]);
                asm volatile("" ::: "memory");   // keep the loads ahead of the table traffic below
#pragma unroll
                for (int k = 0; k < 4; k++) process4(bufa[k]);
                if (i + 64 <= end) {
                    ldg_nc32(rec + i + 32, bufa[0], bufa[1]);
                    ldg_nc32(rec + i + 40, bufa[2], bufa[3]);
                }
                asm volatile("" ::: "memory");
#pragma unroll
                for (int k = 0; k < 4; k++) process4(bufb[k]);
            }
        }
        for (; i < end; i++) process(__ldg(rec + i), true);
        for (; i < part_end && i < end + 3; i++) process(__ldg(rec + i), false);   // trailing followers of my leaders
        // publish the final table
        uint32_t* st = b.state[seg] + (size_t)g * kClasses;
#pragma unroll 4
        for (int c = 0; c < kClasses; c += 4) {
            const uint4 v = make_uint4(table[(c + 0) * 32 + lane], table[(c + 1) * 32 + lane], table[(c + 2) * 32 + lane],
                                       table[(c + 3) * 32 + lane]);
            *reinterpret_cast<uint4*>(st + c) = v;
        }
    }
    for (int o = 16; o; o >>= 1) count += __shfl_xor_sync(kFull, count, o);
    if (lane == 0 && count) atomicAdd(&matches[b.slot[seg]], (unsigned long long)count);
}

// ---- resolve, level 1: per chunk and class, the last piece of the chunk that touched the class -----------
constexpr int kResolveThreads = kClasses;   // one thread per class

__global__ void __launch_bounds__(kResolveThreads) ltu_summarize_kernel(const SortBatch b) {
    const int seg = blockIdx.y;
    const uint32_t total = b.piece_base[seg][kParts];
    const uint32_t g0 = blockIdx.x * kChunkPieces;
    if (g0 >= total) return;
    const uint32_t g1 = min(total, g0 + kChunkPieces);
    const uint32_t c = threadIdx.x;
    const uint32_t* st = b.state[seg];
    uint32_t last = kUnknown, last_part = 0;
    for (uint32_t g = g0; g < g1; g += 8) {
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = g + k < g1 ? __ldg(st + (size_t)(g + k) * kClasses + c) : kUnknown;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (!(w[k] & kUnknown)) last = w[k], last_part = b.part[seg][g + k];
    }
    b.sum_word[seg][(size_t)blockIdx.x * kClasses + c] = last;
    b.sum_part[seg][(size_t)blockIdx.x * kClasses + c] = (uint16_t)last_part;
}

// ---- resolve, level 2: walk the pieces of a chunk in order, decide the parked compares -----------------
__global__ void __launch_bounds__(kResolveThreads) ltu_resolve_kernel(const SortBatch b, unsigned long long* matches) {
    const int seg = blockIdx.y;
    const uint32_t total = b.piece_base[seg][kParts];
    const uint32_t g0 = blockIdx.x * kChunkPieces;
    if (g0 >= total) return;
    const uint32_t g1 = min(total, g0 + kChunkPieces);
    const uint32_t c = threadIdx.x;
    const uint32_t* st = b.state[seg];
    const uint32_t* dk = b.dkey[seg];
    const uint16_t* part = b.part[seg];
    __shared__ uint16_t s_part[kChunkPieces + 1];   // s_part[0] = partition of piece g0 - 1
    if (threadIdx.x <= g1 - g0) {
        const uint32_t g = g0 + threadIdx.x;
        s_part[threadIdx.x] = g == 0 ? 0 : part[g - 1];
    }
    __syncthreads();

    uint32_t cur = 0, count = 0;
    bool known = g0 == 0;   // the very first piece starts a partition: every bucket holds 0
    for (uint32_t gb = g0; gb < g1; gb += 8) {
        uint32_t w[8], d0[8];   // state word and first parked key of 8 pieces, all loads in flight together
#pragma unroll
        for (int k = 0; k < 8; k++) {
            w[k] = gb + k < g1 ? __ldg(st + (size_t)(gb + k) * kClasses + c) : kUnknown;
            d0[k] = gb + k < g1 ? __ldg(dk + (size_t)(gb + k) * 4 * kClasses + c) : 0u;   // garbage unless parked >= 1
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t g = gb + k;
            if (g >= g1) break;
            const uint32_t p = s_part[g - g0 + 1];
            if (g != 0 && s_part[g - g0] != p) cur = 0, known = true;   // first piece of its partition
            const uint32_t parked = (w[k] >> 24) & 7u;
            if (parked) {
                if (!known) {
                    // lazily fetch the in-state: the last earlier piece of partition p that touched class c
                    cur = 0;
                    for (int q = (int)blockIdx.x - 1; q >= 0; q--) {
                        const uint32_t sw = __ldg(b.sum_word[seg] + (size_t)q * kClasses + c);
                        if (!(sw & kUnknown)) {
                            if (b.sum_part[seg][(size_t)q * kClasses + c] == p) cur = sw & kRecKeyMask;
                            break;
                        }
                        if (part[(size_t)(q + 1) * kChunkPieces - 1] != p) break;   // chunk q ends before partition p starts
                    }
                    known = true;
                }
                count += d0[k] == cur;
                for (uint32_t i = 1; i < parked; i++) count += __ldg(dk + ((size_t)g * 4 + i) * kClasses + c) == cur;
            }
            if (!(w[k] & kUnknown)) cur = w[k] & kRecKeyMask, known = true;
        }
    }
    for (int o = 16; o; o >>= 1) count += __shfl_xor_sync(kFull, count, o);
    __shared__ uint32_t ws[kResolveThreads / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = count;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < kResolveThreads / 32; i++) t += ws[i];
        if (t) atomicAdd(&matches[b.slot[seg]], t);
    }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct SegPlan {
    size_t npos, ntiles, max_pieces, max_chunks;
    size_t rec_bytes, cnt_bytes, blk_bytes, poff_bytes, pbase_bytes, part_bytes, state_bytes, dkey_bytes, sumw_bytes, sump_bytes;
    size_t bytes() const {
        return rec_bytes + cnt_bytes + blk_bytes + poff_bytes + pbase_bytes + part_bytes + state_bytes + dkey_bytes + sumw_bytes +
               sump_bytes;
    }
};
SegPlan plan_segment(size_t len, uint32_t run_len) {
    SegPlan p{};
    p.npos = ltu_positions(len);
    p.ntiles = (p.npos + kTile - 1) / kTile;
    p.max_pieces = p.npos / run_len + 1 + kParts;   // one partial piece at either end of every partition
    p.max_chunks = (p.max_pieces + kChunkPieces - 1) / kChunkPieces;
    p.rec_bytes = align_up(p.npos * 4, 256);
    p.cnt_bytes = align_up((p.ntiles + 1) * kParts * 4, 256);
    p.blk_bytes = align_up((p.ntiles + kColChunk - 1) / kColChunk * kParts * 4, 256);
    p.poff_bytes = p.pbase_bytes = align_up((kParts + 1) * 4, 256);
    p.part_bytes = align_up(p.max_pieces * 2, 256);
    p.state_bytes = align_up(p.max_pieces * kClasses * 4, 256);
    p.dkey_bytes = align_up(p.max_pieces * 4 * kClasses * 4, 256);
    p.sumw_bytes = align_up(p.max_chunks * kClasses * 4, 256);
    p.sump_bytes = align_up(p.max_chunks * kClasses * 2, 256);
    return p;
}

constexpr size_t kSmallPositions = 4096;   // at or below: the single-launch kernel

// Piece length for ONE launch set (up to kMaxSegs large segments share a set of launches): one resident wave of
// piece-threads over the set (a second, nearly empty wave would double the time), never shorter than kMinRunLen
// (per-piece overhead: table init and publish, up to 4 parked keys per class) and never longer than four average
// partitions: with many small segments the partitions alone fill the wave, and an unbounded piece length would make
// the heaviest partition of a skewed stream (flat texture regions put most positions into one bucket) ONE thread's
// sequential job.
uint32_t choose_run_len(const size_t* lens, int n) {
    size_t total = 0;
    for (int i = 0; i < n; i++) total += ltu_positions(lens[i]);
    const size_t parts = (size_t)n * (kParts + 1);
    const size_t budget = kTargetPieces > parts + 4096 ? kTargetPieces - parts : 4096;
    size_t len = (total + budget - 1) / budget;
    const size_t cap = n ? 4 * (total / ((size_t)n * kParts) + 1) : 0;
    if (len > cap) len = cap;
    len = (len + 31) / 32 * 32;
    return (uint32_t)(len < kMinRunLen ? kMinRunLen : len);
}
size_t set_bytes(const size_t* lens, int n) {
    const uint32_t run_len = choose_run_len(lens, n);
    size_t total = 0;
    for (int i = 0; i < n; i++) total += plan_segment(lens[i], run_len).bytes();
    return total;
}

// The launch sets of a call: large segments in input order, kMaxSegs at a time, each set with its own piece length.
struct LaunchSet {
    std::vector<int> idx;
    uint32_t run_len;
};
std::vector<LaunchSet> plan_sets(const LtuSegment* segs, int nseg) {
    std::vector<LaunchSet> sets;
    for (int i = 0; i < nseg; i++) {
        if (ltu_positions(segs[i].len) <= kSmallPositions) continue;
        if (sets.empty() || (int)sets.back().idx.size() == kMaxSegs) sets.emplace_back();
        sets.back().idx.push_back(i);
    }
    for (LaunchSet& st : sets) {
        size_t lens[kMaxSegs];
        for (size_t i = 0; i < st.idx.size(); i++) lens[i] = segs[st.idx[i]].len;
        st.run_len = choose_run_len(lens, (int)st.idx.size());
    }
    return sets;
}

}  // namespace

uint64_t estimator_launch_count() { return g_est_launches.load(std::memory_order_relaxed); }

void LtuScratchMeter::add(size_t len) {
    static_assert(kSet == kMaxSegs, "LtuScratchMeter mirrors the launch-set size");
    nseg_++;
    if (ltu_positions(len) <= kSmallPositions) return;
    open_len_[open_++] = len;
    if (open_ == kSet) closed_bytes_ += set_bytes(open_len_, open_), open_ = 0;
}
size_t LtuScratchMeter::bytes() const {
    return align_up((nseg_ > 0 ? nseg_ : 1) * sizeof(uint64_t), 256) + closed_bytes_ + (open_ ? set_bytes(open_len_, open_) : 0);
}

inline size_t result_bytes(int nseg) { return align_up((size_t)(nseg > 0 ? nseg : 1) * sizeof(uint64_t), 256); }

size_t ltu_scratch_bytes(const LtuSegment* segs, int nseg) {
    size_t total = result_bytes(nseg);
    for (const LaunchSet& st : plan_sets(segs, nseg))
        for (int i : st.idx) total += plan_segment(segs[i].len, st.run_len).bytes();
    return total;
}

// All segments are queued on `stream` without intermediate host waits: groups of up to kMaxSegsSmall (single-launch
// kernel) / kMaxSegs (partition + runs pipeline) segments share one set of launches (grid.y), every segment has its
// own slice of the scratch and its own result slot, and there is one copy back + one synchronize at the end.  A
// directory of small textures therefore costs a handful of launches, not a handful per texture.
Status ltu_matches_device(const LtuSegment* segs, int nseg, uint64_t* matches, cudaStream_t stream, uint8_t* scratch,
                          size_t scratch_bytes) {
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
    if (nseg <= 0) return Status::kOk;
    if (scratch_bytes < ltu_scratch_bytes(segs, nseg)) return Status::kOutOfMemory;
    unsigned long long* d_matches = reinterpret_cast<unsigned long long*>(scratch);
    auto fail = [](cudaError_t e) {
        note_cuda_error(e);
        return Status::kCudaError;
    };
    for (int i = 0; i < nseg; i++)
        if (ltu_positions(segs[i].len) > 0xFFFF0000ull) return Status::kCudaError;  // record indices are 32-bit

    cudaError_t e = cudaMemsetAsync(d_matches, 0, result_bytes(nseg), stream);
    if (e != cudaSuccess) return fail(e);
    // function attributes are per device: set it on every call (microseconds), a process may drive several GPUs
    if ((e = cudaFuncSetAttribute(ltu_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmemBytes)) != cudaSuccess)
        return fail(e);

    uint8_t* p = scratch + result_bytes(nseg);
    auto take = [&p](size_t bytes) {
        uint8_t* r = p;
        p += bytes;
        return r;
    };

    // ---- small segments: one launch per group
    {
        SmallBatch sb{};
        int ns = 0;
        auto flush = [&]() {
            if (!ns) return cudaSuccess;
            ltu_scan_filter_kernel<<<dim3(kScanGroups, ns), 32, 0, stream>>>(sb, d_matches);
            g_est_launches.fetch_add(1, std::memory_order_relaxed);
            ns = 0;
            return cudaGetLastError();
        };
        for (int i = 0; i < nseg; i++) {
            if (ltu_positions(segs[i].len) > kSmallPositions) continue;
            sb.s[ns] = segs[i];
            sb.slot[ns] = i;
            if (++ns == kMaxSegsSmall && (e = flush()) != cudaSuccess) return fail(e);
        }
        if ((e = flush()) != cudaSuccess) return fail(e);
    }

    // ---- large segments: nine launches per group
    {
        SortBatch b{};
        int nl = 0;
        uint32_t max_tiles = 0, max_scan_blocks = 0, max_piece_warps = 0, max_chunks = 0;
        auto flush = [&]() {
            if (!nl) return cudaSuccess;
            const dim3 tiles(max_tiles, nl), scan_grid(max_scan_blocks, nl);
            ltu_hist_kernel<<<tiles, kSortThreads, 0, stream>>>(b);
            const bool few_tiles = max_tiles <= (uint32_t)kColChunk;
            if (few_tiles) {
                ltu_colscan_small_kernel<<<nl, kParts, 0, stream>>>(b);   // column scans + plan in one launch
            } else {
                ltu_colsum_kernel<<<scan_grid, kParts, 0, stream>>>(b);
                ltu_colbase_kernel<<<nl, kParts, 0, stream>>>(b);
                ltu_colapply_kernel<<<scan_grid, kParts, 0, stream>>>(b);
            }
            ltu_scatter_kernel<<<tiles, kPThreads, kScatterSmemBytes, stream>>>(b);
            if (!few_tiles) ltu_plan_kernel<<<nl, kParts, 0, stream>>>(b);
            ltu_runs_kernel<<<dim3(max_piece_warps, nl), 32, 0, stream>>>(b, d_matches);
            ltu_summarize_kernel<<<dim3(max_chunks, nl), kResolveThreads, 0, stream>>>(b);
            ltu_resolve_kernel<<<dim3(max_chunks, nl), kResolveThreads, 0, stream>>>(b, d_matches);
            g_est_launches.fetch_add(few_tiles ? 6 : 9, std::memory_order_relaxed);
            nl = 0, max_tiles = max_scan_blocks = max_piece_warps = max_chunks = 0;
            return cudaGetLastError();
        };
        for (const LaunchSet& st : plan_sets(segs, nseg)) {
            b.run_len = st.run_len;
            for (int i : st.idx) {
                const SegPlan pl = plan_segment(segs[i].len, st.run_len);
                const int k = nl++;
                b.seg[k] = segs[i];
                b.slot[k] = (uint32_t)i;
                b.npos[k] = (uint32_t)pl.npos;
                b.ntiles[k] = (uint32_t)pl.ntiles;
                b.rec[k] = reinterpret_cast<uint32_t*>(take(pl.rec_bytes));
                b.cnt[k] = reinterpret_cast<uint32_t*>(take(pl.cnt_bytes));
                b.blk[k] = reinterpret_cast<uint32_t*>(take(pl.blk_bytes));
                b.part_off[k] = reinterpret_cast<uint32_t*>(take(pl.poff_bytes));
                b.piece_base[k] = reinterpret_cast<uint32_t*>(take(pl.pbase_bytes));
                b.part[k] = reinterpret_cast<uint16_t*>(take(pl.part_bytes));
                b.state[k] = reinterpret_cast<uint32_t*>(take(pl.state_bytes));
                b.dkey[k] = reinterpret_cast<uint32_t*>(take(pl.dkey_bytes));
                b.sum_word[k] = reinterpret_cast<uint32_t*>(take(pl.sumw_bytes));
                b.sum_part[k] = reinterpret_cast<uint16_t*>(take(pl.sump_bytes));
                max_tiles = std::max(max_tiles, (uint32_t)pl.ntiles);
                max_scan_blocks = std::max(max_scan_blocks, (uint32_t)((pl.ntiles + kColChunk - 1) / kColChunk));
                max_piece_warps = std::max(max_piece_warps, (uint32_t)((pl.max_pieces + 31) / 32));
                max_chunks = std::max(max_chunks, (uint32_t)pl.max_chunks);
            }
            if ((e = flush()) != cudaSuccess) return fail(e);
        }
    }

    static_assert(sizeof(uint64_t) == sizeof(unsigned long long), "");
    if ((e = cudaMemcpyAsync(matches, d_matches, sizeof(uint64_t) * (size_t)nseg, cudaMemcpyDeviceToHost, stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(stream)) != cudaSuccess)
        return fail(e);
    return Status::kOk;
}

}  // namespace dlt
