// estimator.cu — LTU-semantics LZ match estimator on the GPU (see estimator.h).
//
// The restated CPU algorithm is a sequential scan with a 2^H-entry "last 3-byte key seen in this bucket" table,
// G positions at a time (G compares against the table as it was before the group, then G updates).  Hence:
// position p is a match iff the most recent earlier position of the SAME BUCKET that lies in an EARLIER group
// holds the same key (an untouched bucket holds 0, so only key 0 matches it).
//
// Design ("sequential table machine", round 2 — replaces the partition + piece-thread pipeline of round 1):
// the table IS kept, in shared memory, one per CTA, and the stream is cut into chunks, one CTA per chunk.
//   * An entry is 16 bits.  key -> P = key * GOLDEN (mod 2^32) is a bijection, the bucket is H bits of P, and 16 more
//     bits of P identify the key among the keys of that bucket (checked exhaustively for every supported parameter
//     set: tests/test_ltu_params.py).  0xFFFF means "untouched"; the one key per bucket whose 16 bits are 0xFFFF is
//     stored as 0xFFFE, which no key of such a bucket uses (same exhaustive check).  2^16 entries = 128 KiB.
//   * 15 PRODUCER warps turn 256 consecutive positions at a time (a batch = 8 rows of 32, lane = position in the row)
//     into 8-byte commands {entry address, tag to store}.  A ROW is one window: a position whose predecessor of the
//     same bucket lies in an earlier group of the row is decided on the spot (equal packets <=> a match) and the
//     answer of the table is ignored for it; the last position of a bucket in the row decides what the entry holds
//     afterwards.  Rows are screened through a per-warp byte scratch (which lanes could share a bucket at all?) and only
//     those lanes pay for a MATCH.ANY.  Commands travel through a ring of 16 slots in shared memory.
//   * ONE TABLE warp consumes the ring in stream order: per row one load and one store of all 32 lanes at the same
//     addresses.  A warp's shared-memory instructions are performed in order, so the reference's sequential semantics
//     hold exactly, for any data: a flat texture (every position in one bucket) and random bytes cost the same.  What the
//     loads found goes back to the producer (16 bits per position), which counts the matches one ring cycle later.
//     The table warp is ~30 instructions per batch, unrolled over the ring: it is the serial part of the machine.
//   * Chunks of one stream run concurrently.  A chunk that does not start its segment does not know the table it
//     inherits: a position that finds its entry untouched records its tag in first_seen[bucket][position & 3] (global),
//     the chunk publishes its final table, and one small RESOLVE kernel walks the chunks of a segment in order, one
//     thread per bucket, fully coalesced, and adds the matches of those first touches.
// DRAM traffic: the stream is read once (1 byte per position); hand-over state is 640 KiB per chunk.
// Measured (B200, 4 x 32 Mi positions): 1.30 ms for the first version of this design (one producer lane per 8-position
// window, four table rounds per row), 0.47 ms now; the steps are in DESIGN.md.
#include "estimator.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

// -DDLT_EST_CHECK=1: every shared-memory address the table machine computes is range-checked and a violation traps (a build
// for the tests: compute-sanitizer is not available on the GPU pool).  Off in the product build.
#ifndef DLT_EST_CHECK
#define DLT_EST_CHECK 0
#endif
#define DLT_EST_ASSERT(cond)                 \
    do {                                     \
        if (DLT_EST_CHECK && !(cond)) __trap(); \
    } while (0)

namespace dlt {
namespace {

std::atomic<uint64_t> g_est_launches{0};

std::mutex g_params_mutex;
LtuParams g_params;

constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kUntouched = 0xFFFFu;   // table entry: no position of this bucket seen yet
constexpr uint32_t kAlias = 0xFFFEu;       // stored instead of a tag of 0xFFFF

constexpr int kRows = 8;                                // rows of 32 consecutive positions per batch: one row = one WINDOW = one round of the table warp
constexpr int kBatchPos = 32 * kRows;                  // positions per batch (one producer warp iteration)
constexpr int kSlotBytes = kRows * 32 * 8;             // one batch of commands: 8 bytes per position
constexpr int kSeenBytes = kRows * 32 * 2;             // one batch of answers: what every position found in the table (16 bits)
constexpr int kWarps = 20;                             // the table warp and 19 producers (16 / 18 / 20 warps: 0.688 / 0.679 / 0.664 ms for the 64 MiB search;
                                                       // 96 registers x 640 threads and 223 KiB of shared memory: the SM is full)
constexpr int kRing = kWarps;                          // ring slots (even: the table warp's loop is unrolled over them in pairs)
constexpr int kSinkBytes = 64;                         // where positions that do not exist load and store: one 16-bit word per lane
constexpr int kRowGroup = 2;                            // rows a producer resolves side by side (measured: 1: 0.77, 2: 0.69, 4: 0.70, 8: 0.73 ms)
constexpr int kScratchSlots = 2048;                    // per producer warp: byte slots of the row screen (a power of two)
// The table warp and kWarps - 1 producers.  (Keeping the table warp's scheduler to itself — producers only on the other
// three — did not make it faster: its pace is the LSU's dispatch rate of one instruction per 4 cycles and per warp, with
// or without neighbours; ncu r02_seq_v3 .. v5.)
constexpr int kProducers = kWarps - 1;
constexpr int kSeqThreads = kWarps * 32;
constexpr int kMaxTableEntries = 1 << 16;              // per CTA: 128 KiB of 16-bit entries
constexpr int kTargetChunks = 148;                     // one chunk per SM when there is enough work
constexpr uint32_t kMinChunkBatches = 128;             // 32 Ki positions: below this a chunk's fixed costs dominate
static_assert(kRing >= kProducers && kRing % 2 == 0, "every producer holds one batch in the ring");

// Number of positions the reference loop visits: groups starting at i = 0, G, 2G, .. while i < len - 7.
inline size_t ltu_positions(size_t len, int group) {
    const size_t end = len > (size_t)kLtuTailGuard ? len - kLtuTailGuard : 0;
    return (end + group - 1) / group * group;
}

struct SeqParams {
    uint32_t hash_bits;       // H
    uint32_t sb;              // packet = bucket << sb | tag;  sb = 32 - max(H, 16)
    uint32_t tag_mask;        // (1 << sb) - 1
    uint32_t keep_mask;       // packet bits that select the table entry of this part (the bucket-part bits are dropped)
    uint32_t part_shift;      // part of a packet = (packet >> part_shift) & part_mask (the top bucket bits)
    uint32_t part_mask;       // parts - 1
    uint32_t table_bytes;     // entries per part * 2
};

struct SeqChunk {
    const uint8_t* data;            // segment base (any alignment)
    unsigned long long npos;        // positions of the segment
    unsigned long long first_pos;   // first position of the chunk (multiple of kBatchPos)
    uint32_t nbatches;
    uint32_t slot;                  // index of the segment's result in matches[]
    uint32_t part;                  // which part of the bucket space this CTA owns (tables above 128 KiB are split)
    uint32_t first;                 // 1: the chunk starts its segment (the inherited table is known: all zero)
    uint16_t* out_state;            // [entries]: the chunk's final table (nullptr: nobody needs it)
    uint16_t* first_seen;           // [entries][4]: tags of the heads that found their bucket untouched (first == 0)
};

struct SeqResolve {
    const uint16_t* out_base;       // [nchunks - 1][entries]: final tables of chunks 0 .. n-2
    const uint16_t* first_base;     // [nchunks - 1][entries][4]: first_seen of chunks 1 .. n-1
    uint32_t nchunks;
    uint32_t slot;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Ring hand-over by sequence numbers in shared memory.  ready[slot] = t + 1 once batch t sits in its slot (a release
// store by lane 0 after __syncwarp: MEMBAR.CTA + STS); consumed = t + 1 once the table warp has answered batch t (a plain
// store after the stores of the answers: a warp's shared-memory instructions are performed in order).  Readers poll with
// acquire loads, which are plain LDS on sm_100: ~30 cycles instead of the 90-150 of an mbarrier wait — the table warp is a
// single warp and every cycle of its loop is on the critical path.
__device__ __forceinline__ uint32_t ld_acquire_shared(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_shared(uint32_t* p, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_volatile_shared(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// Blocking wait with a guard: a protocol bug must end in a CUDA error, never in a hung GPU.
// SLEEP: back off between polls, longer the further away the awaited value is (the producers run ahead of the table warp;
// their polls go through the same LSU as the table warp's loads and stores: v3 issued ~200 polls per batch).
template <bool SLEEP>
__device__ __forceinline__ void wait_at_least(const uint32_t* p, uint32_t want) {
    uint32_t spins = 0;
    for (;;) {
        const int32_t behind = (int32_t)(want - ld_acquire_shared(p));
        if (behind <= 0) break;
        if (SLEEP) __nanosleep(32u * (uint32_t)behind);
        if (++spins > (1u << 24)) __trap();
    }
}

// packet = bucket << sb | tag.  key -> key * GOLDEN is a bijection, so equal packets <=> equal keys.
template <bool FAST16, bool TOP>
__device__ __forceinline__ uint32_t packet_of(uint32_t key, const SeqParams& prm) {
    const uint32_t p = key * kLtuGoldenRatio;
    uint32_t pkt = p;   // FAST16 (H = 16, index from the top bits): the packet is the product itself
    if (!FAST16) {
        uint32_t bucket, tag;
        if (TOP) {
            bucket = p >> (32 - prm.hash_bits);
            tag = p & prm.tag_mask;
        } else {
            bucket = p & ((1u << prm.hash_bits) - 1u);
            tag = (p >> prm.hash_bits) & prm.tag_mask;
        }
        pkt = (bucket << prm.sb) | tag;
    }
    if ((~pkt & 0xFFFFu) == 0u) pkt ^= 1u;   // low 16 bits 0xFFFF ("untouched") -> 0xFFFE (kAlias)
    return pkt;
}

// ---- producers --------------------------------------------------------------------------------------------------
// A batch = 8 rows of 32 consecutive positions, lane = position inside the row.  A ROW is a window: everything that can be
// decided inside it is decided here.  With grp(i) = i / G:
//   pred(i) = the highest lane j with grp(j) < grp(i) and the bucket of i -> i is a match iff the packets are equal (same packet
//             <=> same key) and what the table holds is irrelevant; no such j -> i is a HEAD: it is a match iff the table holds its tag
//   tail    = the highest lane of a bucket: the bucket's entry holds its tag after the row
// Two heads of one bucket lie in the same group (else the later one has a pred): the table warp may perform all loads of a
// row, then all stores of the row.  The lane writes one 8-byte command {entry address, tag to store}: EVERY lane loads its
// entry (the answer of a lane that is not a head is ignored) and every lane stores the tag of its bucket's tail (the lanes
// of one bucket store the same value to the same address): the table warp needs no predicate and no second address.
//
// Finding the lanes of one bucket is a MATCH.ANY, and MATCH.ANY is one unit per SM that spends 2 cycles per DISTINCT value:
// 64 cycles for a row of 32 different buckets (tools/microbench/match_bench.cu), which made the whole kernel run at exactly
// 8 x 64 cycles per batch.  But a lane whose bucket no other lane of the row has needs no resolution — it is head and tail —
// and that is the common case wherever MATCH.ANY is slow.  So each row is first screened through a per-warp scratch of 2048
// byte slots: every lane stores its lane id at slot (bucket mod 2048) and loads it back.  Two lanes of one bucket share a
// slot, so at least one of them reads a foreign id (a "loser"; no false negatives, whatever else hits the slot); every loser
// names the lane it read (one REDUX.OR), which tells the winners of contested slots.  The lanes of uncontested slots enter
// MATCH.ANY with one common value: it only pays for the contested buckets, and rows without a loser skip it.

// What a producer remembers of a batch until the table warp has answered it.
struct Pending {
    uint32_t pkt[kRows];   // packet of this lane's position in every row
    uint32_t heads;        // bit r: the position of row r is a head (its answer counts)
};

// FULL: every position of the batch is valid and the table is not split (all batches but the last one of a segment).
// Returns the number of matches decided inside the rows; cmd_a / cmd_b = the commands.
template <int G, bool TOP, bool FAST16, bool FULL>
__device__ __forceinline__ uint32_t produce_batch(const uint32_t (&wa)[kRows], const uint32_t (&wb)[kRows], const uint32_t fsh,
                                                  const uint32_t pos0, const uint32_t chunk_valid, const SeqParams& prm,
                                                  const uint32_t part, const uint32_t table_addr, const uint32_t sink_addr,
                                                  const uint32_t scratch_addr, uint32_t (&cmd_a)[kRows], uint32_t (&cmd_b)[kRows],
                                                  Pending& pend) {
    const unsigned lane = threadIdx.x & 31;
    const uint32_t below = (1u << (lane & ~(unsigned)(G - 1))) - 1u;   // the lanes of earlier groups
    uint32_t count = 0;
    pend.heads = 0;
    // kRowGroup rows at a time, every step for all rows of the group before the next step: the chain store -> load -> vote ->
    // REDUX -> MATCH -> shuffle of ONE row is ~400 cycles of latency, and a producer that walks its rows one by one spends
    // 4700 cycles per batch (ncu r02_seq_v8: with the table warp at ~30 instructions per batch the producers set the pace).
#pragma unroll
    for (int r0 = 0; r0 < kRows; r0 += kRowGroup) {
        uint32_t pkt[kRowGroup], bucket[kRowGroup], slot[kRowGroup], winner[kRowGroup];
        bool act[kRowGroup];
#pragma unroll
        for (int i = 0; i < kRowGroup; i++) {
            const int r = r0 + i;
            pkt[i] = packet_of<FAST16, TOP>(__funnelshift_r(wa[r], wb[r], fsh) & kLtuKeyMask, prm);
            pend.pkt[r] = pkt[i];
            act[i] = true;
            if (!FULL) {
                act[i] = pos0 + 32u * r < chunk_valid;
                if (prm.part_mask) act[i] = act[i] && ((pkt[i] >> prm.part_shift) & prm.part_mask) == part;
            }
            bucket[i] = FAST16 ? pkt[i] >> 16 : pkt[i] >> prm.sb;
            // screen: which lanes share their scratch slot with another lane?  (scratch_addr is 2 KiB aligned.)  A slot may
            // also be overwritten by a lane of another row of the group: that only adds false alarms.
            slot[i] = scratch_addr | (bucket[i] & (kScratchSlots - 1));
            DLT_EST_ASSERT((scratch_addr & (kScratchSlots - 1)) == 0 && slot[i] - scratch_addr < (uint32_t)kScratchSlots);
            if (FULL || act[i]) asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(slot[i]), "r"(lane) : "memory");
        }
        if (!FULL) __syncwarp();   // (store then load of one converged warp: performed in order)
#pragma unroll
        for (int i = 0; i < kRowGroup; i++) {
            winner[i] = lane;
            if (FULL || act[i]) asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(winner[i]) : "r"(slot[i]) : "memory");
        }
        bool any_loser = false;
#pragma unroll
        for (int i = 0; i < kRowGroup; i++) any_loser |= winner[i] != lane;
        const bool resolve = __any_sync(kFull, any_loser);
        bool head[kRowGroup];
        uint32_t store_tag[kRowGroup];
#pragma unroll
        for (int i = 0; i < kRowGroup; i++) head[i] = act[i], store_tag[i] = pkt[i];
        if (resolve) {
            // phase by phase over the rows of the group (warp-level operations are issued in program order: a row's chain
            // written out on its own would be waited for link by link)
            uint32_t named[kRowGroup], same[kRowGroup], lower[kRowGroup], pred_pkt[kRowGroup];
            bool contested[kRowGroup];
            // the lane that won a contested slot must learn it too: every loser names its winner
#pragma unroll
            for (int i = 0; i < kRowGroup; i++) named[i] = __reduce_or_sync(kFull, winner[i] != lane ? 1u << winner[i] : 0u);
#pragma unroll
            for (int i = 0; i < kRowGroup; i++) {
                // may share its bucket with another lane of the row.  (A loser may name a lane of the OTHER row of the group: a
                // false alarm for that lane here, and nothing at all if the position does not exist in this row.)
                contested[i] = act[i] && (winner[i] != lane || ((named[i] >> lane) & 1u));
                // the others all enter with one common value (a real bucket has at most 17 bits)
                same[i] = __match_any_sync(kFull, contested[i] ? bucket[i] : 0xFFFFFFFFu);
            }
#pragma unroll
            for (int i = 0; i < kRowGroup; i++) {
                if (!contested[i]) same[i] = 1u << lane;   // alone in its bucket (the common value's class means nothing)
                lower[i] = same[i] & below;
                // (no predecessor: lane -1 = lane 31 of the shuffle, and the value is not looked at)
                pred_pkt[i] = __shfl_sync(kFull, pkt[i], 31u - (uint32_t)__clz(lower[i]));
            }
#pragma unroll
            for (int i = 0; i < kRowGroup; i++) store_tag[i] = __shfl_sync(kFull, pkt[i], 31u - (uint32_t)__clz(same[i]));
#pragma unroll
            for (int i = 0; i < kRowGroup; i++) {
                head[i] = act[i] && lower[i] == 0u;
                count += lower[i] != 0u && pred_pkt[i] == pkt[i];
            }
        }
#pragma unroll
        for (int i = 0; i < kRowGroup; i++) {
            const int r = r0 + i;
            const uint32_t entry = FAST16 ? table_addr + 2u * bucket[i] : table_addr + (((pkt[i] & prm.keep_mask) >> prm.sb) << 1);
            cmd_a[r] = act[i] ? entry : sink_addr;
            cmd_b[r] = store_tag[i];   // the table warp stores the low 16 bits
            pend.heads |= (uint32_t)head[i] << r;
        }
    }
    return count;
}

// The table warp has answered a batch: count the heads that found their own tag; UNKNOWN (the chunk does not start its
// segment): a head that found its entry untouched records its tag for the resolve kernel.
template <bool FAST16>
__device__ __forceinline__ uint32_t collect_batch(const Pending& pend, const uint4 seen4, const SeqParams& prm, uint16_t* first_seen) {
    const uint32_t packed[4] = {seen4.x, seen4.y, seen4.z, seen4.w};
    uint32_t count = 0, untouched = 0;
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        const uint32_t seen = (r & 1) ? packed[r >> 1] >> 16 : packed[r >> 1] & 0xFFFFu;
        const uint32_t tag = pend.pkt[r] & 0xFFFFu;
        count += ((pend.heads >> r) & 1u) && seen == tag;
        untouched |= (uint32_t)(seen == kUntouched) << r;
    }
    // heads that found their entry untouched: common in the first batches of a chunk, rare afterwards (a lane-level branch)
    untouched &= pend.heads;
    if (first_seen && untouched) {
#pragma unroll
        for (int r = 0; r < kRows; r++)
            if ((untouched >> r) & 1u) {
                const uint32_t idx = FAST16 ? pend.pkt[r] >> 16 : (pend.pkt[r] & prm.keep_mask) >> prm.sb;
                DLT_EST_ASSERT(idx * 2u < prm.table_bytes);
                first_seen[(size_t)idx * 4] = (uint16_t)pend.pkt[r];   // first_seen is already advanced by position & 3
            }
    }
    return count;
}

template <int G, bool TOP, bool FAST16>
__device__ __forceinline__ uint32_t producer_warp(const SeqChunk& ck, const uint32_t nb, const uint32_t pi, const SeqParams& prm,
                                                  const uint32_t table_addr, const uint32_t sink_base, const uint32_t scratch_addr,
                                                  uint8_t* ring, const uint8_t* seen_base, uint32_t* ready, const uint32_t* consumed) {
    const unsigned lane = threadIdx.x & 31;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(ck.data) & 3u);
    const uint32_t* base = reinterpret_cast<const uint32_t*>(ck.data - sh) + (ck.first_pos >> 2);   // chunk-relative words
    // positions of the segment still ahead at the start of the chunk (clamped: batches are 32-bit quantities)
    const unsigned long long ahead = ck.npos - ck.first_pos;
    const uint32_t chunk_valid = ahead > 0xFFFFFF00ull ? 0xFFFFFF00u : (uint32_t)ahead;
    const uint32_t maxw = (sh + chunk_valid + 1) >> 2;   // chunk-relative word of the last byte any valid position reads
    const uint32_t sink_addr = sink_base + 2u * lane;
    uint16_t* first_seen = ck.first_seen ? ck.first_seen + (lane & 3) : nullptr;   // position & 3 == lane & 3
    // the key of position (row base + lane) starts at byte lane + sh of the row's first word
    const uint32_t wofs = (lane + sh) >> 2, fsh = ((lane + sh) & 3u) * 8u;
    auto load = [&](uint32_t t, uint32_t (&a)[kRows], uint32_t (&b)[kRows]) {
        const uint32_t w0 = t * (kBatchPos / 4) + wofs;
        if (w0 + 8u * (kRows - 1) + 1 <= maxw) {   // (a lane-level branch only in the last batch of a segment)
#pragma unroll
            for (int r = 0; r < kRows; r++) {
                a[r] = __ldg(base + w0 + 8 * r);
                b[r] = __ldg(base + w0 + 8 * r + 1);
            }
        } else {
#pragma unroll
            for (int r = 0; r < kRows; r++) {
                const uint32_t w = w0 + 8u * r;
                a[r] = w <= maxw ? __ldg(base + w) : 0u;
                b[r] = w + 1 <= maxw ? __ldg(base + w + 1) : 0u;
            }
        }
    };
    auto answers = [&](uint32_t t) { return reinterpret_cast<const uint4*>(seen_base + (size_t)(t % kRing) * kSeenBytes)[lane]; };
    uint32_t count = 0;
    uint32_t na[kRows], nbw[kRows];
#pragma unroll
    for (int r = 0; r < kRows; r++) na[r] = nbw[r] = 0u;
    if (pi < nb) load(pi, na, nbw);
    Pending pend;        // the batch this warp produced last (t - kProducers), not yet collected
    bool have_pending = false;
    uint32_t t = pi;
    // (Writing the loop out in pairs with two register sets that change roles — no copies of the 16 prefetched words and the
    // 9 remembered ones — made ptxas spill: 0.66 -> 0.72 ms.)
    for (; t < nb; t += kProducers) {
        uint32_t ca[kRows], cb[kRows];
#pragma unroll
        for (int r = 0; r < kRows; r++) ca[r] = na[r], cb[r] = nbw[r];
        if (t + kProducers < nb) load(t + kProducers, na, nbw);
        const uint32_t pos0 = t * kBatchPos + lane;   // chunk-relative
        uint32_t cmd_a[kRows], cmd_b[kRows];
        Pending fresh;
        if (prm.part_mask == 0u && (t + 1) * (uint32_t)kBatchPos <= chunk_valid)
            count += produce_batch<G, TOP, FAST16, true>(ca, cb, fsh, pos0, chunk_valid, prm, ck.part, table_addr, sink_addr, scratch_addr, cmd_a, cmd_b, fresh);
        else
            count += produce_batch<G, TOP, FAST16, false>(ca, cb, fsh, pos0, chunk_valid, prm, ck.part, table_addr, sink_addr, scratch_addr, cmd_a, cmd_b, fresh);
        // the table warp's answers to this warp's previous batch (this also frees the slot: t - kProducers >= t - kRing)
        if (have_pending) {
            wait_at_least<true>(consumed, t - kProducers + 1);
            count += collect_batch<FAST16>(pend, answers(t - kProducers), prm, first_seen);
        }
        pend = fresh;
        have_pending = true;
        // the commands of rows 2k and 2k + 1 of a lane travel together (one 128-bit load of the table warp)
        uint4* sw = reinterpret_cast<uint4*>(ring + (size_t)(t % kRing) * kSlotBytes) + lane;
#pragma unroll
        for (int k = 0; k < kRows / 2; k++) sw[k * 32] = make_uint4(cmd_a[2 * k], cmd_b[2 * k], cmd_a[2 * k + 1], cmd_b[2 * k + 1]);
        __syncwarp();
        if (lane == 0) st_release_shared(ready + t % kRing, t + 1);   // (MEMBAR.CTA + STS; measured: free on this side of the ring)
    }
    if (have_pending) {
        wait_at_least<true>(consumed, t - kProducers + 1);
        count += collect_batch<FAST16>(pend, answers(t - kProducers), prm, first_seen);
    }
    return count;
}

// ---- the table warp ---------------------------------------------------------------------------------------------
// Lane c handles position c of every row; per row ONE load and ONE store of all 32 lanes at the same addresses, in program
// order (a warp's shared-memory instructions are performed in order: this IS the reference's sequential loop).  The warp
// does nothing else: what the loads found goes back to the producer of the batch (16 bits per position), which counts the
// matches and keeps the books of the hand-over between chunks.  Every instruction of this loop is on the critical path of
// the kernel, and the LSU takes one instruction per 4 cycles from a warp: 4 command loads + 16 + 1 answer store per batch.
__device__ __forceinline__ uint32_t ld_volatile_shared(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
// Software pipeline: while batch t goes through the table, the commands of batch t + 1 are already on their way from the
// ring and the answers of batch t - 1 (whose loads have long completed) are packed and stored: no instruction of the loop
// waits for a load it has just issued.  The loop is unrolled over the 16 ring slots, so every ring address is an immediate:
// a single warp issues a dependent instruction every ~4.5 cycles, so the pace of the table warp — of the whole kernel — is
// its instruction count (ncu r02_seq_exp5: with the table accesses themselves removed the loop still took 260 cycles per
// batch, for ~55 instructions of slot arithmetic, flag handling and packing).
struct TableWarp {
    uint32_t ring_lane, seen_lane, ready_addr, consumed_addr;   // shared-memory addresses (this lane's 16 bytes of slot 0)
    uint32_t nb;
    unsigned lane;
    uint32_t table_lo, table_hi, sink_lo;   // DLT_EST_CHECK: what a command may address
    __device__ __forceinline__ bool valid_target(uint32_t a) const {
        return (a & 1u) == 0 && ((a >= table_lo && a < table_hi) || (a >= sink_lo && a < sink_lo + kSinkBytes));
    }

    template <int SLOT>
    __device__ __forceinline__ void load_commands(uint4 (&c)[kRows / 2]) const {
#pragma unroll
        for (int k = 0; k < kRows / 2; k++)   // volatile: after the flag has been seen, never hoisted above it
            asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(c[k].x), "=r"(c[k].y), "=r"(c[k].z), "=r"(c[k].w)
                         : "r"(ring_lane + (uint32_t)(SLOT * kSlotBytes + 512 * k))
                         : "memory");
    }
    template <int SLOT>
    __device__ __forceinline__ uint32_t load_flag() const {
        uint32_t v;
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(ready_addr + (uint32_t)(SLOT * 4)) : "memory");
        return v;
    }
    template <int SLOT>
    __device__ __forceinline__ void wait_flag(uint32_t flag, uint32_t want) const {
        uint32_t spins = 0;
        while (flag < want) {   // (sequence numbers start at 1 and a chunk has far fewer than 2^32 batches)
            flag = load_flag<SLOT>();
            if (++spins > (1u << 26)) __trap();
        }
    }
    // answers of batch t (slot SLOT) and the signal that they are there: the producer may collect them and refill the slot
    template <int SLOT>
    __device__ __forceinline__ void store_answers(const uint32_t (&seen)[kRows], uint32_t t) const {
        asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(seen_lane + (uint32_t)(SLOT * kSeenBytes)), "r"(__byte_perm(seen[0], seen[1], 0x5410)),
                     "r"(__byte_perm(seen[2], seen[3], 0x5410)), "r"(__byte_perm(seen[4], seen[5], 0x5410)), "r"(__byte_perm(seen[6], seen[7], 0x5410))
                     : "memory");
        if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(consumed_addr), "r"(t + 1) : "memory");
    }
    // batch t = base + SLOT: `cur` holds its commands, `flag` = ready[slot of t + 1] as loaded one step ago.
    // INNER: 0 < t and t + 1 < nb are known (all steps of a ring round that is neither the first nor the last).
    template <int SLOT, bool INNER>
    __device__ __forceinline__ void step(uint32_t base, uint4 (&cur)[kRows / 2], uint4 (&nxt)[kRows / 2], uint32_t (&seen)[kRows],
                                         const uint32_t (&prev)[kRows], uint32_t& flag) const {
        constexpr int S1 = (SLOT + 1) % kRing, S2 = (SLOT + 2) % kRing, SP = (SLOT + kRing - 1) % kRing;
        const uint32_t t = base + SLOT;
        if (INNER || t + 1 < nb) {
            wait_flag<S1>(flag, t + 2);
            load_commands<S1>(nxt);
        }
        flag = load_flag<S2>();   // (a stale value of a slot that is never filled is never looked at)
#pragma unroll
        for (int k = 0; k < kRows / 2; k++) {
            DLT_EST_ASSERT(valid_target(cur[k].x) && valid_target(cur[k].z));
            asm volatile(
                "ld.volatile.shared.u16 %0, [%2];\n st.volatile.shared.u16 [%2], %3;\n"
                "ld.volatile.shared.u16 %1, [%4];\n st.volatile.shared.u16 [%4], %5;"
                : "=&r"(seen[2 * k]), "=&r"(seen[2 * k + 1])
                : "r"(cur[k].x), "r"(cur[k].y), "r"(cur[k].z), "r"(cur[k].w)
                : "memory");
        }
        if (INNER || t > 0) store_answers<SP>(prev, t - 1);
    }
    template <int SLOT, bool INNER>
    __device__ __forceinline__ bool steps(uint32_t base, uint4 (&a)[kRows / 2], uint4 (&b)[kRows / 2], uint32_t (&sa)[kRows],
                                          uint32_t (&sb)[kRows], uint32_t& flag) const {
        if constexpr (SLOT < kRing) {
            if (!INNER && base + SLOT >= nb) return false;
            step<SLOT, INNER>(base, a, b, sa, sb, flag);
            if (!INNER && base + SLOT + 1 >= nb) return false;
            step<SLOT + 1, INNER>(base, b, a, sb, sa, flag);
            return steps<SLOT + 2, INNER>(base, a, b, sa, sb, flag);
        } else {
            return true;
        }
    }
    __device__ __forceinline__ void run() const {
        if (nb == 0) return;
        uint4 a[kRows / 2], b[kRows / 2];
        uint32_t sa[kRows], sb[kRows];
#pragma unroll
        for (int r = 0; r < kRows; r++) sa[r] = sb[r] = 0u;
        wait_flag<0>(0u, 1u);
        load_commands<0>(a);
        uint32_t flag = load_flag<1>();
        uint32_t base = 0;
        if (steps<0, false>(base, a, b, sa, sb, flag)) {   // the first round of the ring (t = 0 has no previous answers)
            for (base = kRing; base + kRing < nb; base += kRing) steps<0, true>(base, a, b, sa, sb, flag);
            steps<0, false>(base, a, b, sa, sb, flag);     // the last round (it may be empty: nb a multiple of the ring)
        }
        // the answers of the last batch: in sa if nb is odd, else in sb
        const uint32_t last = nb - 1;
        uint32_t fin[kRows];
#pragma unroll
        for (int r = 0; r < kRows; r++) fin[r] = (last & 1u) ? sb[r] : sa[r];
        asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(seen_lane + (last % kRing) * kSeenBytes), "r"(__byte_perm(fin[0], fin[1], 0x5410)),
                     "r"(__byte_perm(fin[2], fin[3], 0x5410)), "r"(__byte_perm(fin[4], fin[5], 0x5410)), "r"(__byte_perm(fin[6], fin[7], 0x5410))
                     : "memory");
        if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(consumed_addr), "r"(nb) : "memory");
    }
};

// G: positions per group of the reference loop.  TOP: index from the top bits of the product (else: low bits).
// FAST16: H = 16 and TOP (the restated crate's parameters): the packet is the product itself.
template <int G, bool TOP, bool FAST16>
__global__ void __launch_bounds__(kSeqThreads, 1)
ltu_seq_kernel(const SeqChunk* __restrict__ chunks, unsigned long long* __restrict__ matches, const SeqParams prm) {
    extern __shared__ __align__(16) uint8_t seq_smem[];
    // table | pad to 2 KiB | scratch (kProducers x 2 KiB) | ring | answers | sink | ready[kRing], consumed
    uint16_t* table = reinterpret_cast<uint16_t*>(seq_smem);
    const uint32_t scratch0 = (smem_u32(seq_smem) + prm.table_bytes + (kScratchSlots - 1)) & ~(uint32_t)(kScratchSlots - 1);
    uint8_t* ring = seq_smem + (scratch0 - smem_u32(seq_smem)) + (size_t)kProducers * kScratchSlots;
    uint8_t* seen = ring + (size_t)kRing * kSlotBytes;
    uint8_t* sink = seen + (size_t)kRing * kSeenBytes;
    uint32_t* ready = reinterpret_cast<uint32_t*>(sink + kSinkBytes);
    uint32_t* consumed = ready + kRing;
    if (DLT_EST_CHECK) {
        uint32_t dyn_bytes;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
        DLT_EST_ASSERT(smem_u32(consumed) + 4u <= smem_u32(seq_smem) + dyn_bytes);
    }

    const SeqChunk ck = chunks[blockIdx.x];
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nb = ck.nbatches;

    // ---- set-up: table = untouched everywhere (a chunk that starts its segment inherits the reference's zeroed
    // table: bucket(0) holds key 0 -> tag 0, every other bucket can only mismatch, which "untouched" does too)
    {
        uint4* t4 = reinterpret_cast<uint4*>(table);
        const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
        for (uint32_t i = tid; i < prm.table_bytes / 16; i += kSeqThreads) t4[i] = ones;
        if (ck.first_seen) {
            uint4* f4 = reinterpret_cast<uint4*>(ck.first_seen);
            for (uint32_t i = tid; i < prm.table_bytes / 4; i += kSeqThreads) f4[i] = ones;   // entries * 8 bytes
        }
        if (tid <= (unsigned)kRing) ready[tid] = 0;   // ready[0 .. kRing) and consumed
    }
    __syncthreads();
    if (tid == 0 && ck.first && ck.part == 0) table[0] = 0;
    __syncthreads();

    uint32_t count = 0;
    if (warp == 0) {
        TableWarp{smem_u32(ring) + 16u * lane, smem_u32(seen) + 16u * lane, smem_u32(ready), smem_u32(consumed), nb, lane,
                  smem_u32(table), smem_u32(table) + prm.table_bytes, smem_u32(sink)}.run();
    } else {
        const uint32_t pi = warp - 1;
        count = producer_warp<G, TOP, FAST16>(ck, nb, pi, prm, smem_u32(table), smem_u32(sink), scratch0 + pi * kScratchSlots, ring, seen, ready,
                                              consumed);
    }
    for (int o = 16; o; o >>= 1) count += __shfl_xor_sync(kFull, count, o);
    if (lane == 0 && count) atomicAdd(&matches[ck.slot], (unsigned long long)count);
    __syncthreads();
    if (ck.out_state) {
        const uint4* t4 = reinterpret_cast<const uint4*>(table);
        uint4* o4 = reinterpret_cast<uint4*>(ck.out_state);
        for (uint32_t i = tid; i < prm.table_bytes / 16; i += kSeqThreads) o4[i] = t4[i];
    }
}

// ---- resolve: the first touches of the chunks that did not know their inherited table ------------------------
// One thread per bucket walks the chunks of a segment in order: `cur` = the entry the next chunk inherits.
__global__ void __launch_bounds__(256) ltu_resolve_kernel(const SeqResolve* __restrict__ res, unsigned long long* __restrict__ matches,
                                                          const uint32_t entries) {
    const SeqResolve r = res[blockIdx.y];
    const uint32_t b = blockIdx.x * 256 + threadIdx.x;
    uint32_t count = 0;
    if (b < entries) {
        uint32_t cur = r.out_base[b];
        for (uint32_t c = 1; c < r.nchunks; c++) {
            const uint2 f = __ldg(reinterpret_cast<const uint2*>(r.first_base + ((size_t)(c - 1) * entries + b) * 4));
            uint32_t o = kUntouched;
            if (c + 1 < r.nchunks) o = r.out_base[(size_t)c * entries + b];
            // an empty slot holds 0xFFFF and a recorded tag never does: nothing may match while `cur` is still untouched
            if (cur != kUntouched) {
                count += (f.x & 0xFFFFu) == cur;
                count += (f.x >> 16) == cur;
                count += (f.y & 0xFFFFu) == cur;
                count += (f.y >> 16) == cur;
            }
            if (o != kUntouched) cur = o;
        }
    }
    for (int o = 16; o; o >>= 1) count += __shfl_xor_sync(kFull, count, o);
    __shared__ uint32_t ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = count;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; i++) t += ws[i];
        if (t) atomicAdd(&matches[r.slot], t);
    }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline size_t result_bytes(int nseg) { return align_up((size_t)(nseg > 0 ? nseg : 1) * sizeof(uint64_t), 256); }

struct Geometry {
    SeqParams prm;
    uint32_t parts;     // CTAs per chunk of the stream (bucket space split when the table exceeds 128 KiB)
    uint32_t entries;   // table entries per part
    size_t smem_bytes;
};
Geometry geometry(const LtuParams& p) {
    Geometry g{};
    const uint32_t h = (uint32_t)p.hash_bits;
    g.prm.hash_bits = h;
    g.prm.sb = 32 - std::max<uint32_t>(h, 16);
    g.prm.tag_mask = (1u << g.prm.sb) - 1u;
    const uint32_t total = 1u << h;
    g.parts = total > (uint32_t)kMaxTableEntries ? total / kMaxTableEntries : 1u;
    g.entries = total / g.parts;
    uint32_t part_bits = 0;
    while ((1u << part_bits) < g.parts) part_bits++;
    // packet = bucket << sb | tag: the top part_bits of the bucket select the part
    g.prm.part_shift = part_bits ? g.prm.sb + h - part_bits : 0;
    g.prm.part_mask = g.parts - 1u;
    g.prm.keep_mask = part_bits ? ((g.entries - 1u) << g.prm.sb) | g.prm.tag_mask : ~0u;
    g.prm.table_bytes = g.entries * 2u;
    g.smem_bytes = g.prm.table_bytes + (size_t)kScratchSlots + (size_t)kProducers * kScratchSlots + (size_t)kRing * (kSlotBytes + kSeenBytes) + kSinkBytes +
                   (kRing + 1) * sizeof(uint32_t) + 12;   // table, alignment pad, scratch, ring, answers, sink, flags
    return g;
}

// The chunks of a call.  Every segment is cut into chunks of about `cb` batches; cb is chosen so that a large call
// is about one chunk per SM (one wave: the chunks are equal, a second, nearly empty wave would double the time).
struct Plan {
    std::vector<SeqChunk> chunks;
    std::vector<SeqResolve> resolves;
    size_t chunk_desc_bytes = 0, resolve_desc_bytes = 0, state_bytes = 0;
    size_t total() const { return chunk_desc_bytes + resolve_desc_bytes + state_bytes; }
};
// nb batches in chunks of about cb: n equal chunks of `per` batches (the last one may be shorter, never empty)
struct ChunkCut {
    size_t n, per;
};
ChunkCut chunks_of(size_t nb, size_t cb) {
    if (nb == 0) return {0, 0};
    const size_t n0 = (nb + cb - 1) / cb, per = (nb + n0 - 1) / n0;
    return {(nb + per - 1) / per, per};
}
uint32_t chunk_batches(const LtuSegment* segs, int nseg, const LtuParams& p, const Geometry& g) {
    size_t total = 0;
    for (int i = 0; i < nseg; i++) total += (ltu_positions(segs[i].len, p.group) + kBatchPos - 1) / kBatchPos;
    total *= g.parts;
    size_t cb = std::max<size_t>((total + kTargetChunks - 1) / kTargetChunks, kMinChunkBatches);
    // rounding every segment up to whole chunks can overshoot one wave: grow cb until the chunks fit (or it is hopeless)
    for (int it = 0; it < 64; it++) {
        size_t n = 0;
        for (int i = 0; i < nseg; i++) {
            const size_t nb = (ltu_positions(segs[i].len, p.group) + kBatchPos - 1) / kBatchPos;
            n += chunks_of(nb, cb).n * g.parts;
        }
        if (n <= (size_t)kTargetChunks || n >= 4 * (size_t)kTargetChunks) break;
        cb += std::max<size_t>(cb / 32, 1);
    }
    return (uint32_t)std::min<size_t>(cb, 0xFFFFFFFFu);
}
// fill == false: only the sizes.  state / d_chunks / d_resolves: device addresses the descriptors point into.
Plan make_plan(const LtuSegment* segs, int nseg, const LtuParams& p, const Geometry& g, bool fill, uint8_t* state) {
    Plan pl;
    const uint32_t cb = chunk_batches(segs, nseg, p, g);
    const size_t out_bytes = (size_t)g.entries * 2, first_bytes = (size_t)g.entries * 8;
    size_t nchunks_total = 0;
    for (int i = 0; i < nseg; i++) {
        const size_t npos = ltu_positions(segs[i].len, p.group);
        const size_t nb = (npos + kBatchPos - 1) / kBatchPos;
        if (nb == 0) continue;
        const ChunkCut cut = chunks_of(nb, cb);
        const size_t n = cut.n, per = cut.per;
        for (uint32_t part = 0; part < g.parts; part++) {
            uint8_t* out_base = fill ? state + pl.state_bytes : nullptr;
            uint8_t* first_base = fill ? out_base + (n - 1) * out_bytes : nullptr;
            if (n > 1) {
                pl.state_bytes += (n - 1) * (out_bytes + first_bytes);
                if (fill)
                    pl.resolves.push_back(SeqResolve{reinterpret_cast<const uint16_t*>(out_base),
                                                     reinterpret_cast<const uint16_t*>(first_base), (uint32_t)n, (uint32_t)i});
                else
                    pl.resolves.emplace_back();
            }
            for (size_t c = 0; c < n; c++) {
                nchunks_total++;
                if (!fill) continue;
                SeqChunk ck{};
                ck.data = segs[i].d_ptr;
                ck.npos = npos;
                ck.first_pos = (unsigned long long)c * per * kBatchPos;
                ck.nbatches = (uint32_t)std::min(per, nb - c * per);
                ck.slot = (uint32_t)i;
                ck.part = part;
                ck.first = c == 0;
                ck.out_state = c + 1 < n ? reinterpret_cast<uint16_t*>(out_base + c * out_bytes) : nullptr;
                ck.first_seen = c > 0 ? reinterpret_cast<uint16_t*>(first_base + (c - 1) * first_bytes) : nullptr;
                pl.chunks.push_back(ck);
            }
        }
    }
    pl.chunk_desc_bytes = align_up(std::max<size_t>(nchunks_total, 1) * sizeof(SeqChunk), 256);
    pl.resolve_desc_bytes = align_up(std::max<size_t>(pl.resolves.size(), 1) * sizeof(SeqResolve), 256);
    pl.state_bytes = align_up(pl.state_bytes, 256);
    return pl;
}

template <int G, bool TOP, bool FAST16>
cudaError_t launch_seq(const SeqChunk* d_chunks, int n, unsigned long long* d_matches, const Geometry& g, cudaStream_t stream) {
    // function attributes are per device: set it on every call (microseconds), a process may drive several GPUs
    cudaError_t e = cudaFuncSetAttribute(ltu_seq_kernel<G, TOP, FAST16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
    if (e != cudaSuccess) return e;
    ltu_seq_kernel<G, TOP, FAST16><<<n, kSeqThreads, g.smem_bytes, stream>>>(d_chunks, d_matches, g.prm);
    return cudaGetLastError();
}

}  // namespace

bool ltu_params_supported(const LtuParams& p) {
    return p.hash_bits >= 12 && p.hash_bits <= 17 && (p.group == 1 || p.group == 4);
}
bool ltu_set_params(const LtuParams& p) {
    if (!ltu_params_supported(p)) return false;
    std::lock_guard<std::mutex> lock(g_params_mutex);
    g_params = p;
    return true;
}
LtuParams ltu_params() {
    std::lock_guard<std::mutex> lock(g_params_mutex);
    return g_params;
}

uint64_t estimator_launch_count() { return g_est_launches.load(std::memory_order_relaxed); }

// Scratch for `nseg` segments with `batches` batches in total: an upper bound of what any plan of such a call needs,
// monotone in both arguments (callers size the scratch from a worst-case segment list and then estimate a subset).
//   chunks          <= nseg * parts + kTargetChunks      (chunk_batches keeps the cut near one wave; rounding adds one per segment)
//   chunks with hand-over state <= min(kTargetChunks, batches * parts / kMinChunkBatches): 10 bytes per table entry each
static size_t scratch_bound(size_t nseg, size_t batches, const Geometry& g) {
    nseg = std::max<size_t>(nseg, 1);
    const size_t max_chunks = nseg * g.parts + (size_t)kTargetChunks;
    const size_t cut = std::min<size_t>((size_t)kTargetChunks, batches * g.parts / kMinChunkBatches + g.parts);
    return result_bytes((int)std::min<size_t>(nseg, 0x7FFFFFFF)) + align_up(max_chunks * sizeof(SeqChunk), 256) +
           align_up(nseg * g.parts * sizeof(SeqResolve), 256) + align_up(cut * (size_t)g.entries * 10, 256);
}

void LtuScratchMeter::add(size_t len) {
    nseg_++;
    batches_ += (len + kBatchPos - 1) / kBatchPos;   // >= the batches of ltu_positions(len) for any group size
}
size_t LtuScratchMeter::bytes() const { return scratch_bound(nseg_, batches_, geometry(ltu_params())); }

size_t ltu_scratch_bytes(const LtuSegment* segs, int nseg) {
    size_t batches = 0;
    for (int i = 0; i < nseg; i++) batches += (segs[i].len + kBatchPos - 1) / kBatchPos;
    return scratch_bound((size_t)std::max(nseg, 0), batches, geometry(ltu_params()));
}

// Page-locked staging for the descriptors and the results of a call, one per host thread (a call synchronises its stream
// before it returns, so the buffer is free again).  Grown on demand, never freed: a few KiB for one texture, 64 bytes per
// chunk for a directory.  nullptr if page-locked memory cannot be had: the copies then go through pageable memory.
static uint8_t* call_staging(size_t bytes) {
    struct Buf {
        uint8_t* p = nullptr;
        size_t cap = 0;
    };
    thread_local Buf buf;
    if (buf.cap < bytes) {
        if (buf.p) cudaFreeHost(buf.p);
        buf.p = nullptr, buf.cap = 0;
        const size_t want = std::max<size_t>(align_up(bytes, 4096), 64u << 10);
        void* q = nullptr;
        if (cudaHostAlloc(&q, want, cudaHostAllocPortable) != cudaSuccess) {   // one buffer per host thread, whichever device it drives
            (void)cudaGetLastError();
            return nullptr;
        }
        buf.p = static_cast<uint8_t*>(q), buf.cap = want;
    }
    return buf.p;
}

// One launch of the table machine over all chunks of all segments, one resolve launch if any segment was cut; the zeroed
// result slots and both descriptor tables go up in ONE copy from page-locked memory, the results come back in one, and the
// call synchronises once.  A directory of small textures is one chunk per endpoint stream.
Status ltu_matches_device(const LtuSegment* segs, int nseg, uint64_t* matches, cudaStream_t stream, uint8_t* scratch,
                          size_t scratch_bytes) {
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
    if (nseg <= 0) return Status::kOk;
    const LtuParams p = ltu_params();
    const Geometry g = geometry(p);
    auto fail = [](cudaError_t e) {
        note_cuda_error(e);
        return Status::kCudaError;
    };
    const Plan sizes = make_plan(segs, nseg, p, g, false, nullptr);
    if (scratch_bytes < result_bytes(nseg) + sizes.total()) return Status::kOutOfMemory;
    const size_t res_bytes = result_bytes(nseg);
    unsigned long long* d_matches = reinterpret_cast<unsigned long long*>(scratch);
    uint8_t* d_chunk_desc = scratch + res_bytes;
    uint8_t* d_resolve_desc = d_chunk_desc + sizes.chunk_desc_bytes;
    uint8_t* d_state = d_resolve_desc + sizes.resolve_desc_bytes;
    const Plan pl = make_plan(segs, nseg, p, g, true, d_state);

    // [results = 0 | chunk descriptors | resolve descriptors] is contiguous on the device: one upload
    const size_t head_bytes = res_bytes + sizes.chunk_desc_bytes + sizes.resolve_desc_bytes;
    uint8_t* staging = call_staging(head_bytes);
    cudaError_t e;
    if (staging) {
        std::memset(staging, 0, res_bytes);
        std::memcpy(staging + res_bytes, pl.chunks.data(), pl.chunks.size() * sizeof(SeqChunk));
        std::memcpy(staging + res_bytes + sizes.chunk_desc_bytes, pl.resolves.data(), pl.resolves.size() * sizeof(SeqResolve));
        const size_t used = pl.resolves.empty() ? res_bytes + pl.chunks.size() * sizeof(SeqChunk)
                                                : res_bytes + sizes.chunk_desc_bytes + pl.resolves.size() * sizeof(SeqResolve);
        e = cudaMemcpyAsync(scratch, staging, used, cudaMemcpyHostToDevice, stream);
    } else {
        // pageable sources: the driver stages the bytes before the call returns, the vectors may die afterwards
        e = cudaMemsetAsync(d_matches, 0, res_bytes, stream);
        if (e == cudaSuccess && !pl.chunks.empty())
            e = cudaMemcpyAsync(d_chunk_desc, pl.chunks.data(), pl.chunks.size() * sizeof(SeqChunk), cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess && !pl.resolves.empty())
            e = cudaMemcpyAsync(d_resolve_desc, pl.resolves.data(), pl.resolves.size() * sizeof(SeqResolve), cudaMemcpyHostToDevice, stream);
    }
    if (e != cudaSuccess) return fail(e);
    if (!pl.chunks.empty()) {
        const SeqChunk* dc = reinterpret_cast<const SeqChunk*>(d_chunk_desc);
        const int n = (int)pl.chunks.size();
        if (p.group == 4 && p.index_top && p.hash_bits == 16) e = launch_seq<4, true, true>(dc, n, d_matches, g, stream);   // the restated crate
        else if (p.group == 4) e = p.index_top ? launch_seq<4, true, false>(dc, n, d_matches, g, stream) : launch_seq<4, false, false>(dc, n, d_matches, g, stream);
        else e = p.index_top ? launch_seq<1, true, false>(dc, n, d_matches, g, stream) : launch_seq<1, false, false>(dc, n, d_matches, g, stream);
        if (e != cudaSuccess) return fail(e);
        g_est_launches.fetch_add(1, std::memory_order_relaxed);
        if (!pl.resolves.empty()) {
            ltu_resolve_kernel<<<dim3((g.entries + 255) / 256, (unsigned)pl.resolves.size()), 256, 0, stream>>>(
                reinterpret_cast<const SeqResolve*>(d_resolve_desc), d_matches, g.entries);
            if ((e = cudaGetLastError()) != cudaSuccess) return fail(e);
            g_est_launches.fetch_add(1, std::memory_order_relaxed);
        }
    }
    void* back = staging ? static_cast<void*>(staging) : static_cast<void*>(matches);
    if ((e = cudaMemcpyAsync(back, d_matches, sizeof(uint64_t) * (size_t)nseg, cudaMemcpyDeviceToHost, stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(stream)) != cudaSuccess)
        return fail(e);
    if (staging) std::memcpy(matches, staging, sizeof(uint64_t) * (size_t)nseg);
    return Status::kOk;
}

}  // namespace dlt
