// bcn_kernels.cu — BCn lossless transform / untransform kernels for B200 (sm_100a).
//
// What they compute (reference: /root/reference/src/core, all little-endian, integer only):
//   transform   : N interleaved BCn blocks  ->  per-field streams (bcn_layout.h), colour endpoints
//                 optionally YCoCg-R decorrelated (common/src/color_565/decorrelate.rs:101-300)
//                 and optionally split into c0 / c1 (and a0 / a1 for BC3).
//   untransform : the exact inverse (decorrelate.rs:148-344).
// One kernel family covers BC1/BC2/BC3 x every settings combination; the format, the splits and
// the YCoCg variant are template parameters, the stream table comes from bcn_layout.h.
//
// Shape of a tiled kernel (HBM-bound streaming permutation, 2 x len bytes of traffic):
//   * a CTA owns a tile of 16 KiB of blocks (2048 BC1 / 1024 BC2,BC3 blocks);
//   * transform: every thread issues 4 coalesced 128-bit `ld.global.nc.L1::no_allocate` loads up
//     front, does the colour arithmetic on two RGB565 values packed in one 32-bit register, and
//     scatters the fields into a shared-memory staging area that is laid out per stream; after one
//     barrier the CTA streams each stream segment to HBM with 128-bit stores;
//   * untransform mirrors it: 16-byte cp.async (LDGSTS) stream copies straight into shared memory ->
//     per-thread gather -> recorrelate -> one 128-bit block store per thread and vector.
//   * stream bases need not be 16-byte aligned (in the reference layout they sit at N*k bytes, and
//     real mip chains give odd N): the staging area of stream s is shifted by (address & 127) so
//     shared and global addresses are congruent mod 128; the 16-byte aligned interior moves as
//     128-bit vectors in whole 128-byte lines per warp and only the (at most two) ragged edges of a
//     segment (< 16 bytes each) move bytewise.
#include "bcn_kernels.h"

#include <algorithm>
#include <atomic>
#include <cstdint>

namespace dlt {
namespace {

#ifndef DLT_MIN_CTAS
#define DLT_MIN_CTAS 6  // resident CTAs per SM the tiled kernels are compiled for (register cap 40)
#endif
constexpr int kThreads = 256;
constexpr int kUnroll = 4;  // 128-bit vectors per thread per tile
// Staging of stream s is shifted by (global address & 127): shared and global addresses are then
// congruent mod 128, so 16-byte chunk k of a segment is the same chunk on both sides AND a warp's 32
// consecutive chunks cover four whole 128-byte lines (no partially written sectors inside a tile).
#ifndef DLT_SHIFT_ALIGN
#define DLT_SHIFT_ALIGN 128
#endif
constexpr int kShiftAlign = DLT_SHIFT_ALIGN;
static_assert(kThreads * kUnroll * 16 == kTileBytes, "tile geometry");

std::atomic<uint64_t> g_launches{0};

// ------------------------------------------------------------------------------------------------
// Streaming global accesses
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream8(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream16(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
// -DDLT_BULK_STORE=1: the RAGGED transform writes every stream segment with one TMA bulk copy (cp.async.bulk
// shared -> global, UBLKCP) instead of a 128-bit store loop.  Bit-exact (parity + property tests) and 8 % fewer
// instructions, but no consistent gain at 1 GiB - 3 blocks: BC1 6.69-6.84 -> 6.71-6.78 TB/s, BC2 6.34-6.73 -> 6.66-6.70,
// BC3 5.11-6.08 -> 5.34-6.66 (one setting +15 %, one -3 %); ncu shows the ragged kernel waiting on its loads (long
// scoreboard 14 vs 10.7 per issue in the aligned kernel at identical DRAM bytes), not on the store loop.  Off by default.
#ifndef DLT_BULK_STORE
#define DLT_BULK_STORE 0
#endif
__device__ __forceinline__ void bulk_store(void* gmem, const void* smem, uint32_t bytes) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_wait() {
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void stg_stream8(void* p, const uint2& v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// ------------------------------------------------------------------------------------------------
// YCoCg-R on two RGB565 colours packed in one 32-bit register (lanes [15:0] and [31:16]).
// Every intermediate is a 5-bit field per lane; +32 per lane before a subtraction keeps borrows
// from crossing lanes.  Restates Color565::{de,re}correlate_ycocg_r_var{1,2,3}.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kM5 = 0x001F001Fu, kM4 = 0x000F000Fu, kM1 = 0x00010001u, kBias = 0x00200020u;

template <int VAR>
__device__ __forceinline__ uint32_t decorrelate2(uint32_t v) {
    if constexpr (VAR == kNone) {
        return v;
    } else {
        const uint32_t r = (v >> 11) & kM5, g = (v >> 6) & kM5, gl = (v >> 5) & kM1, b = v & kM5;
        const uint32_t co = (r + kBias - b) & kM5;
        const uint32_t t = (b + ((co >> 1) & kM4)) & kM5;
        const uint32_t cg = (g + kBias - t) & kM5;
        const uint32_t y = (t + ((cg >> 1) & kM4)) & kM5;
        if constexpr (VAR == kVariant1) return (y << 11) | (co << 6) | (gl << 5) | cg;
        if constexpr (VAR == kVariant2) return (gl << 15) | (y << 10) | (co << 5) | cg;
        return (y << 11) | (co << 6) | (cg << 1) | gl;
    }
}

template <int VAR>
__device__ __forceinline__ uint32_t recorrelate2(uint32_t v) {
    if constexpr (VAR == kNone) {
        return v;
    } else {
        uint32_t y, co, cg, gl;
        if constexpr (VAR == kVariant1) {
            y = (v >> 11) & kM5, co = (v >> 6) & kM5, gl = (v >> 5) & kM1, cg = v & kM5;
        } else if constexpr (VAR == kVariant2) {
            gl = (v >> 15) & kM1, y = (v >> 10) & kM5, co = (v >> 5) & kM5, cg = v & kM5;
        } else {
            y = (v >> 11) & kM5, co = (v >> 6) & kM5, cg = (v >> 1) & kM5, gl = v & kM1;
        }
        const uint32_t t = (y + kBias - ((cg >> 1) & kM4)) & kM5;
        const uint32_t g = (cg + t) & kM5;
        const uint32_t b = (t + kBias - ((co >> 1) & kM4)) & kM5;
        const uint32_t r = (b + co) & kM5;
        return (r << 11) | (g << 6) | (gl << 5) | b;
    }
}

// Runtime-variant forms for the byte-granular kernel (same arithmetic, one colour in the low lane).
__device__ __forceinline__ uint32_t decorrelate_rt(uint32_t v, int var) {
    switch (var) {
        case kVariant1: return decorrelate2<kVariant1>(v);
        case kVariant2: return decorrelate2<kVariant2>(v);
        case kVariant3: return decorrelate2<kVariant3>(v);
        default: return v;
    }
}
__device__ __forceinline__ uint32_t recorrelate_rt(uint32_t v, int var) {
    switch (var) {
        case kVariant1: return recorrelate2<kVariant1>(v);
        case kVariant2: return recorrelate2<kVariant2>(v);
        case kVariant3: return recorrelate2<kVariant3>(v);
        default: return v;
    }
}

// ------------------------------------------------------------------------------------------------
// experimental::normalize_blocks on one BC1 block held as (c0 | c1 << 16, indices).
// Restates decode_bc1_block (core/dxt-lossless-transform-bc1/src/util/bc1_decode.rs:7-63), Color565::{red,green,blue}
// (common/src/color_565/mod.rs:154-191), from_rgb (:108-115) and normalize_blocks_impl / write_normalized_solid_color_block
// (experimental/normalize_blocks/normalize.rs:104-247).  Returns the BlockCase: 0 cannot normalize (block unchanged),
// 1 fully transparent (block = 0xFF..), 2 solid round-trippable colour (block rewritten per `mode`).
// Fast reject: when more than one index VALUE is used and c0 != c1 the pixels cannot all be equal (the expanded
// endpoints differ by >= 4 in some channel, so every interpolated entry differs from both endpoints and from the
// other interpolated entry, and the transparent entry differs in alpha) — nearly every real block leaves here.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t expand565(uint32_t c) {   // r | g << 8 | b << 16, 8 bits each
    const uint32_t r = (c >> 11) & 31u, g = (c >> 5) & 63u, b = c & 31u;
    return ((r << 3) | (r >> 2)) | (((g << 2) | (g >> 4)) << 8) | (((b << 3) | (b >> 2)) << 16);
}
__device__ __forceinline__ uint32_t mix3(uint32_t a, uint32_t b, uint32_t wa, uint32_t wb, uint32_t div) {
    // per channel (wa * a + wb * b) / div on packed 8-bit channels
    uint32_t out = 0;
#pragma unroll
    for (int sft = 0; sft < 24; sft += 8) {
        const uint32_t x = wa * ((a >> sft) & 0xFFu) + wb * ((b >> sft) & 0xFFu);
        out |= (div == 3 ? (x * 0xAAABu) >> 17 : x >> 1) << sft;   // x <= 765: exact
    }
    return out;
}
// Slow part of normalize_bc1_block: every pixel uses interpolated entry `sel` (2 or 3) of a block with c0 != c1.
// Returns the RGBA8888 pixel (alpha 0 = the transparent entry of the three-colour mode).
__device__ __noinline__ uint32_t bc1_interpolated_entry(uint32_t c0, uint32_t c1, uint32_t sel) {
    const uint32_t e0 = expand565(c0), e1 = expand565(c1);
    if (c0 > c1) return (sel == 2 ? mix3(e0, e1, 2, 1, 3) : mix3(e0, e1, 1, 2, 3)) | 0xFF000000u;
    return sel == 2 ? (mix3(e0, e1, 1, 1, 2) | 0xFF000000u) : 0u;
}
__device__ __forceinline__ int normalize_bc1_block(uint32_t& c01, uint32_t& idx, const int mode) {
    const uint32_t c0 = c01 & 0xFFFFu, c1 = c01 >> 16;
    uint32_t c565;
    if (c0 == c1) {
        // entries 0, 1 and 2 decode to the colour itself (c0 > c1 is false: entry 2 = (e0 + e0) / 2), entry 3 is transparent
        const uint32_t threes = idx & (idx >> 1) & 0x55555555u;
        if (threes != 0) {
            if (threes != 0x55555555u) return 0;   // opaque and transparent pixels
            c01 = idx = 0xFFFFFFFFu;
            return 1;
        }
        c565 = c0;   // an expanded RGB565 colour always converts back to itself
    } else {
        const uint32_t sel = idx & 3u;
        if (idx != sel * 0x55555555u) return 0;   // two index values in use: the pixels differ (see above)
        if (sel < 2) {
            c565 = sel == 0 ? c0 : c1;
        } else {
            const uint32_t px = bc1_interpolated_entry(c0, c1, sel);
            if ((px >> 24) == 0) {   // fully transparent
                c01 = idx = 0xFFFFFFFFu;
                return 1;
            }
            const uint32_t r = px & 0xFFu, g = (px >> 8) & 0xFFu, b = (px >> 16) & 0xFFu;
            c565 = ((r & 0xF8u) << 8) | ((g & 0xFCu) << 3) | (b >> 3);
            if ((expand565(c565) | 0xFF000000u) != px) return 0;   // not representable as one RGB565 colour
        }
    }
    if (mode == kNormColor0Only) c01 = c565, idx = 0;
    else if (mode == kNormReplicateColor) c01 = c565 | (c565 << 16), idx = 0;
    return 2;   // mode None: the reference writes the source block back
}

// ---- the same decision for EIGHT blocks per lane, warp-cooperative -----------------------------------------------
// A block painted entirely with an interpolated palette entry (c0 != c1, all indices 2 or 3) needs ~90 instructions of
// arithmetic; encoders that store flat areas through their single-colour tables (stb_dxt and friends) produce exactly
// that, so it is a common block in textures with flat regions.  Taken lane by lane under a branch, every one of a lane's
// eight blocks that needs it stalls the other 31 lanes (ncu on adversarial data: 10.3 of 32 lanes active per
// instruction).  Here the fast cases are decided inline and the (colours, entry) pairs of the slow blocks of the whole
// warp are compacted into a shared-memory queue (ballot + popc ranks); then ALL lanes work the queue 32 entries at a
// time and the results go back by the same ranks.  Used by the stand-alone normalization pass (3.6 instead of 2.35 TB/s
// on adversarial data, and faster on ordinary data too).  The fused transform keeps the per-block branch: with the
// queue it lost 7 % on ordinary data (6.1 instead of 6.6 TB/s) for +22 % on randomly scattered slow blocks, and real flat
// regions are spatially coherent, so whole warps take the branch together.
// Result word: BlockCase in bits 0-1 (0 unchanged, 1 transparent, 2 solid colour), the RGB565 colour in bits 16-31.
constexpr uint32_t kNeedsEntry = 3u;
__device__ __forceinline__ uint32_t classify_bc1_block(uint32_t c01, uint32_t idx) {
    const uint32_t c0 = c01 & 0xFFFFu, c1 = c01 >> 16;
    if (c0 == c1) {
        const uint32_t threes = idx & (idx >> 1) & 0x55555555u;
        if (threes != 0) return threes == 0x55555555u ? 1u : 0u;
        return 2u | (c0 << 16);
    }
    const uint32_t sel = idx & 3u;
    if (idx != sel * 0x55555555u) return 0u;
    if (sel < 2) return 2u | ((sel ? c1 : c0) << 16);
    return kNeedsEntry;
}
__device__ __forceinline__ uint32_t classify_interpolated(uint32_t c01, uint32_t sel) {
    const uint32_t c0 = c01 & 0xFFFFu, c1 = c01 >> 16;
    const uint32_t e0 = expand565(c0), e1 = expand565(c1);
    uint32_t px;
    if (c0 > c1) px = (sel == 2 ? mix3(e0, e1, 2, 1, 3) : mix3(e0, e1, 1, 2, 3)) | 0xFF000000u;
    else px = sel == 2 ? (mix3(e0, e1, 1, 1, 2) | 0xFF000000u) : 0u;
    if ((px >> 24) == 0) return 1u;
    const uint32_t r = px & 0xFFu, g = (px >> 8) & 0xFFu, b = (px >> 16) & 0xFFu;
    const uint32_t c565 = ((r & 0xF8u) << 8) | ((g & 0xFCu) << 3) | (b >> 3);
    return (expand565(c565) | 0xFF000000u) != px ? 0u : (2u | (c565 << 16));
}
__device__ __forceinline__ void apply_normalization(uint32_t& c01, uint32_t& idx, const uint32_t res, const int mode) {
    const uint32_t bcase = res & 3u, c565 = res >> 16;
    if (bcase == 1) c01 = idx = 0xFFFFFFFFu;
    else if (bcase == 2 && mode == kNormColor0Only) c01 = c565, idx = 0;
    else if (bcase == 2 && mode == kNormReplicateColor) c01 = c565 | (c565 << 16), idx = 0;
}
constexpr int kNormQueueEntries = 8 * 32;   // per warp: eight blocks per lane
// ALL 32 lanes of the warp must call this.  `queue` = this warp's kNormQueueEntries entries of shared memory.
__device__ __forceinline__ void classify8_warp(const uint4 (&v)[kUnroll], uint32_t (&res)[2 * kUnroll], uint2* queue) {
    static_assert(kUnroll == 4, "eight BC1 blocks per lane");
    bool any_slow = false;
#pragma unroll
    for (int s = 0; s < 2 * kUnroll; s++) {
        res[s] = classify_bc1_block((s & 1) ? v[s >> 1].z : v[s >> 1].x, (s & 1) ? v[s >> 1].w : v[s >> 1].y);
        any_slow |= res[s] == kNeedsEntry;
    }
    if (!__any_sync(0xFFFFFFFFu, any_slow)) return;   // the common case costs one vote
    const unsigned lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
    unsigned queued = 0;
#pragma unroll
    for (int s = 0; s < 2 * kUnroll; s++) {
        const bool slow = res[s] == kNeedsEntry;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, slow);
        if (slow) queue[queued + __popc(m & below)] = make_uint2((s & 1) ? v[s >> 1].z : v[s >> 1].x, ((s & 1) ? v[s >> 1].w : v[s >> 1].y) & 3u);
        queued += __popc(m);
    }
    __syncwarp();
    for (unsigned i = lane; i < queued; i += 32) {
        const uint2 q = queue[i];
        queue[i].x = classify_interpolated(q.x, q.y);
    }
    __syncwarp();
    queued = 0;
#pragma unroll
    for (int s = 0; s < 2 * kUnroll; s++) {
        const bool slow = res[s] == kNeedsEntry;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, slow);
        if (slow) res[s] = queue[queued + __popc(m & below)].x;
        queued += __popc(m);
    }
}

// ------------------------------------------------------------------------------------------------
// Compile-time view of one (format, split_alpha, split_colour) layout.
// ------------------------------------------------------------------------------------------------
template <int FMT, bool SA, bool SC, int HALO = 0>
struct Lay {
    static constexpr int NS = num_streams(FMT, SA, SC);
    static constexpr int BPB = block_bytes(FMT);
    static constexpr int T = kTileBytes / BPB;    // blocks per tile
    static constexpr int TS = T + HALO;           // blocks staged in shared memory (tile + halo of the next tile)
    static constexpr int BPV = 16 / BPB;          // blocks per 128-bit vector
    DLT_HD static constexpr int w(int s) { return stream_width(FMT, SA, SC, s); }
    // Staging region of stream s: w*TS payload bytes + kShiftAlign bytes of slack for the alignment shift.
    DLT_HD static constexpr int region(int s) { return TS * stream_prefix(FMT, SA, SC, s) + kShiftAlign * s; }
    static constexpr int kStageBytes = TS * BPB + kShiftAlign * NS;
    // 128-bit chunks per stream segment of a full tile (+ the shift, + one sector of halo).
    DLT_HD static constexpr int iters(int s) {
        return (w(s) * T / 16 + kShiftAlign / 16 + (HALO ? 2 : 0) + kThreads - 1) / kThreads;
    }
    // Stream indices of the logical fields.
    static constexpr int sAlpha = 0;                          // BC2 alpha:8 / BC3 a0a1:2 or a0:1
    [[maybe_unused]] static constexpr int sA1 = 1;            // BC3 split alpha only
    static constexpr int sAIdx = SA ? 2 : 1;                  // BC3 only
    static constexpr int sCol = FMT == 1 ? 0 : FMT == 2 ? 1 : sAIdx + 1;  // c0c1:4 or c0:2
    [[maybe_unused]] static constexpr int sC1 = sCol + 1;     // split colour only
    static constexpr int sIdx = NS - 1;
};

template <typename T>
__device__ __forceinline__ void sts(uint8_t* p, T v) {
    *reinterpret_cast<T*>(p) = v;
}
template <typename T>
__device__ __forceinline__ T lds(const uint8_t* p) {
    return *reinterpret_cast<const T*>(p);
}

// ------------------------------------------------------------------------------------------------
// Tiled transform kernel
// ------------------------------------------------------------------------------------------------
// One 128-bit vector of blocks (two BC1 blocks or one BC2/BC3 block starting at tile-relative block b0; blocks at or
// beyond `limit` are not staged): colour arithmetic + scatter of every field into the per-stream staging area.
template <int FMT, bool SA, bool SC, int VAR, int NORM, typename L>
__device__ __forceinline__ void stage_vector(uint8_t* stage, const int* sh, const int b0, const int limit, uint4 v) {
    if constexpr (FMT == 1) {
        const bool two = b0 + 1 < limit;
        if constexpr (NORM != kNormNone) {   // experimental::transform_bc1_with_normalize_blocks: normalize, then transform
            constexpr int kMode = NORM == kNormAllModesNone ? (int)kNormNone : NORM;   // transparent blocks only
            normalize_bc1_block(v.x, v.y, kMode);
            normalize_bc1_block(v.z, v.w, kMode);
        }
        const uint32_t ca = decorrelate2<VAR>(v.x), cb = decorrelate2<VAR>(v.z);
        uint8_t* pi = stage + L::region(L::sIdx) + sh[L::sIdx] + 4 * b0;
        sts<uint32_t>(pi, v.y);
        if (two) sts<uint32_t>(pi + 4, v.w);
        if constexpr (SC) {
            uint8_t* p0 = stage + L::region(L::sCol) + sh[L::sCol] + 2 * b0;
            uint8_t* p1 = stage + L::region(L::sC1) + sh[L::sC1] + 2 * b0;
            sts<uint16_t>(p0, (uint16_t)ca);
            sts<uint16_t>(p1, (uint16_t)(ca >> 16));
            if (two) {
                sts<uint16_t>(p0 + 2, (uint16_t)cb);
                sts<uint16_t>(p1 + 2, (uint16_t)(cb >> 16));
            }
        } else {
            uint8_t* pc = stage + L::region(L::sCol) + sh[L::sCol] + 4 * b0;
            sts<uint32_t>(pc, ca);
            if (two) sts<uint32_t>(pc + 4, cb);
        }
    } else {
        // BC2 / BC3: one block per vector: [x y] = alpha part, z = colours, w = indices
        if constexpr (FMT == 2) {
            sts<uint2>(stage + L::region(L::sAlpha) + sh[L::sAlpha] + 8 * b0, make_uint2(v.x, v.y));
        } else {
            if constexpr (SA) {
                stage[L::region(L::sAlpha) + sh[L::sAlpha] + b0] = (uint8_t)v.x;
                stage[L::region(L::sA1) + sh[L::sA1] + b0] = (uint8_t)(v.x >> 8);
            } else {
                sts<uint16_t>(stage + L::region(L::sAlpha) + sh[L::sAlpha] + 2 * b0, (uint16_t)v.x);
            }
            uint8_t* pa = stage + L::region(L::sAIdx) + sh[L::sAIdx] + 6 * b0;
            sts<uint16_t>(pa, (uint16_t)(v.x >> 16));
            sts<uint16_t>(pa + 2, (uint16_t)v.y);
            sts<uint16_t>(pa + 4, (uint16_t)(v.y >> 16));
        }
        const uint32_t c = decorrelate2<VAR>(v.z);
        if constexpr (SC) {
            sts<uint16_t>(stage + L::region(L::sCol) + sh[L::sCol] + 2 * b0, (uint16_t)c);
            sts<uint16_t>(stage + L::region(L::sC1) + sh[L::sC1] + 2 * b0, (uint16_t)(c >> 16));
        } else {
            sts<uint32_t>(stage + L::region(L::sCol) + sh[L::sCol] + 4 * b0, c);
        }
        sts<uint32_t>(stage + L::region(L::sIdx) + sh[L::sIdx] + 4 * b0, v.w);
    }
}

template <typename L>
__device__ __forceinline__ uint4 load_vector(const uint8_t* tin, const int j, const int limit) {
    const int b0 = j * L::BPV;
    if (b0 + L::BPV <= limit) return ldg_stream16(tin + (size_t)j * 16);
    if (L::BPV == 2 && b0 < limit) {
        const uint2 h = ldg_stream8(tin + (size_t)j * 16);
        return make_uint4(h.x, h.y, 0u, 0u);
    }
    return make_uint4(0u, 0u, 0u, 0u);
}

// RAGGED = false: every stream base is 16-byte aligned and every stream's total length is a multiple of 16; a tile's
// segments are whole 128-byte lines and nothing else is needed.
// RAGGED = true (odd block counts: real mip chains): segments start and end inside 32-byte sectors.  Writing such a
// sector from two tiles costs three PARTIAL L2 writes per stream and tile boundary, and a partial write into the
// ECC-protected L2 is ~5x a full one (measured: -25 % on BC2/BC3).  Instead a sector belongs to the tile that holds its
// FIRST byte: the tile also stages kHalo blocks of the next tile (3 % extra reads, L2 hits) and writes the boundary
// sector whole; only the first and the last sector of the launch's range are written bytewise.
constexpr int kHalo = 32;   // blocks: covers 31 bytes of the narrowest stream (1 byte per block)

template <int FMT, bool SA, bool SC, int VAR, bool RAGGED, int NORM>
__device__ __forceinline__ void transform_tile(const uint8_t* __restrict__ in, const StreamPtrs& out, const uint64_t nblocks) {
    using L = Lay<FMT, SA, SC, RAGGED ? kHalo : 0>;
    __shared__ __align__(16) uint8_t stage[L::kStageBytes];

    const int tid = threadIdx.x;
    const uint64_t tile_first = (uint64_t)blockIdx.x * L::T;
    const uint64_t left = nblocks - tile_first;
    const int nb = left < (uint64_t)L::T ? (int)left : L::T;        // blocks this tile owns
    const int nbs = left < (uint64_t)L::TS ? (int)left : L::TS;     // blocks it stages (owned + halo)

    // (address & 127) of each stream segment; w*T is a multiple of 128 so it is tile-independent.
    int sh[L::NS];
#pragma unroll
    for (int s = 0; s < L::NS; s++) sh[s] = (int)(reinterpret_cast<uintptr_t>(out.p[s]) & (kShiftAlign - 1));

    // ---- phase 1: 4 coalesced 128-bit loads in flight per thread; warp 0 also fetches the halo (a warp-uniform branch:
    // predicating a fifth vector through every warp costs +27 % issued instructions on the six-stream BC3 layout)
    const uint8_t* tin = in + tile_first * L::BPB;
    uint4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) v[u] = load_vector<L>(tin, u * kThreads + tid, nbs);
    uint4 vh = make_uint4(0u, 0u, 0u, 0u);
    const bool halo_lane = RAGGED && tid < kHalo / L::BPV;
    if (RAGGED && tid < 32) {
        if (halo_lane) vh = load_vector<L>(tin, kUnroll * kThreads + tid, nbs);
    }

    // ---- phase 2: colour arithmetic + scatter into the per-stream staging area
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
        const int b0 = (u * kThreads + tid) * L::BPV;
        if (b0 < nbs) stage_vector<FMT, SA, SC, VAR, NORM, L>(stage, sh, b0, nbs, v[u]);
    }
    if (RAGGED && tid < 32) {
        const int b0 = (kUnroll * kThreads + tid) * L::BPV;
        if (halo_lane && b0 < nbs) stage_vector<FMT, SA, SC, VAR, NORM, L>(stage, sh, b0, nbs, vh);
    }
#if DLT_BULK_STORE
    if (RAGGED) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async-proxy reads
#endif
    __syncthreads();

    // ---- phase 3: stream every segment out as 128-bit vectors.  Coordinates are relative to `gal`, the 128-byte
    // aligned address below the segment: the tile owns [lo_valid, hi_valid), has data up to `avail`.
#pragma unroll
    for (int s = 0; s < L::NS; s++) {
        if (out.p[s] == nullptr) continue;   // a stream nobody wants (the search only estimates the endpoint streams)
        const int w = L::w(s);
        const int lo_valid = sh[s], hi_valid = sh[s] + w * nb;
        uint8_t* gal = out.p[s] + (uint64_t)w * tile_first - sh[s];  // 128-byte aligned
        const uint8_t* reg = stage + L::region(s);
        int vec_lo, vec_hi;   // multiples of 16
        if constexpr (RAGGED) {
            const int avail = sh[s] + w * nbs;
            vec_lo = (lo_valid + 31) & ~31;                      // the sector holding lo_valid belongs to the tile before
            vec_hi = min((hi_valid + 31) & ~31, avail & ~15);    // ... and the one holding hi_valid - 1 to this tile
        } else {
            vec_lo = lo_valid, vec_hi = hi_valid;
        }
#if DLT_BULK_STORE
        // One bulk copy (TMA engine, shared -> global) per stream segment instead of a 128-bit store loop through every
        // thread: thread s issues segment s.  Source and destination are 16-byte aligned (the staging shift keeps them
        // congruent mod 128), the size is a multiple of 16.
        if (RAGGED && tid == s && vec_hi > vec_lo) bulk_store(gal + vec_lo, reg + vec_lo, (uint32_t)(vec_hi - vec_lo));
        if (RAGGED) continue;
#endif
#pragma unroll
        for (int it = 0; it < L::iters(s); it++) {
            const int lo = (it * kThreads + tid) << 4;
            if (lo >= vec_lo && lo + 16 <= vec_hi) stg_stream16(gal + lo, lds<uint4>(reg + lo));
        }
    }
#if DLT_BULK_STORE
    if (RAGGED && tid < L::NS) bulk_store_wait();   // the staging area must outlive the engine's reads
#endif
    // ... and, in the first / last tile of the launch's range only, the bytes outside whole sectors (< 32 each):
    // the range may be a shard, so nothing beyond it is touched.  Warp 0, one lane per byte.
    if constexpr (RAGGED) {
        const bool first_tile = blockIdx.x == 0, last_tile = left <= (uint64_t)L::TS;
        if ((first_tile || last_tile) && tid < 32) {
#pragma unroll
            for (int s = 0; s < L::NS; s++) {
                if (out.p[s] == nullptr) continue;
                const int w = L::w(s);
                const int lo_valid = sh[s], hi_valid = sh[s] + w * nb, avail = sh[s] + w * nbs;
                uint8_t* gal = out.p[s] + (uint64_t)w * tile_first - sh[s];
                const uint8_t* reg = stage + L::region(s);
                const int vec_lo = (lo_valid + 31) & ~31, vec_hi = min((hi_valid + 31) & ~31, avail & ~15);
                if (first_tile) {   // head: [lo_valid, vec_lo), clipped to the data
                    const int i = lo_valid + tid;
                    if (i < min(vec_lo, avail)) gal[i] = reg[i];
                }
                if (last_tile) {    // tail: what the vector stores could not cover of [.., min(owned sectors, data))
                    const int end = min((hi_valid + 31) & ~31, avail);
                    const int i = max(vec_hi, max(vec_lo, lo_valid)) + tid;
                    if (vec_hi >= vec_lo && i < end) gal[i] = reg[i];
                }
            }
        }
    }
}

// Resident CTAs per SM the transform kernels are compiled for.  The RAGGED kernels with many streams wait longer per tile
// (halo, more segment edges): compiled for 8 CTAs / SM (32 registers, no spills) every BC3 settings combination runs at
// 6.82-6.86 TB/s on 1 GiB - 3 blocks against 5.6-6.4 at 6, and the BC2 split-colour layouts at 6.83-6.87 against 6.62-6.80;
// the BC2 layouts with one colour stream (6.83 at 6, 6.27-6.67 at 8), BC1 (6.83-6.98 at 6) and the aligned kernels are
// best at 6 (profiles/r02_ragged_ctas.jsonl: 6 / 7 / 8 for every settings combination).
#ifdef DLT_RAGGED_CTAS
constexpr int transform_min_ctas(int fmt, bool sc, bool ragged) { return ragged ? DLT_RAGGED_CTAS : DLT_MIN_CTAS; }
#else
constexpr int transform_min_ctas(int fmt, bool sc, bool ragged) { return ragged && (fmt == 3 || (fmt == 2 && sc)) ? 8 : DLT_MIN_CTAS; }
#endif

template <int FMT, bool SA, bool SC, int VAR, bool RAGGED, int NORM = kNormNone>
__global__ void __launch_bounds__(kThreads, transform_min_ctas(FMT, SC, RAGGED))
    transform_tiled(const uint8_t* __restrict__ in, const StreamPtrs out, const uint64_t nblocks) {
    transform_tile<FMT, SA, SC, VAR, RAGGED, NORM>(in, out, nblocks);
}

// Many independent payloads with the same settings in ONE launch (the candidates of a batched best-settings search:
// a directory of small textures is launch-bound otherwise): blockIdx.y picks the payload, blockIdx.x its tile.
template <int FMT, bool SA, bool SC, int VAR, bool RAGGED>
__global__ void __launch_bounds__(kThreads, transform_min_ctas(FMT, SC, RAGGED)) transform_tiled_batch(const TransformBatchItem* __restrict__ items) {
    const TransformBatchItem it = items[blockIdx.y];
    using L = Lay<FMT, SA, SC>;
    if ((uint64_t)blockIdx.x * L::T >= it.nblocks) return;
    transform_tile<FMT, SA, SC, VAR, RAGGED, kNormNone>(it.in, it.out, it.nblocks);
}

// ------------------------------------------------------------------------------------------------
// Tiled untransform kernel
// ------------------------------------------------------------------------------------------------
template <int FMT, bool SA, bool SC, int VAR>
__device__ __forceinline__ void untransform_tile(const StreamPtrs& in, uint8_t* __restrict__ out, const uint64_t nblocks) {
    using L = Lay<FMT, SA, SC>;
    __shared__ __align__(16) uint8_t stage[L::kStageBytes];

    const int tid = threadIdx.x;
    const uint64_t tile_first = (uint64_t)blockIdx.x * L::T;
    const uint64_t left = nblocks - tile_first;
    const int nb = left < (uint64_t)L::T ? (int)left : L::T;

    int sh[L::NS];
#pragma unroll
    for (int s = 0; s < L::NS; s++) sh[s] = (int)(reinterpret_cast<uintptr_t>(in.p[s]) & (kShiftAlign - 1));

    // ---- phase 1: every stream segment of the tile goes global -> shared with 16-byte cp.async
    // (LDGSTS, L2-only caching): no staging registers, all copies of the tile in flight at once.
#pragma unroll
    for (int s = 0; s < L::NS; s++) {
        const int w = L::w(s);
        const int lo_valid = sh[s], hi_valid = sh[s] + w * nb;
        const uint8_t* gal = in.p[s] + (uint64_t)w * tile_first - sh[s];
        uint8_t* reg = stage + L::region(s);
        const int nch = (hi_valid + 15) >> 4;
#pragma unroll
        for (int it = 0; it < L::iters(s); it++) {
            const int k = it * kThreads + tid;
            const int lo = k << 4;
            if (k < nch && lo >= lo_valid && lo + 16 <= hi_valid) cp_async16(reg + lo, gal + lo);
        }
    }
    // ragged head / tail of every segment (< 16 bytes each): warp 0, lanes 0-15 head, lanes 16-31 tail
    if (tid < 32) {
#pragma unroll
        for (int s = 0; s < L::NS; s++) {
            const int lo_valid = sh[s], hi_valid = sh[s] + L::w(s) * nb;
            const int head_end = min(hi_valid, (lo_valid + 15) & ~15);
            const int tail_start = max(head_end, hi_valid & ~15);
            const int i = tid < 16 ? lo_valid + tid : tail_start + tid - 16;
            if (tid < 16 ? i < head_end : i < hi_valid)
                stage[L::region(s) + i] = (in.p[s] + (uint64_t)L::w(s) * tile_first - sh[s])[i];
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- phase 2: gather the fields of each block, recorrelate, one 128-bit block store per vector
    uint8_t* tout = out + tile_first * L::BPB;
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
        const int j = u * kThreads + tid;
        const int b0 = j * L::BPV;
        if (b0 >= nb) continue;
        uint4 r;
        if constexpr (FMT == 1) {
            const bool two = b0 + 1 < nb;
            const uint8_t* pi = stage + L::region(L::sIdx) + sh[L::sIdx] + 4 * b0;
            uint32_t ca, cb = 0;
            r.y = lds<uint32_t>(pi);
            r.w = two ? lds<uint32_t>(pi + 4) : 0u;
            if constexpr (SC) {
                const uint8_t* p0 = stage + L::region(L::sCol) + sh[L::sCol] + 2 * b0;
                const uint8_t* p1 = stage + L::region(L::sC1) + sh[L::sC1] + 2 * b0;
                ca = (uint32_t)lds<uint16_t>(p0) | ((uint32_t)lds<uint16_t>(p1) << 16);
                if (two) cb = (uint32_t)lds<uint16_t>(p0 + 2) | ((uint32_t)lds<uint16_t>(p1 + 2) << 16);
            } else {
                const uint8_t* pc = stage + L::region(L::sCol) + sh[L::sCol] + 4 * b0;
                ca = lds<uint32_t>(pc);
                if (two) cb = lds<uint32_t>(pc + 4);
            }
            r.x = recorrelate2<VAR>(ca);
            r.z = recorrelate2<VAR>(cb);
            if (two) stg_stream16(tout + (size_t)j * 16, r);
            else stg_stream8(tout + (size_t)j * 16, make_uint2(r.x, r.y));
        } else {
            if constexpr (FMT == 2) {
                const uint2 a = lds<uint2>(stage + L::region(L::sAlpha) + sh[L::sAlpha] + 8 * b0);
                r.x = a.x;
                r.y = a.y;
            } else {
                uint32_t a01;
                if constexpr (SA) {
                    a01 = (uint32_t)stage[L::region(L::sAlpha) + sh[L::sAlpha] + b0] |
                          ((uint32_t)stage[L::region(L::sA1) + sh[L::sA1] + b0] << 8);
                } else {
                    a01 = lds<uint16_t>(stage + L::region(L::sAlpha) + sh[L::sAlpha] + 2 * b0);
                }
                const uint8_t* pa = stage + L::region(L::sAIdx) + sh[L::sAIdx] + 6 * b0;
                r.x = a01 | ((uint32_t)lds<uint16_t>(pa) << 16);
                r.y = (uint32_t)lds<uint16_t>(pa + 2) | ((uint32_t)lds<uint16_t>(pa + 4) << 16);
            }
            uint32_t c;
            if constexpr (SC) {
                c = (uint32_t)lds<uint16_t>(stage + L::region(L::sCol) + sh[L::sCol] + 2 * b0) |
                    ((uint32_t)lds<uint16_t>(stage + L::region(L::sC1) + sh[L::sC1] + 2 * b0) << 16);
            } else {
                c = lds<uint32_t>(stage + L::region(L::sCol) + sh[L::sCol] + 4 * b0);
            }
            r.z = recorrelate2<VAR>(c);
            r.w = lds<uint32_t>(stage + L::region(L::sIdx) + sh[L::sIdx] + 4 * b0);
            stg_stream16(tout + (size_t)j * 16, r);
        }
    }
}


template <int FMT, bool SA, bool SC, int VAR>
__global__ void __launch_bounds__(kThreads, DLT_MIN_CTAS)
    untransform_tiled(const StreamPtrs in, uint8_t* __restrict__ out, const uint64_t nblocks) {
    untransform_tile<FMT, SA, SC, VAR>(in, out, nblocks);
}
// Many payloads with the same settings in one launch (see transform_tiled_batch).
template <int FMT, bool SA, bool SC, int VAR>
__global__ void __launch_bounds__(kThreads, DLT_MIN_CTAS) untransform_tiled_batch(const UntransformBatchItem* __restrict__ items) {
    const UntransformBatchItem it = items[blockIdx.y];
    using L = Lay<FMT, SA, SC>;
    if ((uint64_t)blockIdx.x * L::T >= it.nblocks) return;
    untransform_tile<FMT, SA, SC, VAR>(it.in, it.out, it.nblocks);
}

// ------------------------------------------------------------------------------------------------
// Byte-granular kernel: one thread per block, no alignment assumptions at all.  Only taken when a
// caller hands in pointers the tiled kernels cannot address (the reference accepts any alignment).
// ------------------------------------------------------------------------------------------------
struct RtLayout {
    int fmt, var, ns, norm;
    int w[kMaxStreams];
    int s_alpha, s_a1, s_aidx, s_col, s_c1, s_idx;
    bool sa, sc;
};

__global__ void __launch_bounds__(kThreads)
    bytewise_kernel(const RtLayout L, const bool inverse, const uint8_t* blocks_in, uint8_t* blocks_out,
                    const StreamPtrs streams, const uint64_t nblocks) {
    const uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= nblocks) return;
    const int bpb = L.fmt == 1 ? 8 : 16;
    const int coff = L.fmt == 1 ? 0 : 8;  // colour part offset inside the block
    uint8_t blk[16];
    if (!inverse) {
        for (int k = 0; k < bpb; k++) blk[k] = blocks_in[i * bpb + k];
        if (L.fmt == 2) {
            for (int k = 0; k < 8; k++) streams.p[L.s_alpha][i * 8 + k] = blk[k];
        } else if (L.fmt == 3) {
            if (L.sa) {
                streams.p[L.s_alpha][i] = blk[0];
                streams.p[L.s_a1][i] = blk[1];
            } else {
                streams.p[L.s_alpha][i * 2] = blk[0];
                streams.p[L.s_alpha][i * 2 + 1] = blk[1];
            }
            for (int k = 0; k < 6; k++) streams.p[L.s_aidx][i * 6 + k] = blk[2 + k];
        }
        uint32_t c = (uint32_t)blk[coff] | ((uint32_t)blk[coff + 1] << 8) | ((uint32_t)blk[coff + 2] << 16) |
                     ((uint32_t)blk[coff + 3] << 24);
        if (L.fmt == 1 && L.norm != kNormNone) {   // normalize the block before it is transformed
            uint32_t x = (uint32_t)blk[4] | ((uint32_t)blk[5] << 8) | ((uint32_t)blk[6] << 16) | ((uint32_t)blk[7] << 24);
            normalize_bc1_block(c, x, L.norm == kNormAllModesNone ? (int)kNormNone : L.norm);
            blk[4] = (uint8_t)x, blk[5] = (uint8_t)(x >> 8), blk[6] = (uint8_t)(x >> 16), blk[7] = (uint8_t)(x >> 24);
        }
        c = decorrelate_rt(c, L.var);
        uint8_t* p0 = L.sc ? streams.p[L.s_col] + i * 2 : streams.p[L.s_col] + i * 4;
        uint8_t* p1 = L.sc ? streams.p[L.s_c1] + i * 2 : p0 + 2;
        p0[0] = (uint8_t)c;
        p0[1] = (uint8_t)(c >> 8);
        p1[0] = (uint8_t)(c >> 16);
        p1[1] = (uint8_t)(c >> 24);
        for (int k = 0; k < 4; k++) streams.p[L.s_idx][i * 4 + k] = blk[coff + 4 + k];
    } else {
        if (L.fmt == 2) {
            for (int k = 0; k < 8; k++) blk[k] = streams.p[L.s_alpha][i * 8 + k];
        } else if (L.fmt == 3) {
            if (L.sa) {
                blk[0] = streams.p[L.s_alpha][i];
                blk[1] = streams.p[L.s_a1][i];
            } else {
                blk[0] = streams.p[L.s_alpha][i * 2];
                blk[1] = streams.p[L.s_alpha][i * 2 + 1];
            }
            for (int k = 0; k < 6; k++) blk[2 + k] = streams.p[L.s_aidx][i * 6 + k];
        }
        const uint8_t* p0 = L.sc ? streams.p[L.s_col] + i * 2 : streams.p[L.s_col] + i * 4;
        const uint8_t* p1 = L.sc ? streams.p[L.s_c1] + i * 2 : p0 + 2;
        uint32_t c = (uint32_t)p0[0] | ((uint32_t)p0[1] << 8) | ((uint32_t)p1[0] << 16) | ((uint32_t)p1[1] << 24);
        c = recorrelate_rt(c, L.var);
        blk[coff] = (uint8_t)c;
        blk[coff + 1] = (uint8_t)(c >> 8);
        blk[coff + 2] = (uint8_t)(c >> 16);
        blk[coff + 3] = (uint8_t)(c >> 24);
        for (int k = 0; k < 4; k++) blk[coff + 4 + k] = streams.p[L.s_idx][i * 4 + k];
        for (int k = 0; k < bpb; k++) blocks_out[i * bpb + k] = blk[k];
    }
}

// ------------------------------------------------------------------------------------------------
// split_color_endpoints: [c0 c1] x n  ->  c0 x n | c1 x n
// (common/src/transforms/split_565_color_endpoints/portable32.rs:18-64; the colour part of the BC1
// split layout on its own).  One 128-bit load = 4 pairs per thread, two coalesced 64-bit stores.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    split_endpoints_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ c0, uint8_t* __restrict__ c1,
                           const uint64_t npairs, const bool vector_ok) {
    const uint64_t q = (uint64_t)blockIdx.x * kThreads + threadIdx.x;  // group of 4 pairs
    const uint64_t first = q * 4;
    if (first >= npairs) return;
    if (vector_ok && first + 4 <= npairs) {
        const uint4 v = ldg_stream16(in + first * 4);
        stg_stream8(c0 + first * 2, make_uint2(__byte_perm(v.x, v.y, 0x5410), __byte_perm(v.z, v.w, 0x5410)));
        stg_stream8(c1 + first * 2, make_uint2(__byte_perm(v.x, v.y, 0x7632), __byte_perm(v.z, v.w, 0x7632)));
    } else {
        for (uint64_t i = first; i < first + 4 && i < npairs; i++) {
            c0[2 * i] = in[4 * i], c0[2 * i + 1] = in[4 * i + 1];
            c1[2 * i] = in[4 * i + 2], c1[2 * i + 1] = in[4 * i + 3];
        }
    }
}

template <int FMT, bool SA, bool SC>
RtLayout make_rt_layout(int var, int norm = kNormNone) {
    using L = Lay<FMT, SA, SC>;
    RtLayout r{};
    r.fmt = FMT;
    r.var = var;
    r.norm = norm;
    r.ns = L::NS;
    for (int s = 0; s < L::NS; s++) r.w[s] = L::w(s);
    r.s_alpha = L::sAlpha;
    r.s_a1 = L::sA1;
    r.s_aidx = L::sAIdx;
    r.s_col = L::sCol;
    r.s_c1 = L::sC1;
    r.s_idx = L::sIdx;
    r.sa = SA;
    r.sc = SC;
    return r;
}

// ------------------------------------------------------------------------------------------------
// Host dispatch
// ------------------------------------------------------------------------------------------------
template <int FMT, bool SA, bool SC>
bool tiled_ok(const void* blocks, const StreamPtrs& sp) {
    using L = Lay<FMT, SA, SC>;
    if (reinterpret_cast<uintptr_t>(blocks) & 15) return false;
    for (int s = 0; s < L::NS; s++)
        if (reinterpret_cast<uintptr_t>(sp.p[s]) & (uintptr_t)(stream_align(L::w(s)) - 1)) return false;
    return true;
}

template <int FMT, bool SA, bool SC, int VAR, bool RAGGED>
void launch_tiled_transform(int norm, unsigned tiles, const uint8_t* blocks_in, const StreamPtrs& sp, uint64_t n, cudaStream_t stream) {
    if constexpr (FMT == 1) {   // experimental normalization exists for BC1 only
        if (norm == kNormColor0Only) {
            transform_tiled<FMT, SA, SC, VAR, RAGGED, kNormColor0Only><<<tiles, kThreads, 0, stream>>>(blocks_in, sp, n);
            return;
        }
        if (norm == kNormReplicateColor) {
            transform_tiled<FMT, SA, SC, VAR, RAGGED, kNormReplicateColor><<<tiles, kThreads, 0, stream>>>(blocks_in, sp, n);
            return;
        }
        if (norm == kNormAllModesNone) {
            transform_tiled<FMT, SA, SC, VAR, RAGGED, kNormAllModesNone><<<tiles, kThreads, 0, stream>>>(blocks_in, sp, n);
            return;
        }
    }
    transform_tiled<FMT, SA, SC, VAR, RAGGED><<<tiles, kThreads, 0, stream>>>(blocks_in, sp, n);
}

template <int FMT, bool SA, bool SC, int VAR>
cudaError_t run(bool inverse, int norm, const uint8_t* blocks_in, uint8_t* blocks_out, const StreamPtrs& sp, uint64_t n,
                cudaStream_t stream) {
    using L = Lay<FMT, SA, SC>;
    if (n == 0) return cudaSuccess;
    const void* blocks = inverse ? (const void*)blocks_out : (const void*)blocks_in;
    if (tiled_ok<FMT, SA, SC>(blocks, sp)) {
        const uint64_t tiles = (n + L::T - 1) / L::T;
        if (tiles > 0x7fffffffull) return cudaErrorInvalidValue;
        bool ragged = false;
        for (int s = 0; s < L::NS; s++)
            ragged |= ((reinterpret_cast<uintptr_t>(sp.p[s]) | ((uint64_t)L::w(s) * n)) & 15) != 0;
        if (!inverse && ragged) launch_tiled_transform<FMT, SA, SC, VAR, true>(norm, (unsigned)tiles, blocks_in, sp, n, stream);
        else if (!inverse) launch_tiled_transform<FMT, SA, SC, VAR, false>(norm, (unsigned)tiles, blocks_in, sp, n, stream);
        else untransform_tiled<FMT, SA, SC, VAR><<<(unsigned)tiles, kThreads, 0, stream>>>(sp, blocks_out, n);
    } else {
        const uint64_t ctas = (n + kThreads - 1) / kThreads;
        if (ctas > 0x7fffffffull) return cudaErrorInvalidValue;
        bytewise_kernel<<<(unsigned)ctas, kThreads, 0, stream>>>(make_rt_layout<FMT, SA, SC>(VAR, inverse ? kNormNone : norm),
                                                                 inverse, blocks_in, blocks_out, sp, n);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <int FMT, bool SA, bool SC>
cudaError_t run_var(int var, int norm, bool inverse, const uint8_t* bi, uint8_t* bo, const StreamPtrs& sp, uint64_t n,
                    cudaStream_t stream) {
    switch (var) {
        case kNone: return run<FMT, SA, SC, kNone>(inverse, norm, bi, bo, sp, n, stream);
        case kVariant1: return run<FMT, SA, SC, kVariant1>(inverse, norm, bi, bo, sp, n, stream);
        case kVariant2: return run<FMT, SA, SC, kVariant2>(inverse, norm, bi, bo, sp, n, stream);
        case kVariant3: return run<FMT, SA, SC, kVariant3>(inverse, norm, bi, bo, sp, n, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t dispatch(const Settings& st, bool inverse, const uint8_t* bi, uint8_t* bo, const StreamPtrs& sp,
                     uint64_t n, cudaStream_t stream) {
    const bool sc = st.split_colour, sa = st.split_alpha;
    switch (st.format) {
        case 1:
            return sc ? run_var<1, false, true>(st.variant, st.normalize, inverse, bi, bo, sp, n, stream)
                      : run_var<1, false, false>(st.variant, st.normalize, inverse, bi, bo, sp, n, stream);
        case 2:
            return sc ? run_var<2, false, true>(st.variant, st.normalize, inverse, bi, bo, sp, n, stream)
                      : run_var<2, false, false>(st.variant, st.normalize, inverse, bi, bo, sp, n, stream);
        case 3:
            if (sa) {
                return sc ? run_var<3, true, true>(st.variant, st.normalize, inverse, bi, bo, sp, n, stream)
                          : run_var<3, true, false>(st.variant, st.normalize, inverse, bi, bo, sp, n, stream);
            }
            return sc ? run_var<3, false, true>(st.variant, st.normalize, inverse, bi, bo, sp, n, stream)
                      : run_var<3, false, false>(st.variant, st.normalize, inverse, bi, bo, sp, n, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <int FMT, bool SA, bool SC, int VAR>
cudaError_t run_batch(bool inverse, const void* d_items, int nitems, uint64_t max_blocks, bool ragged, cudaStream_t stream) {
    using L = Lay<FMT, SA, SC>;
    const uint64_t tiles = (max_blocks + L::T - 1) / L::T;
    if (tiles > 0x7fffffffull || nitems > 65535) return cudaErrorInvalidValue;
    const dim3 grid((unsigned)tiles, (unsigned)nitems);
    if (inverse) untransform_tiled_batch<FMT, SA, SC, VAR><<<grid, kThreads, 0, stream>>>(static_cast<const UntransformBatchItem*>(d_items));
    else if (ragged) transform_tiled_batch<FMT, SA, SC, VAR, true><<<grid, kThreads, 0, stream>>>(static_cast<const TransformBatchItem*>(d_items));
    else transform_tiled_batch<FMT, SA, SC, VAR, false><<<grid, kThreads, 0, stream>>>(static_cast<const TransformBatchItem*>(d_items));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}
template <int FMT, bool SA, bool SC>
cudaError_t run_batch_var(int var, bool inverse, const void* d, int n, uint64_t mb, bool ragged, cudaStream_t s) {
    switch (var) {
        case kNone: return run_batch<FMT, SA, SC, kNone>(inverse, d, n, mb, ragged, s);
        case kVariant1: return run_batch<FMT, SA, SC, kVariant1>(inverse, d, n, mb, ragged, s);
        case kVariant2: return run_batch<FMT, SA, SC, kVariant2>(inverse, d, n, mb, ragged, s);
        case kVariant3: return run_batch<FMT, SA, SC, kVariant3>(inverse, d, n, mb, ragged, s);
        default: return cudaErrorInvalidValue;
    }
}
cudaError_t dispatch_batch(const Settings& st, bool inverse, const void* d_items, int nitems, uint64_t max_blocks, bool ragged,
                           cudaStream_t stream) {
    if (nitems <= 0 || max_blocks == 0) return cudaSuccess;
    if (st.normalize != kNormNone) return cudaErrorInvalidValue;
    const bool sc = st.split_colour, sa = st.split_alpha;
    switch (st.format) {
        case 1:
            return sc ? run_batch_var<1, false, true>(st.variant, inverse, d_items, nitems, max_blocks, ragged, stream)
                      : run_batch_var<1, false, false>(st.variant, inverse, d_items, nitems, max_blocks, ragged, stream);
        case 2:
            return sc ? run_batch_var<2, false, true>(st.variant, inverse, d_items, nitems, max_blocks, ragged, stream)
                      : run_batch_var<2, false, false>(st.variant, inverse, d_items, nitems, max_blocks, ragged, stream);
        case 3:
            if (sa)
                return sc ? run_batch_var<3, true, true>(st.variant, inverse, d_items, nitems, max_blocks, ragged, stream)
                          : run_batch_var<3, true, false>(st.variant, inverse, d_items, nitems, max_blocks, ragged, stream);
            return sc ? run_batch_var<3, false, true>(st.variant, inverse, d_items, nitems, max_blocks, ragged, stream)
                      : run_batch_var<3, false, false>(st.variant, inverse, d_items, nitems, max_blocks, ragged, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

bool transform_batch_item_ok(const Settings& st, const TransformBatchItem& it, bool* ragged) {
    const int ns = num_streams(st.format, st.split_alpha, st.split_colour);
    if (reinterpret_cast<uintptr_t>(it.in) & 15) return false;
    for (int s = 0; s < ns; s++) {
        const int w = stream_width(st.format, st.split_alpha, st.split_colour, s);
        if (reinterpret_cast<uintptr_t>(it.out.p[s]) & (uintptr_t)(stream_align(w) - 1)) return false;
        *ragged |= ((reinterpret_cast<uintptr_t>(it.out.p[s]) | ((uint64_t)w * it.nblocks)) & 15) != 0;
    }
    return true;
}

cudaError_t launch_transform_batch(const Settings& st, const TransformBatchItem* d_items, int nitems, uint64_t max_blocks,
                                   bool ragged, cudaStream_t stream) {
    return dispatch_batch(st, false, d_items, nitems, max_blocks, ragged, stream);
}

bool untransform_batch_item_ok(const Settings& st, const UntransformBatchItem& it) {
    bool ragged = false;   // the untransform only reads the streams: any natural alignment goes
    return transform_batch_item_ok(st, TransformBatchItem{it.out, it.in, it.nblocks}, &ragged);
}

cudaError_t launch_untransform_batch(const Settings& st, const UntransformBatchItem* d_items, int nitems, uint64_t max_blocks,
                                     cudaStream_t stream) {
    return dispatch_batch(st, true, d_items, nitems, max_blocks, false, stream);
}

cudaError_t launch_transform(const Settings& st, const uint8_t* in, const StreamPtrs& out, uint64_t nblocks,
                             cudaStream_t stream) {
    return dispatch(st, false, in, nullptr, out, nblocks, stream);
}

cudaError_t launch_untransform(const Settings& st, const StreamPtrs& in, uint8_t* out, uint64_t nblocks,
                               cudaStream_t stream) {
    return dispatch(st, true, nullptr, out, in, nblocks, stream);
}

cudaError_t launch_split_color_endpoints(const uint8_t* in, uint8_t* out, uint64_t len_bytes, cudaStream_t stream) {
    const uint64_t npairs = len_bytes / 4;
    if (npairs == 0) return cudaSuccess;
    uint8_t* c1 = out + len_bytes / 2;
    const bool vector_ok = !(reinterpret_cast<uintptr_t>(in) & 15) && !(reinterpret_cast<uintptr_t>(out) & 7) &&
                           !(reinterpret_cast<uintptr_t>(c1) & 7);
    const uint64_t ctas = ((npairs + 3) / 4 + kThreads - 1) / kThreads;
    if (ctas > 0x7fffffffull) return cudaErrorInvalidValue;
    split_endpoints_kernel<<<(unsigned)ctas, kThreads, 0, stream>>>(in, out, c1, npairs, vector_ok);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// experimental::normalize_blocks as stand-alone passes (normalize.rs:38-101, 286-386, 417-484)
// ------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ uint2 load_block8(const uint8_t* p, bool aligned) {
    if (aligned) return *reinterpret_cast<const uint2*>(p);
    uint2 r;
    r.x = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    r.y = (uint32_t)p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24);
    return r;
}
__device__ __forceinline__ void store_block8(uint8_t* p, uint2 v, bool aligned) {
    if (aligned) {
        *reinterpret_cast<uint2*>(p) = v;
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) p[k] = (uint8_t)(v.x >> (8 * k)), p[4 + k] = (uint8_t)(v.y >> (8 * k));
}

// outs[m] (m = mode) may be null; `any` (optional) is set when a block is normalizable.
struct NormOuts {
    uint8_t* p[3];
};
// The three outputs of one block from ONE evaluation: the BlockCase does not depend on the mode.
__device__ __forceinline__ bool normalize_all_modes(const uint2 src, uint2 (&o)[3]) {
    uint2 c0only = src;
    const int bcase = normalize_bc1_block(c0only.x, c0only.y, kNormColor0Only);
    o[kNormNone] = bcase == 1 ? c0only : src;
    o[kNormColor0Only] = c0only;
    o[kNormReplicateColor] = bcase == 2 ? make_uint2(c0only.x | (c0only.x << 16), 0u) : c0only;
    return bcase != 0;
}
// Byte-aligned fallback: one block per thread.
__global__ void __launch_bounds__(kThreads)
    normalize_blocks_kernel(const uint8_t* in, const NormOuts outs, const uint64_t nblocks, const bool aligned, unsigned int* any) {
    const uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    bool hit = false;
    if (i < nblocks) {
        uint2 o[3];
        hit = normalize_all_modes(load_block8(in + 8 * i, aligned), o);
#pragma unroll
        for (int m = 0; m < 3; m++)
            if (outs.p[m]) store_block8(outs.p[m] + 8 * i, o[m], aligned);
    }
    if (any && __syncthreads_or(hit) && threadIdx.x == 0) atomicOr(any, 1u);
}
// 16-byte aligned pointers: a CTA owns a 16 KiB tile, every thread keeps four 128-bit loads in flight (plain loads, not
// .nc: an output may be the input itself) and writes 128-bit vectors — the shape of the transform kernels.
__global__ void __launch_bounds__(kThreads)
    normalize_blocks_vec_kernel(const uint8_t* in, const NormOuts outs, const uint64_t nblocks, unsigned int* any) {
    const uint64_t vec0 = (uint64_t)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
    uint4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
        const uint64_t j = vec0 + (uint64_t)u * kThreads;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (2 * j + 1 < nblocks) {
            asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                         : "l"(in + 16 * j)
                         : "memory");
        } else if (2 * j < nblocks) {
            asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v[u].x), "=r"(v[u].y) : "l"(in + 16 * j) : "memory");
        }
    }
    __shared__ uint2 queue[kNormQueueEntries * (kThreads / 32)];
    uint32_t res[2 * kUnroll];
    classify8_warp(v, res, queue + (threadIdx.x >> 5) * kNormQueueEntries);   // blocks past the end are zeros: a fast case
    bool hit = false;
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
        const uint64_t j = vec0 + (uint64_t)u * kThreads;
        if (2 * j >= nblocks) continue;
        const bool two = 2 * j + 1 < nblocks;
        hit |= (res[2 * u] & 3u) != 0 || (two && (res[2 * u + 1] & 3u) != 0);
#pragma unroll
        for (int m = 0; m < 3; m++) {
            if (!outs.p[m]) continue;
            uint4 o = v[u];
            apply_normalization(o.x, o.y, res[2 * u], m);
            apply_normalization(o.z, o.w, res[2 * u + 1], m);
            if (two) stg_stream16(outs.p[m] + 16 * j, o);
            else stg_stream8(outs.p[m] + 16 * j, make_uint2(o.x, o.y));
        }
    }
    if (any && __syncthreads_or(hit) && threadIdx.x == 0) atomicOr(any, 1u);
}

// normalize_split_blocks_in_place: colours (c0 c1 pairs) and indices live in separate arrays.
__global__ void __launch_bounds__(kThreads)
    normalize_split_kernel(uint8_t* colors, uint8_t* indices, const uint64_t nblocks, const int mode, const bool aligned) {
    const uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= nblocks) return;
    uint32_t c, x;
    if (aligned) {
        c = *reinterpret_cast<const uint32_t*>(colors + 4 * i), x = *reinterpret_cast<const uint32_t*>(indices + 4 * i);
    } else {
        c = x = 0;
        for (int k = 0; k < 4; k++) c |= (uint32_t)colors[4 * i + k] << (8 * k), x |= (uint32_t)indices[4 * i + k] << (8 * k);
    }
    if (normalize_bc1_block(c, x, mode) == 0) return;   // in place: nothing to write
    if (aligned) {
        *reinterpret_cast<uint32_t*>(colors + 4 * i) = c, *reinterpret_cast<uint32_t*>(indices + 4 * i) = x;
    } else {
        for (int k = 0; k < 4; k++) colors[4 * i + k] = (uint8_t)(c >> (8 * k)), indices[4 * i + k] = (uint8_t)(x >> (8 * k));
    }
}

}  // namespace

cudaError_t launch_normalize_blocks(const uint8_t* in, uint8_t* out_none, uint8_t* out_color0, uint8_t* out_replicate,
                                    uint64_t nblocks, unsigned int* d_any, cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    const uint64_t ctas = (nblocks + kThreads - 1) / kThreads;
    if (ctas > 0x7fffffffull) return cudaErrorInvalidValue;
    const NormOuts outs{{out_none, out_color0, out_replicate}};
    const uintptr_t bits = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out_none) |
                           reinterpret_cast<uintptr_t>(out_color0) | reinterpret_cast<uintptr_t>(out_replicate);
    if ((bits & 15) == 0) {
        const uint64_t per_cta = 2ull * kThreads * kUnroll;
        const uint64_t vctas = (nblocks + per_cta - 1) / per_cta;
        normalize_blocks_vec_kernel<<<(unsigned)vctas, kThreads, 0, stream>>>(in, outs, nblocks, d_any);
    } else {
        normalize_blocks_kernel<<<(unsigned)ctas, kThreads, 0, stream>>>(in, outs, nblocks, (bits & 7) == 0, d_any);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_normalize_split_blocks(uint8_t* colors, uint8_t* indices, uint64_t nblocks, int mode, cudaStream_t stream) {
    if (nblocks == 0 || mode == kNormNone) return cudaSuccess;
    const uint64_t ctas = (nblocks + kThreads - 1) / kThreads;
    if (ctas > 0x7fffffffull) return cudaErrorInvalidValue;
    const bool aligned = ((reinterpret_cast<uintptr_t>(colors) | reinterpret_cast<uintptr_t>(indices)) & 3) == 0;
    normalize_split_kernel<<<(unsigned)ctas, kThreads, 0, stream>>>(colors, indices, nblocks, mode, aligned);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// Gather: many small host->device (or any device-visible) copies as ONE launch — grid.y = copy, 8-byte vectors (BCn
// payloads are multiples of 8 bytes; the pointers must be 8-byte aligned).
namespace {
__global__ void __launch_bounds__(kThreads) copy_batch_kernel(const CopyBatchItem* __restrict__ items) {
    const CopyBatchItem it = items[blockIdx.y];
    const uint64_t first = (uint64_t)blockIdx.x * kThreads + threadIdx.x, step = (uint64_t)gridDim.x * kThreads;
    if (((reinterpret_cast<uintptr_t>(it.src) | reinterpret_cast<uintptr_t>(it.dst)) & 7) == 0) {
        const uint2* src = reinterpret_cast<const uint2*>(it.src);
        uint2* dst = reinterpret_cast<uint2*>(it.dst);
        for (uint64_t i = first; i < it.bytes / 8; i += step) stg_stream8(dst + i, ldg_stream8(src + i));
    } else if (reinterpret_cast<uintptr_t>(it.src) & 15) {
        // Gather from a 4-byte aligned source (a payload behind a 148-byte header, in mapped host memory): the host side is
        // read as whole aligned 128-bit vectors (4-byte reads over the host link crawl), the device side takes words.
        const uintptr_t shift = reinterpret_cast<uintptr_t>(it.src) & 15;
        const uint8_t* a0 = it.src - shift;
        uint32_t* dst = reinterpret_cast<uint32_t*>(it.dst);
        const uint64_t nvec = (shift + it.bytes + 15) / 16;
        for (uint64_t t = first; t < nvec; t += step) {
            const uint4 v = ldg_stream16(a0 + 16 * t);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int64_t pos = (int64_t)(16 * t + 4 * k) - (int64_t)shift;
                if (pos >= 0 && (uint64_t)pos < it.bytes) dst[pos / 4] = w[k];
            }
        }
    } else {
        // Scatter to a 4-byte aligned destination: whole aligned 128-bit vectors on the host side, words on the device side.
        const uintptr_t shift = reinterpret_cast<uintptr_t>(it.dst) & 15;
        uint8_t* a0 = it.dst - shift;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(it.src);
        const uint64_t nvec = (shift + it.bytes + 15) / 16;
        for (uint64_t t = first; t < nvec; t += step) {
            uint32_t w[4];
            bool ok[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int64_t pos = (int64_t)(16 * t + 4 * k) - (int64_t)shift;
                ok[k] = pos >= 0 && (uint64_t)pos < it.bytes;
                w[k] = ok[k] ? src[pos / 4] : 0u;
            }
            if (ok[0] && ok[3]) {
                stg_stream16(a0 + 16 * t, make_uint4(w[0], w[1], w[2], w[3]));
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (ok[k]) reinterpret_cast<uint32_t*>(a0 + 16 * t)[k] = w[k];
            }
        }
    }
}
}  // namespace

cudaError_t launch_copy_batch(const CopyBatchItem* d_items, int nitems, uint64_t max_bytes, cudaStream_t stream) {
    if (nitems <= 0 || max_bytes == 0) return cudaSuccess;
    if (nitems > 65535) return cudaErrorInvalidValue;
    const uint64_t ctas = std::min<uint64_t>((max_bytes / 4 + kThreads - 1) / kThreads, 1024);
    copy_batch_kernel<<<dim3((unsigned)ctas, (unsigned)nitems), kThreads, 0, stream>>>(d_items);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Endpoint streams of several candidates from one read (see bcn_kernels.h)
// ------------------------------------------------------------------------------------------------
namespace {

struct EndpointCandidateSet {
    EndpointCandidate c[kMaxEndpointCandidates];
    int n;
    unsigned variant_mask;   // which YCoCg variants occur
};

// eight elements of `W` bytes each, held in the low bits of v[0..8), to dst (element-aligned); one or two 128-bit stores
// when dst is 16-byte aligned, element stores otherwise (odd block counts put the second stream of a split pair anywhere)
template <int W>
__device__ __forceinline__ void store8(uint8_t* dst, const uint32_t (&v)[8]) {
    if constexpr (W == 4) {
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            reinterpret_cast<uint4*>(dst)[0] = make_uint4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<uint4*>(dst)[1] = make_uint4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) reinterpret_cast<uint32_t*>(dst)[j] = v[j];
        }
    } else if constexpr (W == 2) {
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(__byte_perm(v[0], v[1], 0x5410), __byte_perm(v[2], v[3], 0x5410),
                                                        __byte_perm(v[4], v[5], 0x5410), __byte_perm(v[6], v[7], 0x5410));
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) reinterpret_cast<uint16_t*>(dst)[j] = (uint16_t)v[j];
        }
    } else {
        if ((reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
            *reinterpret_cast<uint2*>(dst) = make_uint2((v[0] & 0xFF) | (v[1] & 0xFF) << 8 | (v[2] & 0xFF) << 16 | v[3] << 24,
                                                        (v[4] & 0xFF) | (v[5] & 0xFF) << 8 | (v[6] & 0xFF) << 16 | v[7] << 24);
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) dst[j] = (uint8_t)v[j];
        }
    }
}

template <int FMT>
__global__ void __launch_bounds__(256) endpoint_candidates_kernel(const uint8_t* __restrict__ blocks, const uint64_t nblocks,
                                                                   const EndpointCandidateSet set) {
    const uint64_t g = (uint64_t)blockIdx.x * 256 + threadIdx.x;   // group of eight blocks
    const uint64_t b0 = g * 8;
    if (b0 >= nblocks) return;
    const int valid = nblocks - b0 >= 8 ? 8 : (int)(nblocks - b0);
    uint32_t cw[8], aw[8];   // colour words (c0 | c1 << 16), BC3 alpha endpoints (a0 | a1 << 8)
    if constexpr (FMT == 1) {
        const uint4* p = reinterpret_cast<const uint4*>(blocks) + g * 4;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (2 * j < valid) {
                if (2 * j + 1 < valid) v = ldg_stream16(p + j);
                else v.x = __ldg(reinterpret_cast<const uint32_t*>(p + j));   // the last, odd block: 8 bytes exist
            }
            cw[2 * j] = v.x, cw[2 * j + 1] = v.z;
            aw[2 * j] = aw[2 * j + 1] = 0;
        }
    } else {
        const uint4* p = reinterpret_cast<const uint4*>(blocks) + b0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint4 v = j < valid ? ldg_stream16(p + j) : make_uint4(0, 0, 0, 0);
            cw[j] = v.z, aw[j] = v.x & 0xFFFFu;
        }
    }
#pragma unroll
    for (int var = 0; var < 4; var++) {
        if (!(set.variant_mask & (1u << var))) continue;
        uint32_t d[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
            d[j] = var == kNone ? cw[j] : var == kVariant1 ? decorrelate2<kVariant1>(cw[j]) : var == kVariant2 ? decorrelate2<kVariant2>(cw[j])
                                                                                                          : decorrelate2<kVariant3>(cw[j]);
#pragma unroll
        for (int k = 0; k < kMaxEndpointCandidates; k++) {   // (constant indices: the set stays in the parameter space)
            if (k >= set.n) break;
            const EndpointCandidate c = set.c[k];
            if (c.variant != var || c.colour == nullptr) continue;
            if (valid == 8) {
                if (!c.split_colour) {
                    store8<4>(c.colour + 4 * b0, d);
                } else {
                    uint32_t hi[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) hi[j] = d[j] >> 16;
                    store8<2>(c.colour + 2 * b0, d);
                    store8<2>(c.colour + 2 * nblocks + 2 * b0, hi);
                }
            } else {
                for (int j = 0; j < valid; j++) {
                    if (!c.split_colour) {
                        reinterpret_cast<uint32_t*>(c.colour)[b0 + j] = d[j];
                    } else {
                        reinterpret_cast<uint16_t*>(c.colour)[b0 + j] = (uint16_t)d[j];
                        reinterpret_cast<uint16_t*>(c.colour + 2 * nblocks)[b0 + j] = (uint16_t)(d[j] >> 16);
                    }
                }
            }
        }
    }
    if constexpr (FMT == 3) {
#pragma unroll
        for (int k = 0; k < kMaxEndpointCandidates; k++) {
            if (k >= set.n) break;
            const EndpointCandidate c = set.c[k];
            if (c.alpha == nullptr) continue;
            if (valid == 8) {
                if (!c.split_alpha) {
                    store8<2>(c.alpha + 2 * b0, aw);
                } else {
                    uint32_t hi[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) hi[j] = aw[j] >> 8;
                    store8<1>(c.alpha + b0, aw);
                    store8<1>(c.alpha + nblocks + b0, hi);
                }
            } else {
                for (int j = 0; j < valid; j++) {
                    if (!c.split_alpha) {
                        reinterpret_cast<uint16_t*>(c.alpha)[b0 + j] = (uint16_t)aw[j];
                    } else {
                        c.alpha[b0 + j] = (uint8_t)aw[j];
                        c.alpha[nblocks + b0 + j] = (uint8_t)(aw[j] >> 8);
                    }
                }
            }
        }
    }
}

}  // namespace

cudaError_t launch_endpoint_candidates(int format, const uint8_t* blocks, uint64_t nblocks, const EndpointCandidate* cands, int count,
                                       cudaStream_t stream) {
    if (nblocks == 0 || count <= 0) return cudaSuccess;
    if (count > kMaxEndpointCandidates || (reinterpret_cast<uintptr_t>(blocks) & 15)) return cudaErrorInvalidValue;
    EndpointCandidateSet set{};
    set.n = count;
    for (int k = 0; k < count; k++) {
        set.c[k] = cands[k];
        if (cands[k].colour) set.variant_mask |= 1u << cands[k].variant;
    }
    const uint64_t groups = (nblocks + 7) / 8;
    const unsigned grid = (unsigned)((groups + 255) / 256);
    if (format == 1) endpoint_candidates_kernel<1><<<grid, 256, 0, stream>>>(blocks, nblocks, set);
    else if (format == 2) endpoint_candidates_kernel<2><<<grid, 256, 0, stream>>>(blocks, nblocks, set);
    else endpoint_candidates_kernel<3><<<grid, 256, 0, stream>>>(blocks, nblocks, set);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

uint64_t kernel_launch_count() { return g_launches.load(std::memory_order_relaxed); }

}  // namespace dlt
