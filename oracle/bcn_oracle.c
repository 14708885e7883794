/*
 * bcn_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see bcn_oracle.h).
 *
 * Scalar, little-endian restatement of the reference's ground-truth code paths.
 * Reference paths are relative to /root/reference/src.  Every routine works on a
 * block range [b0, b1) of a payload of n blocks so the multi-threaded CPU baseline
 * and the single call share one body; section offsets are always functions of the
 * WHOLE payload's n, exactly as in the reference dispatchers.
 */
#include "bcn_oracle.h"
#include "ltu_params.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Color565 YCoCg-R   (core/dxt-lossless-transform-common/src/color_565/decorrelate.rs)
 * ---------------------------------------------------------------------------------------- */

/* decorrelate_ycocg_r_var1/2/3: decorrelate.rs:101-127 / 187-214 / 274-300.  The arithmetic is
 * shared; only the final packing differs per variant. */
uint16_t orc_decorrelate(uint16_t v, int variant) {
    if (variant == ORC_VARIANT_NONE) return v;
    uint32_t r = (v >> 11) & 0x1F, g = (v >> 6) & 0x1F, gl = (v >> 5) & 1, b = v & 0x1F;
    uint32_t co = (r - b) & 0x1F;
    uint32_t t = (b + (co >> 1)) & 0x1F;
    uint32_t cg = (g - t) & 0x1F;
    uint32_t y = (t + (cg >> 1)) & 0x1F;
    switch (variant) {
    case ORC_VARIANT_1: return (uint16_t)((y << 11) | (co << 6) | (gl << 5) | cg);
    case ORC_VARIANT_2: return (uint16_t)((gl << 15) | (y << 10) | (co << 5) | cg);
    default: return (uint16_t)((y << 11) | (co << 6) | (cg << 1) | gl);
    }
}

/* recorrelate_ycocg_r_var1/2/3: decorrelate.rs:148-171 / 235-258 / 321-344. */
uint16_t orc_recorrelate(uint16_t v, int variant) {
    uint32_t y, co, cg, gl;
    switch (variant) {
    case ORC_VARIANT_NONE: return v;
    case ORC_VARIANT_1: y = (v >> 11) & 0x1F; co = (v >> 6) & 0x1F; gl = (v >> 5) & 1; cg = v & 0x1F; break;
    case ORC_VARIANT_2: gl = v >> 15; y = (v >> 10) & 0x1F; co = (v >> 5) & 0x1F; cg = v & 0x1F; break;
    default: y = (v >> 11) & 0x1F; co = (v >> 6) & 0x1F; cg = (v >> 1) & 0x1F; gl = v & 1; break;
    }
    uint32_t t = (y - (cg >> 1)) & 0x1F;
    uint32_t g = (cg + t) & 0x1F;
    uint32_t b = (t - (co >> 1)) & 0x1F;
    uint32_t r = (b + co) & 0x1F;
    return (uint16_t)((r << 11) | (g << 6) | (gl << 5) | b);
}

static inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline void wr16(uint8_t *p, uint16_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

/* Colour part shared by BC1/BC2/BC3: `col` points at [c0:u16][c1:u16] of one block.
 * no split → (c0,c1) pair at colours + 4*i; split → c0 at c0s + 2*i, c1 at c1s + 2*i.
 * BC1 with_split_colour_and_recorr/transform/generic.rs:39-83, with_recorrelate/transform/generic.rs:38-83. */
static inline void colour_fwd(const uint8_t *col, uint8_t *c0dst, uint8_t *c1dst, int variant) {
    wr16(c0dst, orc_decorrelate(rd16(col), variant));
    wr16(c1dst, orc_decorrelate(rd16(col + 2), variant));
}
static inline void colour_inv(const uint8_t *c0src, const uint8_t *c1src, uint8_t *col, int variant) {
    wr16(col, orc_recorrelate(rd16(c0src), variant));
    wr16(col + 2, orc_recorrelate(rd16(c1src), variant));
}

/* ------------------------------------------------------------------------------------------
 * BC1   (core/dxt-lossless-transform-bc1/src/transform/transform_with_settings.rs:31-135)
 *   no split: colours[4n] @0 | indices[4n] @len/2      (standard/transform/portable32.rs:8-47)
 *   split   : c0[2n] @0 | c1[2n] @len/4 | indices @len/2
 * ---------------------------------------------------------------------------------------- */
static void bc1_range(int inverse, const uint8_t *in, uint8_t *out, size_t n, size_t b0, size_t b1,
                      int variant, int split) {
    size_t len = n * 8;
    for (size_t i = b0; i < b1; i++) {
        size_t c0o = split ? 2 * i : 4 * i;
        size_t c1o = split ? len / 4 + 2 * i : 4 * i + 2;
        size_t ixo = len / 2 + 4 * i;
        if (!inverse) {
            const uint8_t *blk = in + 8 * i;
            colour_fwd(blk, out + c0o, out + c1o, variant);
            memcpy(out + ixo, blk + 4, 4);
        } else {
            uint8_t *blk = out + 8 * i;
            colour_inv(in + c0o, in + c1o, blk, variant);
            memcpy(blk + 4, in + ixo, 4);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * BC2   (core/dxt-lossless-transform-bc2/src/transform/transform_with_settings.rs:30-137)
 *   alpha[8n] @0 | colours[4n] @len/2 (split: c0 @len/2, c1 @len/2+len/8) | indices @len/2+len/4
 *   (standard/transform/portable32.rs:6-56)
 * ---------------------------------------------------------------------------------------- */
static void bc2_range(int inverse, const uint8_t *in, uint8_t *out, size_t n, size_t b0, size_t b1,
                      int variant, int split) {
    size_t len = n * 16;
    for (size_t i = b0; i < b1; i++) {
        size_t alo = 8 * i;
        size_t c0o = len / 2 + (split ? 2 * i : 4 * i);
        size_t c1o = split ? len / 2 + len / 8 + 2 * i : len / 2 + 4 * i + 2;
        size_t ixo = len / 2 + len / 4 + 4 * i;
        if (!inverse) {
            const uint8_t *blk = in + 16 * i;
            memcpy(out + alo, blk, 8);
            colour_fwd(blk + 8, out + c0o, out + c1o, variant);
            memcpy(out + ixo, blk + 12, 4);
        } else {
            uint8_t *blk = out + 16 * i;
            memcpy(blk, in + alo, 8);
            colour_inv(in + c0o, in + c1o, blk + 8, variant);
            memcpy(blk + 12, in + ixo, 4);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * BC3   (core/dxt-lossless-transform-bc3/src/transform/transform_with_settings.rs:32-290)
 *   alpha endpoints [0,2n): pairs, or split a0[n] @0 | a1[n] @n
 *   alpha indices   [2n,8n): 6 B/block verbatim
 *   colours         [8n,12n): pairs, or split c0[2n] @8n | c1[2n] @10n
 *   colour indices  [12n,16n)
 *   (standard/transform/portable32.rs:10-65; with_split_alphas_colour_and_recorr/transform/generic.rs:24-88)
 * ---------------------------------------------------------------------------------------- */
static void bc3_range(int inverse, const uint8_t *in, uint8_t *out, size_t n, size_t b0, size_t b1,
                      int variant, int split_a, int split_c) {
    for (size_t i = b0; i < b1; i++) {
        size_t a0o = split_a ? i : 2 * i;
        size_t a1o = split_a ? n + i : 2 * i + 1;
        size_t aio = 2 * n + 6 * i;
        size_t c0o = 8 * n + (split_c ? 2 * i : 4 * i);
        size_t c1o = split_c ? 10 * n + 2 * i : 8 * n + 4 * i + 2;
        size_t ixo = 12 * n + 4 * i;
        if (!inverse) {
            const uint8_t *blk = in + 16 * i;
            out[a0o] = blk[0];
            out[a1o] = blk[1];
            memcpy(out + aio, blk + 2, 6);
            colour_fwd(blk + 8, out + c0o, out + c1o, variant);
            memcpy(out + ixo, blk + 12, 4);
        } else {
            uint8_t *blk = out + 16 * i;
            blk[0] = in[a0o];
            blk[1] = in[a1o];
            memcpy(blk + 2, in + aio, 6);
            colour_inv(in + c0o, in + c1o, blk + 8, variant);
            memcpy(blk + 12, in + ixo, 4);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Word-wise BC1 path for the CPU BASELINE timing (bench.py): the same bytes as bc1_range, but
 * written the way the reference's SIMD tiers work (two colours per 32-bit lane, SWAR YCoCg-R as in
 * common/src/intrinsics/color_565/decorrelate/avx512bw.rs:42-96) so that gcc -O3 auto-vectorises
 * it.  tests/test_oracle.py checks it against bc1_range for every settings combination.
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t decorr2(uint32_t v, int variant) {
    const uint32_t M5 = 0x001F001Fu, M4 = 0x000F000Fu, M1 = 0x00010001u, B = 0x00200020u;
    uint32_t r = (v >> 11) & M5, g = (v >> 6) & M5, gl = (v >> 5) & M1, b = v & M5;
    uint32_t co = (r + B - b) & M5;
    uint32_t t = (b + ((co >> 1) & M4)) & M5;
    uint32_t cg = (g + B - t) & M5;
    uint32_t y = (t + ((cg >> 1) & M4)) & M5;
    if (variant == ORC_VARIANT_1) return (y << 11) | (co << 6) | (gl << 5) | cg;
    if (variant == ORC_VARIANT_2) return (gl << 15) | (y << 10) | (co << 5) | cg;
    return (y << 11) | (co << 6) | (cg << 1) | gl;
}
static inline uint32_t recorr2(uint32_t v, int variant) {
    const uint32_t M5 = 0x001F001Fu, M4 = 0x000F000Fu, M1 = 0x00010001u, B = 0x00200020u;
    uint32_t y, co, cg, gl;
    if (variant == ORC_VARIANT_1) { y = (v >> 11) & M5; co = (v >> 6) & M5; gl = (v >> 5) & M1; cg = v & M5; }
    else if (variant == ORC_VARIANT_2) { gl = (v >> 15) & M1; y = (v >> 10) & M5; co = (v >> 5) & M5; cg = v & M5; }
    else { y = (v >> 11) & M5; co = (v >> 6) & M5; cg = (v >> 1) & M5; gl = v & M1; }
    uint32_t t = (y + B - ((cg >> 1) & M4)) & M5;
    uint32_t g = (cg + t) & M5;
    uint32_t b = (t + B - ((co >> 1) & M4)) & M5;
    uint32_t r = (b + co) & M5;
    return (r << 11) | (g << 6) | (gl << 5) | b;
}

#define BC1_FAST_LOOP(VARIANT)                                                                      \
    for (size_t i = b0; i < b1; i++) {                                                              \
        uint32_t c, x;                                                                              \
        if (!inverse) {                                                                             \
            memcpy(&c, in + 8 * i, 4);                                                              \
            memcpy(&x, in + 8 * i + 4, 4);                                                          \
            if (VARIANT) c = decorr2(c, VARIANT);                                                   \
            if (split) {                                                                            \
                uint16_t lo = (uint16_t)c, hi = (uint16_t)(c >> 16);                                \
                memcpy(out + 2 * i, &lo, 2);                                                        \
                memcpy(out + len / 4 + 2 * i, &hi, 2);                                              \
            } else memcpy(out + 4 * i, &c, 4);                                                      \
            memcpy(out + len / 2 + 4 * i, &x, 4);                                                   \
        } else {                                                                                    \
            if (split) {                                                                            \
                uint16_t lo, hi;                                                                    \
                memcpy(&lo, in + 2 * i, 2);                                                         \
                memcpy(&hi, in + len / 4 + 2 * i, 2);                                               \
                c = (uint32_t)lo | ((uint32_t)hi << 16);                                            \
            } else memcpy(&c, in + 4 * i, 4);                                                       \
            memcpy(&x, in + len / 2 + 4 * i, 4);                                                    \
            if (VARIANT) c = recorr2(c, VARIANT);                                                   \
            memcpy(out + 8 * i, &c, 4);                                                             \
            memcpy(out + 8 * i + 4, &x, 4);                                                         \
        }                                                                                           \
    }

static void bc1_range_fast(int inverse, const uint8_t *restrict in, uint8_t *restrict out, size_t n, size_t b0,
                           size_t b1, int variant, int split) {
    const size_t len = n * 8;
    switch (variant) {
    case ORC_VARIANT_NONE: BC1_FAST_LOOP(0) break;
    case ORC_VARIANT_1: BC1_FAST_LOOP(1) break;
    case ORC_VARIANT_2: BC1_FAST_LOOP(2) break;
    default: BC1_FAST_LOOP(3) break;
    }
}

/* ------------------------------------------------------------------------------------------
 * Explicit AVX2 BC1 path for the CPU BASELINE timing: eight blocks per iteration, colours gathered with dword
 * shuffles, YCoCg-R on sixteen 16-bit lanes, endpoints split with a byte shuffle — the shape of the reference's
 * AVX2 tier (core/dxt-lossless-transform-bc1/src/transform/standard/transform/avx2.rs and
 * common/src/intrinsics/color_565/decorrelate/avx2.rs), restated.  Used by orc_bcn_run_mt / orc_bcn_run_range when
 * the host CPU has AVX2; tests/test_oracle.py checks it against bc1_range for every settings combination.
 * ---------------------------------------------------------------------------------------- */
#if defined(__x86_64__)
#include <immintrin.h>
#define ORC_AVX2 __attribute__((target("avx2")))

ORC_AVX2 static inline __m256i decorr16(__m256i v, int variant) {
    const __m256i m5 = _mm256_set1_epi16(31), m1 = _mm256_set1_epi16(1);
    __m256i r = _mm256_srli_epi16(v, 11), g = _mm256_and_si256(_mm256_srli_epi16(v, 6), m5);
    __m256i gl = _mm256_and_si256(_mm256_srli_epi16(v, 5), m1), b = _mm256_and_si256(v, m5);
    __m256i co = _mm256_and_si256(_mm256_sub_epi16(r, b), m5);
    __m256i t = _mm256_and_si256(_mm256_add_epi16(b, _mm256_srli_epi16(co, 1)), m5);
    __m256i cg = _mm256_and_si256(_mm256_sub_epi16(g, t), m5);
    __m256i y = _mm256_and_si256(_mm256_add_epi16(t, _mm256_srli_epi16(cg, 1)), m5);
    if (variant == ORC_VARIANT_1)
        return _mm256_or_si256(_mm256_or_si256(_mm256_slli_epi16(y, 11), _mm256_slli_epi16(co, 6)), _mm256_or_si256(_mm256_slli_epi16(gl, 5), cg));
    if (variant == ORC_VARIANT_2)
        return _mm256_or_si256(_mm256_or_si256(_mm256_slli_epi16(gl, 15), _mm256_slli_epi16(y, 10)), _mm256_or_si256(_mm256_slli_epi16(co, 5), cg));
    return _mm256_or_si256(_mm256_or_si256(_mm256_slli_epi16(y, 11), _mm256_slli_epi16(co, 6)), _mm256_or_si256(_mm256_slli_epi16(cg, 1), gl));
}
ORC_AVX2 static inline __m256i recorr16(__m256i v, int variant) {
    const __m256i m5 = _mm256_set1_epi16(31), m1 = _mm256_set1_epi16(1);
    __m256i y, co, cg, gl;
    if (variant == ORC_VARIANT_1) {
        y = _mm256_srli_epi16(v, 11), co = _mm256_and_si256(_mm256_srli_epi16(v, 6), m5);
        gl = _mm256_and_si256(_mm256_srli_epi16(v, 5), m1), cg = _mm256_and_si256(v, m5);
    } else if (variant == ORC_VARIANT_2) {
        gl = _mm256_srli_epi16(v, 15), y = _mm256_and_si256(_mm256_srli_epi16(v, 10), m5);
        co = _mm256_and_si256(_mm256_srli_epi16(v, 5), m5), cg = _mm256_and_si256(v, m5);
    } else {
        y = _mm256_srli_epi16(v, 11), co = _mm256_and_si256(_mm256_srli_epi16(v, 6), m5);
        cg = _mm256_and_si256(_mm256_srli_epi16(v, 1), m5), gl = _mm256_and_si256(v, m1);
    }
    __m256i t = _mm256_and_si256(_mm256_sub_epi16(y, _mm256_srli_epi16(cg, 1)), m5);
    __m256i g = _mm256_and_si256(_mm256_add_epi16(cg, t), m5);
    __m256i b = _mm256_and_si256(_mm256_sub_epi16(t, _mm256_srli_epi16(co, 1)), m5);
    __m256i r = _mm256_and_si256(_mm256_add_epi16(b, co), m5);
    return _mm256_or_si256(_mm256_or_si256(_mm256_slli_epi16(r, 11), _mm256_slli_epi16(g, 6)), _mm256_or_si256(_mm256_slli_epi16(gl, 5), b));
}

#define BC1_AVX2_LOOP(VARIANT)                                                                                       \
    for (; i + 8 <= b1; i += 8) {                                                                                    \
        if (!inverse) {                                                                                              \
            const __m256 a = _mm256_castsi256_ps(_mm256_loadu_si256((const __m256i *)(in + 8 * i)));                 \
            const __m256 b = _mm256_castsi256_ps(_mm256_loadu_si256((const __m256i *)(in + 8 * i + 32)));            \
            __m256i col = _mm256_permute4x64_epi64(_mm256_castps_si256(_mm256_shuffle_ps(a, b, 0x88)), 0xD8);        \
            const __m256i idx = _mm256_permute4x64_epi64(_mm256_castps_si256(_mm256_shuffle_ps(a, b, 0xDD)), 0xD8);  \
            if (VARIANT) col = decorr16(col, VARIANT);                                                               \
            if (split) {                                                                                             \
                const __m256i s = _mm256_permute4x64_epi64(_mm256_shuffle_epi8(col, split_mask), 0xD8);              \
                _mm_storeu_si128((__m128i *)(out + 2 * i), _mm256_castsi256_si128(s));                               \
                _mm_storeu_si128((__m128i *)(out + len / 4 + 2 * i), _mm256_extracti128_si256(s, 1));                \
            } else _mm256_storeu_si256((__m256i *)(out + 4 * i), col);                                               \
            _mm256_storeu_si256((__m256i *)(out + len / 2 + 4 * i), idx);                                            \
        } else {                                                                                                     \
            __m256i col;                                                                                             \
            if (split) {                                                                                             \
                const __m256i c0 = _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i *)(in + 2 * i)));            \
                const __m256i c1 = _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i *)(in + len / 4 + 2 * i)));  \
                col = _mm256_or_si256(c0, _mm256_slli_epi32(c1, 16));                                                \
            } else col = _mm256_loadu_si256((const __m256i *)(in + 4 * i));                                          \
            const __m256i idx = _mm256_loadu_si256((const __m256i *)(in + len / 2 + 4 * i));                         \
            if (VARIANT) col = recorr16(col, VARIANT);                                                               \
            const __m256i lo = _mm256_unpacklo_epi32(col, idx), hi = _mm256_unpackhi_epi32(col, idx);                \
            _mm256_storeu_si256((__m256i *)(out + 8 * i), _mm256_permute2x128_si256(lo, hi, 0x20));                  \
            _mm256_storeu_si256((__m256i *)(out + 8 * i + 32), _mm256_permute2x128_si256(lo, hi, 0x31));             \
        }                                                                                                            \
    }

ORC_AVX2 static void bc1_range_avx2(int inverse, const uint8_t *restrict in, uint8_t *restrict out, size_t n, size_t b0,
                                    size_t b1, int variant, int split) {
    const size_t len = n * 8;
    const __m256i split_mask = _mm256_setr_epi8(0, 1, 4, 5, 8, 9, 12, 13, 2, 3, 6, 7, 10, 11, 14, 15,
                                                0, 1, 4, 5, 8, 9, 12, 13, 2, 3, 6, 7, 10, 11, 14, 15);
    size_t i = b0;
    switch (variant) {
    case ORC_VARIANT_NONE: BC1_AVX2_LOOP(0) break;
    case ORC_VARIANT_1: BC1_AVX2_LOOP(1) break;
    case ORC_VARIANT_2: BC1_AVX2_LOOP(2) break;
    default: BC1_AVX2_LOOP(3) break;
    }
    if (i < b1) bc1_range_fast(inverse, in, out, n, i, b1, variant, split);   /* < 8 blocks left */
}
static int have_avx2(void) { return __builtin_cpu_supports("avx2"); }
#else
static int have_avx2(void) { return 0; }
static void bc1_range_avx2(int inverse, const uint8_t *in, uint8_t *out, size_t n, size_t b0, size_t b1, int variant, int split) {
    bc1_range_fast(inverse, in, out, n, b0, b1, variant, split);
}
#endif
int orc_cpu_baseline_uses_avx2(void) { return have_avx2(); }

/* ------------------------------------------------------------------------------------------
 * AVX-512 paths for the CPU BASELINE timing — what the reference itself runs on an AVX-512 host (its best tier):
 *   BC1  32 blocks per iteration: four 64-byte loads, vpermt2d (colours / indices), vpermt2w (c0 / c1), YCoCg-R on 32
 *        16-bit lanes — the shape of core/dxt-lossless-transform-bc1/src/transform/with_split_colour_and_recorr/
 *        transform/avx512bw.rs and standard/transform/avx512f.rs:40-106, restated;
 *   BC2  16 blocks per iteration: vpermt2q (alpha), vpermt2d (colour part) — bc2 standard/transform/avx512*.rs;
 *   BC3  16 blocks per iteration: one vpermt2b per eight blocks gathers the alpha endpoints and the 6-byte alpha
 *        indices — core/dxt-lossless-transform-bc3/src/transform/standard/transform/avx512vbmi.rs, restated; the five
 *        variants the reference only has in scalar form get the same vector loop here, so the CPU arm is
 *        never slower than the reference would be.
 * Selected at run time (avx512bw + avx512vbmi); tests/test_oracle.py checks every settings combination, both
 * directions, odd block counts and unaligned buffers against the scalar restatement above.
 * ---------------------------------------------------------------------------------------- */
#if defined(__x86_64__)
#define ORC_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,avx512vbmi")))

ORC_AVX512 static inline __m512i decorr32(__m512i v, int variant) {
    const __m512i m5 = _mm512_set1_epi16(31), m1 = _mm512_set1_epi16(1);
    __m512i r = _mm512_srli_epi16(v, 11), g = _mm512_and_si512(_mm512_srli_epi16(v, 6), m5);
    __m512i gl = _mm512_and_si512(_mm512_srli_epi16(v, 5), m1), b = _mm512_and_si512(v, m5);
    __m512i co = _mm512_and_si512(_mm512_sub_epi16(r, b), m5);
    __m512i t = _mm512_and_si512(_mm512_add_epi16(b, _mm512_srli_epi16(co, 1)), m5);
    __m512i cg = _mm512_and_si512(_mm512_sub_epi16(g, t), m5);
    __m512i y = _mm512_and_si512(_mm512_add_epi16(t, _mm512_srli_epi16(cg, 1)), m5);
    if (variant == ORC_VARIANT_1)
        return _mm512_or_si512(_mm512_or_si512(_mm512_slli_epi16(y, 11), _mm512_slli_epi16(co, 6)), _mm512_or_si512(_mm512_slli_epi16(gl, 5), cg));
    if (variant == ORC_VARIANT_2)
        return _mm512_or_si512(_mm512_or_si512(_mm512_slli_epi16(gl, 15), _mm512_slli_epi16(y, 10)), _mm512_or_si512(_mm512_slli_epi16(co, 5), cg));
    return _mm512_or_si512(_mm512_or_si512(_mm512_slli_epi16(y, 11), _mm512_slli_epi16(co, 6)), _mm512_or_si512(_mm512_slli_epi16(cg, 1), gl));
}
ORC_AVX512 static inline __m512i recorr32(__m512i v, int variant) {
    const __m512i m5 = _mm512_set1_epi16(31), m1 = _mm512_set1_epi16(1);
    __m512i y, co, cg, gl;
    if (variant == ORC_VARIANT_1) {
        y = _mm512_srli_epi16(v, 11), co = _mm512_and_si512(_mm512_srli_epi16(v, 6), m5);
        gl = _mm512_and_si512(_mm512_srli_epi16(v, 5), m1), cg = _mm512_and_si512(v, m5);
    } else if (variant == ORC_VARIANT_2) {
        gl = _mm512_srli_epi16(v, 15), y = _mm512_and_si512(_mm512_srli_epi16(v, 10), m5);
        co = _mm512_and_si512(_mm512_srli_epi16(v, 5), m5), cg = _mm512_and_si512(v, m5);
    } else {
        y = _mm512_srli_epi16(v, 11), co = _mm512_and_si512(_mm512_srli_epi16(v, 6), m5);
        cg = _mm512_and_si512(_mm512_srli_epi16(v, 1), m5), gl = _mm512_and_si512(v, m1);
    }
    __m512i t = _mm512_and_si512(_mm512_sub_epi16(y, _mm512_srli_epi16(cg, 1)), m5);
    __m512i g = _mm512_and_si512(_mm512_add_epi16(cg, t), m5);
    __m512i b = _mm512_and_si512(_mm512_sub_epi16(t, _mm512_srli_epi16(co, 1)), m5);
    __m512i r = _mm512_and_si512(_mm512_add_epi16(b, co), m5);
    return _mm512_or_si512(_mm512_or_si512(_mm512_slli_epi16(r, 11), _mm512_slli_epi16(g, 6)), _mm512_or_si512(_mm512_slli_epi16(gl, 5), b));
}

/* index vectors (built once per call; a few dozen scalar stores against megabytes of payload) */
typedef struct {
    __m512i even_d, odd_d;        /* dwords 0,2,..,30 / 1,3,..,31 of a register pair */
    __m512i even_w, odd_w;        /* words 0,2,..,62 / 1,3,..,63 of a pair */
    __m512i zip_w_lo, zip_w_hi;   /* words (0,32,1,33,..) / (16,48,17,49,..): c0|c1 -> pairs */
    __m512i zip_d_lo, zip_d_hi;   /* dwords (0,16,1,17,..) / (8,24,..): colours|indices -> blocks */
    __m512i split_w, unsplit_w;   /* one register: words (0,2,..,30,1,3,..,31) and its inverse */
    __m512i even_q;               /* qwords 0,2,..,14 of a pair */
    __m512i colidx_d;             /* BC2/BC3: dwords 2,6,..,30 then 3,7,..,31 of a pair */
    __m512i lo_q, hi_q;           /* qwords (0,1,2,3,8,9,10,11) / (4,..,7,12,..,15) of a pair */
    __m512i bc2_z_lo, bc2_z_hi;   /* (alpha qwords, colour|index) -> blocks 0..3 / 4..7 */
} orc_idx512;

ORC_AVX512 static void orc_idx512_init(orc_idx512 *t) {
    uint32_t d[16];
    uint16_t w[32];
    uint64_t q[8];
    for (int i = 0; i < 16; i++) d[i] = 2 * i;
    t->even_d = _mm512_loadu_si512(d);
    for (int i = 0; i < 16; i++) d[i] = 2 * i + 1;
    t->odd_d = _mm512_loadu_si512(d);
    for (int i = 0; i < 32; i++) w[i] = 2 * i;
    t->even_w = _mm512_loadu_si512(w);
    for (int i = 0; i < 32; i++) w[i] = 2 * i + 1;
    t->odd_w = _mm512_loadu_si512(w);
    for (int i = 0; i < 16; i++) w[2 * i] = i, w[2 * i + 1] = 32 + i;
    t->zip_w_lo = _mm512_loadu_si512(w);
    for (int i = 0; i < 16; i++) w[2 * i] = 16 + i, w[2 * i + 1] = 48 + i;
    t->zip_w_hi = _mm512_loadu_si512(w);
    for (int i = 0; i < 8; i++) d[2 * i] = i, d[2 * i + 1] = 16 + i;
    t->zip_d_lo = _mm512_loadu_si512(d);
    for (int i = 0; i < 8; i++) d[2 * i] = 8 + i, d[2 * i + 1] = 24 + i;
    t->zip_d_hi = _mm512_loadu_si512(d);
    for (int i = 0; i < 16; i++) w[i] = 2 * i, w[16 + i] = 2 * i + 1;
    t->split_w = _mm512_loadu_si512(w);
    for (int i = 0; i < 16; i++) w[2 * i] = i, w[2 * i + 1] = 16 + i;
    t->unsplit_w = _mm512_loadu_si512(w);
    for (int i = 0; i < 8; i++) q[i] = 2 * i;
    t->even_q = _mm512_loadu_si512(q);
    for (int i = 0; i < 8; i++) d[i] = 4 * i + 2, d[8 + i] = 4 * i + 3;
    t->colidx_d = _mm512_loadu_si512(d);
    for (int i = 0; i < 4; i++) q[i] = i, q[4 + i] = 8 + i;
    t->lo_q = _mm512_loadu_si512(q);
    for (int i = 0; i < 4; i++) q[i] = 4 + i, q[4 + i] = 12 + i;
    t->hi_q = _mm512_loadu_si512(q);
    for (int k = 0; k < 4; k++) d[4 * k] = 2 * k, d[4 * k + 1] = 2 * k + 1, d[4 * k + 2] = 16 + k, d[4 * k + 3] = 24 + k;
    t->bc2_z_lo = _mm512_loadu_si512(d);
    for (int k = 0; k < 4; k++) d[4 * k] = 8 + 2 * k, d[4 * k + 1] = 9 + 2 * k, d[4 * k + 2] = 20 + k, d[4 * k + 3] = 28 + k;
    t->bc2_z_hi = _mm512_loadu_si512(d);
}

#define BC1_AVX512_LOOP(VARIANT)                                                                                     \
    for (; i + 32 <= b1; i += 32) {                                                                                  \
        if (!inverse) {                                                                                              \
            const __m512i z0 = _mm512_loadu_si512(in + 8 * i), z1 = _mm512_loadu_si512(in + 8 * i + 64);             \
            const __m512i z2 = _mm512_loadu_si512(in + 8 * i + 128), z3 = _mm512_loadu_si512(in + 8 * i + 192);      \
            __m512i ca = _mm512_permutex2var_epi32(z0, T.even_d, z1), cb = _mm512_permutex2var_epi32(z2, T.even_d, z3); \
            const __m512i xa = _mm512_permutex2var_epi32(z0, T.odd_d, z1), xb = _mm512_permutex2var_epi32(z2, T.odd_d, z3); \
            if (VARIANT) ca = decorr32(ca, VARIANT), cb = decorr32(cb, VARIANT);                                     \
            if (split) {                                                                                             \
                _mm512_storeu_si512(out + 2 * i, _mm512_permutex2var_epi16(ca, T.even_w, cb));                       \
                _mm512_storeu_si512(out + len / 4 + 2 * i, _mm512_permutex2var_epi16(ca, T.odd_w, cb));              \
            } else {                                                                                                 \
                _mm512_storeu_si512(out + 4 * i, ca);                                                                \
                _mm512_storeu_si512(out + 4 * i + 64, cb);                                                           \
            }                                                                                                        \
            _mm512_storeu_si512(out + len / 2 + 4 * i, xa);                                                          \
            _mm512_storeu_si512(out + len / 2 + 4 * i + 64, xb);                                                     \
        } else {                                                                                                     \
            __m512i ca, cb;                                                                                          \
            if (split) {                                                                                             \
                const __m512i c0 = _mm512_loadu_si512(in + 2 * i), c1 = _mm512_loadu_si512(in + len / 4 + 2 * i);    \
                ca = _mm512_permutex2var_epi16(c0, T.zip_w_lo, c1), cb = _mm512_permutex2var_epi16(c0, T.zip_w_hi, c1); \
            } else ca = _mm512_loadu_si512(in + 4 * i), cb = _mm512_loadu_si512(in + 4 * i + 64);                    \
            const __m512i xa = _mm512_loadu_si512(in + len / 2 + 4 * i), xb = _mm512_loadu_si512(in + len / 2 + 4 * i + 64); \
            if (VARIANT) ca = recorr32(ca, VARIANT), cb = recorr32(cb, VARIANT);                                     \
            _mm512_storeu_si512(out + 8 * i, _mm512_permutex2var_epi32(ca, T.zip_d_lo, xa));                         \
            _mm512_storeu_si512(out + 8 * i + 64, _mm512_permutex2var_epi32(ca, T.zip_d_hi, xa));                    \
            _mm512_storeu_si512(out + 8 * i + 128, _mm512_permutex2var_epi32(cb, T.zip_d_lo, xb));                   \
            _mm512_storeu_si512(out + 8 * i + 192, _mm512_permutex2var_epi32(cb, T.zip_d_hi, xb));                   \
        }                                                                                                            \
    }

ORC_AVX512 static void bc1_range_avx512(int inverse, const uint8_t *restrict in, uint8_t *restrict out, size_t n, size_t b0,
                                        size_t b1, int variant, int split) {
    const size_t len = n * 8;
    orc_idx512 T;
    orc_idx512_init(&T);
    size_t i = b0;
    switch (variant) {
    case ORC_VARIANT_NONE: BC1_AVX512_LOOP(0) break;
    case ORC_VARIANT_1: BC1_AVX512_LOOP(1) break;
    case ORC_VARIANT_2: BC1_AVX512_LOOP(2) break;
    default: BC1_AVX512_LOOP(3) break;
    }
    if (i < b1) bc1_range_fast(inverse, in, out, n, i, b1, variant, split);   /* < 32 blocks left */
}

/* colour part shared by BC2 and BC3: 16 blocks, colours at byte 8 and indices at byte 12 of every 16-byte block;
 * cso / c1o / ixo are the byte offsets of the colour (or c0), c1 and index sections */
#define BCX_COLOURS_FWD(VARIANT)                                                                                     \
    {                                                                                                                \
        const __m512i t01 = _mm512_permutex2var_epi32(z0, T.colidx_d, z1), t23 = _mm512_permutex2var_epi32(z2, T.colidx_d, z3); \
        __m512i col = _mm512_permutex2var_epi64(t01, T.lo_q, t23);                                                   \
        const __m512i idx = _mm512_permutex2var_epi64(t01, T.hi_q, t23);                                             \
        if (VARIANT) col = decorr32(col, VARIANT);                                                                   \
        if (split_c) {                                                                                               \
            const __m512i s = _mm512_permutexvar_epi16(T.split_w, col);                                              \
            _mm256_storeu_si256((__m256i *)(out + cso + 2 * i), _mm512_castsi512_si256(s));                          \
            _mm256_storeu_si256((__m256i *)(out + c1o + 2 * i), _mm512_extracti64x4_epi64(s, 1));                    \
        } else _mm512_storeu_si512(out + cso + 4 * i, col);                                                          \
        _mm512_storeu_si512(out + ixo + 4 * i, idx);                                                                 \
    }
#define BCX_COLOURS_INV(VARIANT)                                                                                     \
    __m512i col;                                                                                                     \
    if (split_c) {                                                                                                   \
        const __m512i s = _mm512_inserti64x4(_mm512_castsi256_si512(_mm256_loadu_si256((const __m256i *)(in + cso + 2 * i))), \
                                             _mm256_loadu_si256((const __m256i *)(in + c1o + 2 * i)), 1);           \
        col = _mm512_permutexvar_epi16(T.unsplit_w, s);                                                              \
    } else col = _mm512_loadu_si512(in + cso + 4 * i);                                                               \
    const __m512i idx = _mm512_loadu_si512(in + ixo + 4 * i);                                                        \
    if (VARIANT) col = recorr32(col, VARIANT);                                                                       \
    const __m512i t01 = _mm512_permutex2var_epi64(col, T.lo_q, idx), t23 = _mm512_permutex2var_epi64(col, T.hi_q, idx);

#define BC2_AVX512_LOOP(VARIANT)                                                                                     \
    for (; i + 16 <= b1; i += 16) {                                                                                  \
        if (!inverse) {                                                                                              \
            const __m512i z0 = _mm512_loadu_si512(in + 16 * i), z1 = _mm512_loadu_si512(in + 16 * i + 64);           \
            const __m512i z2 = _mm512_loadu_si512(in + 16 * i + 128), z3 = _mm512_loadu_si512(in + 16 * i + 192);    \
            _mm512_storeu_si512(out + 8 * i, _mm512_permutex2var_epi64(z0, T.even_q, z1));                           \
            _mm512_storeu_si512(out + 8 * i + 64, _mm512_permutex2var_epi64(z2, T.even_q, z3));                      \
            BCX_COLOURS_FWD(VARIANT)                                                                                 \
        } else {                                                                                                     \
            const __m512i a01 = _mm512_loadu_si512(in + 8 * i), a23 = _mm512_loadu_si512(in + 8 * i + 64);           \
            BCX_COLOURS_INV(VARIANT)                                                                                 \
            _mm512_storeu_si512(out + 16 * i, _mm512_permutex2var_epi32(a01, T.bc2_z_lo, t01));                      \
            _mm512_storeu_si512(out + 16 * i + 64, _mm512_permutex2var_epi32(a01, T.bc2_z_hi, t01));                 \
            _mm512_storeu_si512(out + 16 * i + 128, _mm512_permutex2var_epi32(a23, T.bc2_z_lo, t23));                \
            _mm512_storeu_si512(out + 16 * i + 192, _mm512_permutex2var_epi32(a23, T.bc2_z_hi, t23));                \
        }                                                                                                            \
    }

ORC_AVX512 static void bc2_range_avx512(int inverse, const uint8_t *restrict in, uint8_t *restrict out, size_t n, size_t b0,
                                        size_t b1, int variant, int split_c) {
    const size_t len = n * 16, cso = len / 2, c1o = len / 2 + len / 8, ixo = len / 2 + len / 4;
    orc_idx512 T;
    orc_idx512_init(&T);
    size_t i = b0;
    switch (variant) {
    case ORC_VARIANT_NONE: BC2_AVX512_LOOP(0) break;
    case ORC_VARIANT_1: BC2_AVX512_LOOP(1) break;
    case ORC_VARIANT_2: BC2_AVX512_LOOP(2) break;
    default: BC2_AVX512_LOOP(3) break;
    }
    if (i < b1) bc2_range(inverse, in, out, n, i, b1, variant, split_c);
}

/* BC3 alpha part, eight blocks (two registers) at a time with vpermt2b.
 *   forward : P = [endpoint bytes: 16 | alpha index bytes: 48]   (split_a: a0 x 8 | a1 x 8 | indices)
 *   inverse : block register = vpermt2b(P, sel, T) with T = [colours of 8 blocks : 32 | indices : 32]            */
typedef struct {
    __m512i fwd;          /* (z_lo, z_hi) -> P */
    __m512i inv_lo, inv_hi; /* (P, T) -> blocks 0..3 / 4..7 */
} orc_bc3_sel;
ORC_AVX512 static void orc_bc3_sel_init(orc_bc3_sel *s, int split_a) {
    uint8_t f[64], lo[64], hi[64];
    for (int b = 0; b < 8; b++) {
        const int src = 16 * b;   /* byte index in the concatenation (z_lo, z_hi): vpermt2b takes 7 index bits */
        const int e0 = split_a ? b : 2 * b, e1 = split_a ? 8 + b : 2 * b + 1;
        f[e0] = (uint8_t)src, f[e1] = (uint8_t)(src + 1);
        for (int k = 0; k < 6; k++) f[16 + 6 * b + k] = (uint8_t)(src + 2 + k);
        uint8_t *z = b < 4 ? lo + 16 * b : hi + 16 * (b - 4);
        z[0] = (uint8_t)e0, z[1] = (uint8_t)e1;
        for (int k = 0; k < 6; k++) z[2 + k] = (uint8_t)(16 + 6 * b + k);
        for (int k = 0; k < 4; k++) z[8 + k] = (uint8_t)(64 + 4 * b + k), z[12 + k] = (uint8_t)(64 + 32 + 4 * b + k);
    }
    s->fwd = _mm512_loadu_si512(f), s->inv_lo = _mm512_loadu_si512(lo), s->inv_hi = _mm512_loadu_si512(hi);
}

#define BC3_AVX512_LOOP(VARIANT)                                                                                     \
    for (; i + 16 <= b1; i += 16) {                                                                                  \
        if (!inverse) {                                                                                              \
            const __m512i z0 = _mm512_loadu_si512(in + 16 * i), z1 = _mm512_loadu_si512(in + 16 * i + 64);           \
            const __m512i z2 = _mm512_loadu_si512(in + 16 * i + 128), z3 = _mm512_loadu_si512(in + 16 * i + 192);    \
            const __m512i p01 = _mm512_permutex2var_epi8(z0, S.fwd, z1), p23 = _mm512_permutex2var_epi8(z2, S.fwd, z3); \
            const __m128i e01 = _mm512_castsi512_si128(p01), e23 = _mm512_castsi512_si128(p23);                      \
            if (split_a) {                                                                                           \
                _mm_storeu_si128((__m128i *)(out + i), _mm_unpacklo_epi64(e01, e23));                                \
                _mm_storeu_si128((__m128i *)(out + n + i), _mm_unpackhi_epi64(e01, e23));                            \
            } else {                                                                                                 \
                _mm_storeu_si128((__m128i *)(out + 2 * i), e01);                                                     \
                _mm_storeu_si128((__m128i *)(out + 2 * i + 16), e23);                                                \
            }                                                                                                        \
            /* 48 index bytes each: bytes 16..63 of the register land at aio + 6 i (the masked-off bytes are never touched) */ \
            _mm512_mask_storeu_epi8(out + aio + 6 * i - 16, 0xFFFFFFFFFFFF0000ull, p01);                             \
            _mm512_mask_storeu_epi8(out + aio + 6 * i + 48 - 16, 0xFFFFFFFFFFFF0000ull, p23);                        \
            BCX_COLOURS_FWD(VARIANT)                                                                                 \
        } else {                                                                                                     \
            __m128i e01, e23;                                                                                        \
            if (split_a) {                                                                                           \
                const __m128i a0 = _mm_loadu_si128((const __m128i *)(in + i)), a1 = _mm_loadu_si128((const __m128i *)(in + n + i)); \
                e01 = _mm_unpacklo_epi64(a0, a1), e23 = _mm_unpackhi_epi64(a0, a1);                                  \
            } else e01 = _mm_loadu_si128((const __m128i *)(in + 2 * i)), e23 = _mm_loadu_si128((const __m128i *)(in + 2 * i + 16)); \
            const __m512i p01 = _mm512_mask_loadu_epi8(_mm512_castsi128_si512(e01), 0xFFFFFFFFFFFF0000ull, in + aio + 6 * i - 16); \
            const __m512i p23 = _mm512_mask_loadu_epi8(_mm512_castsi128_si512(e23), 0xFFFFFFFFFFFF0000ull, in + aio + 6 * i + 48 - 16); \
            BCX_COLOURS_INV(VARIANT)                                                                                 \
            _mm512_storeu_si512(out + 16 * i, _mm512_permutex2var_epi8(p01, S.inv_lo, t01));                         \
            _mm512_storeu_si512(out + 16 * i + 64, _mm512_permutex2var_epi8(p01, S.inv_hi, t01));                    \
            _mm512_storeu_si512(out + 16 * i + 128, _mm512_permutex2var_epi8(p23, S.inv_lo, t23));                   \
            _mm512_storeu_si512(out + 16 * i + 192, _mm512_permutex2var_epi8(p23, S.inv_hi, t23));                   \
        }                                                                                                            \
    }

ORC_AVX512 static void bc3_range_avx512(int inverse, const uint8_t *restrict in, uint8_t *restrict out, size_t n, size_t b0,
                                        size_t b1, int variant, int split_a, int split_c) {
    const size_t aio = 2 * n, cso = 8 * n, c1o = 10 * n, ixo = 12 * n;
    orc_idx512 T;
    orc_bc3_sel S;
    orc_idx512_init(&T);
    orc_bc3_sel_init(&S, split_a);
    size_t i = b0;
    switch (variant) {
    case ORC_VARIANT_NONE: BC3_AVX512_LOOP(0) break;
    case ORC_VARIANT_1: BC3_AVX512_LOOP(1) break;
    case ORC_VARIANT_2: BC3_AVX512_LOOP(2) break;
    default: BC3_AVX512_LOOP(3) break;
    }
    if (i < b1) bc3_range(inverse, in, out, n, i, b1, variant, split_a, split_c);
}
static int have_avx512(void) {
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
           __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vbmi");
}
#else
static int have_avx512(void) { return 0; }
static void bc1_range_avx512(int inverse, const uint8_t *in, uint8_t *out, size_t n, size_t b0, size_t b1, int variant, int split) {
    bc1_range_fast(inverse, in, out, n, b0, b1, variant, split);
}
static void bc2_range_avx512(int inverse, const uint8_t *in, uint8_t *out, size_t n, size_t b0, size_t b1, int variant, int split) {
    bc2_range(inverse, in, out, n, b0, b1, variant, split);
}
static void bc3_range_avx512(int inverse, const uint8_t *in, uint8_t *out, size_t n, size_t b0, size_t b1, int variant, int sa, int sc) {
    bc3_range(inverse, in, out, n, b0, b1, variant, sa, sc);
}
#endif
/* 0: scalar / word-wise, 2: AVX2 (BC1 only), 5: AVX-512 (BC1, BC2, BC3).  ORC_ISA=scalar|avx2 caps it (tests). */
static int baseline_isa(void) {
    static int isa = -1;
    if (isa < 0) {
        int best = have_avx512() ? 5 : have_avx2() ? 2 : 0;
        const char *cap = getenv("ORC_ISA");
        if (cap && !strcmp(cap, "scalar")) best = 0;
        if (cap && !strcmp(cap, "avx2") && best > 2) best = 2;
        isa = best;
    }
    return isa;
}
int orc_cpu_baseline_isa(void) { return baseline_isa(); }

void orc_bc1_transform(const uint8_t *in, uint8_t *out, size_t len, int v, int s) { bc1_range(0, in, out, len / 8, 0, len / 8, v, s); }
void orc_bc1_untransform(const uint8_t *in, uint8_t *out, size_t len, int v, int s) { bc1_range(1, in, out, len / 8, 0, len / 8, v, s); }
void orc_bc2_transform(const uint8_t *in, uint8_t *out, size_t len, int v, int s) { bc2_range(0, in, out, len / 16, 0, len / 16, v, s); }
void orc_bc2_untransform(const uint8_t *in, uint8_t *out, size_t len, int v, int s) { bc2_range(1, in, out, len / 16, 0, len / 16, v, s); }
void orc_bc3_transform(const uint8_t *in, uint8_t *out, size_t len, int v, int sa, int sc) { bc3_range(0, in, out, len / 16, 0, len / 16, v, sa, sc); }
void orc_bc3_untransform(const uint8_t *in, uint8_t *out, size_t len, int v, int sa, int sc) { bc3_range(1, in, out, len / 16, 0, len / 16, v, sa, sc); }

/* split_color_endpoints: [c0 c1]×k → c0×k | c1×k
 * (common/src/transforms/split_565_color_endpoints/portable32.rs:18-64). */
void orc_split_color_endpoints(const uint8_t *in, uint8_t *out, size_t len_bytes) {
    size_t pairs = len_bytes / 4;
    for (size_t i = 0; i < pairs; i++) {
        memcpy(out + 2 * i, in + 4 * i, 2);
        memcpy(out + len_bytes / 2 + 2 * i, in + 4 * i + 2, 2);
    }
}

/* ------------------------------------------------------------------------------------------
 * LTU estimator (restated; PARITY UNPINNED — see ltu_params.h)
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t rd32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* The parts of the restatement that are unverified against the crate are run-time parameters (defaults: ltu_params.h),
 * so that every variant the product's estimator supports has a CPU checker:
 *   hash_bits 12..17; index_top != 0: index = product >> (32 - hash_bits), else index = product & mask;
 *   group = positions per loop iteration (all compares, then all updates): 4 or 1. */
static int g_ltu_hash_bits = LTU_HASH_BITS, g_ltu_index_top = 1, g_ltu_group = LTU_GROUP;

int orc_ltu_set_params(int hash_bits, int index_top, int group) {
    if (hash_bits < 8 || hash_bits > 24 || group < 1 || group > 8) return 1;
    g_ltu_hash_bits = hash_bits, g_ltu_index_top = index_top != 0, g_ltu_group = group;
    return 0;
}

size_t orc_ltu_num_lz_matches_params(const uint8_t *data, size_t len, int hash_bits, int index_top, int group) {
    uint32_t *table = (uint32_t *)calloc((size_t)1 << hash_bits, sizeof(uint32_t));
    if (!table) abort();
    size_t end = len > LTU_TAIL_GUARD ? len - LTU_TAIL_GUARD : 0;
    size_t matches = 0;
    for (size_t i = 0; i < end; i += (size_t)group) {
        uint32_t d[8], idx[8];
        for (int k = 0; k < group; k++) {
            d[k] = rd32(data + i + k) & LTU_KEY_MASK;
            const uint32_t product = (uint32_t)(d[k] * LTU_GOLDEN_RATIO);
            idx[k] = index_top ? product >> (32 - hash_bits) : product & (((uint32_t)1 << hash_bits) - 1u);
        }
        for (int k = 0; k < group; k++) matches += table[idx[k]] == d[k];
        for (int k = 0; k < group; k++) table[idx[k]] = d[k];
    }
    free(table);
    return matches;
}

size_t orc_ltu_num_lz_matches(const uint8_t *data, size_t len) {
    return orc_ltu_num_lz_matches_params(data, len, g_ltu_hash_bits, g_ltu_index_top, g_ltu_group);
}

/* estimate_size / estimate_compressed_size: extensions/estimators/dxt-lossless-transform-ltu/src/lib.rs:67-119
 * (null or empty → 0; otherwise len.saturating_sub(matches)). */
size_t orc_ltu_estimate(const uint8_t *data, size_t len) {
    if (!data || len == 0) return 0;
    size_t m = orc_ltu_num_lz_matches(data, len);
    return len > m ? len - m : 0;
}

static int ltu_cb(void *ctx, const uint8_t *data, size_t len, size_t *out) {
    (void)ctx;
    *out = orc_ltu_estimate(data, len);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * transform_bcN_auto  (brute-force search, strict '<', first in order wins ties, final
 * re-transform if the winner was not the last candidate tested)
 * ---------------------------------------------------------------------------------------- */

/* bc1/src/transform/settings.rs:81-98 (BC2 settings.rs:81-98 is identical). */
static const int BC12_FAST[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}};
static const int BC12_ALL[8][2] = {{2, 0}, {0, 0}, {0, 1}, {3, 0}, {3, 1}, {2, 1}, {1, 0}, {1, 1}};
/* bc3/src/transform/settings.rs:91-121: (variant, split_alpha, split_colour). */
static const int BC3_FAST[8][3] = {{1, 1, 0}, {1, 1, 1}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {0, 0, 0}, {1, 0, 0}};
static const int BC3_ALL[16][3] = {{2, 1, 0}, {2, 1, 1}, {3, 1, 1}, {3, 1, 0}, {1, 1, 0}, {3, 0, 1}, {1, 1, 1}, {2, 0, 1},
                                   {2, 0, 0}, {3, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {0, 0, 0}, {1, 0, 0}};

/* bc1/src/transform/transform_auto.rs:200-270 (estimate(out, len/2));
 * bc2/src/transform/transform_auto.rs:196-272 (estimate(out+len/2, len/4)). */
static int bc12_auto(int fmt, const uint8_t *in, uint8_t *out, size_t len, int use_all, orc_estimate_fn est,
                     void *ctx, int *ov, int *os, size_t *sizes) {
    const int(*order)[2] = use_all ? BC12_ALL : BC12_FAST;
    int k = use_all ? 8 : 4;
    if (!est) est = ltu_cb;
    /* Bc1TransformSettings::default() = (Variant1, split) — settings.rs:35-43 */
    int best_v = 1, best_s = 1, last_v = 1, last_s = 1;
    size_t best = SIZE_MAX;
    for (int c = 0; c < k; c++) {
        int v = order[c][0], s = order[c][1];
        size_t sz;
        int rc;
        if (fmt == 1) {
            orc_bc1_transform(in, out, len, v, s);
            rc = est(ctx, out, len / 2, &sz);
        } else {
            orc_bc2_transform(in, out, len, v, s);
            rc = est(ctx, out + len / 2, len / 4, &sz);
        }
        last_v = v;
        last_s = s;
        if (rc) return rc;
        if (sizes) sizes[c] = sz;
        if (sz < best) {
            best = sz;
            best_v = v;
            best_s = s;
        }
    }
    if (best_v != last_v || best_s != last_s) {
        if (fmt == 1) orc_bc1_transform(in, out, len, best_v, best_s);
        else orc_bc2_transform(in, out, len, best_v, best_s);
    }
    if (ov) *ov = best_v;
    if (os) *os = best_s;
    return 0;
}

int orc_bc1_transform_auto(const uint8_t *in, uint8_t *out, size_t len, int use_all, orc_estimate_fn est, void *ctx,
                           int *ov, int *os) {
    return bc12_auto(1, in, out, len, use_all, est, ctx, ov, os, NULL);
}
int orc_bc2_transform_auto(const uint8_t *in, uint8_t *out, size_t len, int use_all, orc_estimate_fn est, void *ctx,
                           int *ov, int *os) {
    return bc12_auto(2, in, out, len, use_all, est, ctx, ov, os, NULL);
}

/* bc3/src/transform/transform_auto.rs:196-293: estimate(out, 2n) + estimate(out + len/2, 4n). */
static int bc3_auto(const uint8_t *in, uint8_t *out, size_t len, int use_all, orc_estimate_fn est, void *ctx, int *ov,
                    int *osa, int *osc, size_t *sizes) {
    const int(*order)[3] = use_all ? BC3_ALL : BC3_FAST;
    int k = use_all ? 16 : 8;
    if (!est) est = ltu_cb;
    /* Bc3TransformSettings::default() = (Variant1, split_alpha, split_colour) — settings.rs:39-48 */
    int bv = 1, ba = 1, bc = 1, lv = 1, la = 1, lc = 1;
    size_t best = SIZE_MAX, n = len / 16;
    for (int c = 0; c < k; c++) {
        int v = order[c][0], sa = order[c][1], sc = order[c][2];
        orc_bc3_transform(in, out, len, v, sa, sc);
        lv = v;
        la = sa;
        lc = sc;
        size_t a_sz, c_sz;
        int rc = est(ctx, out, n * 2, &a_sz);
        if (rc) return rc;
        rc = est(ctx, out + len / 2, n * 4, &c_sz);
        if (rc) return rc;
        size_t total = a_sz + c_sz;
        if (sizes) sizes[c] = total;
        if (total < best) {
            best = total;
            bv = v;
            ba = sa;
            bc = sc;
        }
    }
    if (bv != lv || ba != la || bc != lc) orc_bc3_transform(in, out, len, bv, ba, bc);
    if (ov) *ov = bv;
    if (osa) *osa = ba;
    if (osc) *osc = bc;
    return 0;
}

int orc_bc3_transform_auto(const uint8_t *in, uint8_t *out, size_t len, int use_all, orc_estimate_fn est, void *ctx,
                           int *ov, int *osa, int *osc) {
    return bc3_auto(in, out, len, use_all, est, ctx, ov, osa, osc, NULL);
}

int orc_bc1_auto_estimates(const uint8_t *in, uint8_t *scratch, size_t len, int use_all, size_t *sizes) {
    bc12_auto(1, in, scratch, len, use_all, NULL, NULL, NULL, NULL, sizes);
    return use_all ? 8 : 4;
}
int orc_bc2_auto_estimates(const uint8_t *in, uint8_t *scratch, size_t len, int use_all, size_t *sizes) {
    bc12_auto(2, in, scratch, len, use_all, NULL, NULL, NULL, NULL, sizes);
    return use_all ? 8 : 4;
}
int orc_bc3_auto_estimates(const uint8_t *in, uint8_t *scratch, size_t len, int use_all, size_t *sizes) {
    bc3_auto(in, scratch, len, use_all, NULL, NULL, NULL, NULL, NULL, sizes);
    return use_all ? 16 : 8;
}

/* ------------------------------------------------------------------------------------------
 * Reference test-data generators
 * ---------------------------------------------------------------------------------------- */

/* core/dxt-lossless-transform-bc1/src/test_prelude.rs:81-105 */
void orc_generate_bc1_test_data(uint8_t *p, size_t num_blocks) {
    uint8_t color = 0, index = 128;
    for (size_t i = 0; i < num_blocks; i++, p += 8) {
        for (int k = 0; k < 4; k++) p[k] = (uint8_t)(color + k);
        color = (uint8_t)(color + 4);
        for (int k = 0; k < 4; k++) p[4 + k] = (uint8_t)(index + k);
        index = (uint8_t)(index + 4);
    }
}

/* core/dxt-lossless-transform-bc2/src/test_prelude.rs:151-190 */
void orc_generate_bc2_test_data(uint8_t *p, size_t num_blocks) {
    uint8_t alpha = 0, color = 0x80, index = 0xC0;
    for (size_t i = 0; i < num_blocks; i++, p += 16) {
        for (int k = 0; k < 8; k++) p[k] = (uint8_t)(alpha + k);
        alpha = (uint8_t)(alpha + 8);
        for (int k = 0; k < 4; k++) p[8 + k] = (uint8_t)(color + k);
        color = (uint8_t)(color + 4);
        for (int k = 0; k < 4; k++) p[12 + k] = (uint8_t)(index + k);
        index = (uint8_t)(index + 4);
    }
}

/* core/dxt-lossless-transform-bc3/src/test_prelude.rs:45-101 (note the per-field wrap rules). */
void orc_generate_bc3_test_data(uint8_t *p, size_t num_blocks) {
    uint8_t alpha = 0, aidx = 32, color = 128, index = 192;
    for (size_t i = 0; i < num_blocks; i++, p += 16) {
        p[0] = alpha;
        p[1] = (uint8_t)(alpha + 1);
        alpha = (uint8_t)(alpha + 2);
        if (alpha >= 32) alpha = (uint8_t)(alpha - 32);
        for (int k = 0; k < 6; k++) p[2 + k] = (uint8_t)(aidx + k);
        aidx = (uint8_t)(aidx + 6);
        if (aidx >= 128) aidx = (uint8_t)(aidx - 96);
        for (int k = 0; k < 4; k++) p[8 + k] = (uint8_t)(color + k);
        color = (uint8_t)(color + 4);
        if (color >= 192) color = (uint8_t)(color - 64);
        for (int k = 0; k < 4; k++) p[12 + k] = (uint8_t)(index + k);
        index = (uint8_t)(index + 4);
        if (index < 192) index = (uint8_t)(index - 64);
    }
}

/* ------------------------------------------------------------------------------------------
 * experimental::normalize_blocks (BC1) — core/dxt-lossless-transform-bc1/src/experimental/normalize_blocks/
 *   decode_bc1_block            core/dxt-lossless-transform-bc1/src/util/bc1_decode.rs:7-63
 *   Color565::{red,green,blue}  core/dxt-lossless-transform-common/src/color_565/mod.rs:154-191
 *   Color565::from_rgb          .../color_565/mod.rs:108-115        (to_565_lossy)
 *   normalize_blocks            normalize.rs:38-101, normalize_blocks_impl :104-176
 *   write_normalized_solid_color_block   normalize.rs:199-247
 *   normalize_split_blocks_in_place      normalize.rs:286-386
 *   normalize_blocks_all_modes           normalize.rs:417-484
 *   transform_bc1_with_normalize_blocks  transform.rs:65-176
 *   transform_bc1_auto_with_normalization transform.rs:222-340, test_normalize_variant_with_normalization :344-408
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint8_t r, g, b, a; } rgba8;
static inline int rgba_eq(rgba8 x, rgba8 y) { return x.r == y.r && x.g == y.g && x.b == y.b && x.a == y.a; }
static inline uint8_t red8(uint16_t v) { unsigned r = (v & 0xF800u) >> 11; return (uint8_t)((r << 3) | (r >> 2)); }
static inline uint8_t green8(uint16_t v) { unsigned g = (v & 0x07E0u) >> 5; return (uint8_t)((g << 2) | (g >> 4)); }
static inline uint8_t blue8(uint16_t v) { unsigned b = v & 0x001Fu; return (uint8_t)((b << 3) | (b >> 2)); }

static void decode_bc1_block(const uint8_t *src, rgba8 px[16]) {
    const uint16_t c0 = rd16(src), c1 = rd16(src + 2);
    const uint32_t idx = (uint32_t)src[4] | ((uint32_t)src[5] << 8) | ((uint32_t)src[6] << 16) | ((uint32_t)src[7] << 24);
    const unsigned r0 = red8(c0), g0 = green8(c0), b0 = blue8(c0), r1 = red8(c1), g1 = green8(c1), b1 = blue8(c1);
    rgba8 dict[4];
    dict[0] = (rgba8){(uint8_t)r0, (uint8_t)g0, (uint8_t)b0, 255};
    dict[1] = (rgba8){(uint8_t)r1, (uint8_t)g1, (uint8_t)b1, 255};
    if (c0 > c1) { /* four-colour block */
        dict[2] = (rgba8){(uint8_t)((2 * r0 + r1) / 3), (uint8_t)((2 * g0 + g1) / 3), (uint8_t)((2 * b0 + b1) / 3), 255};
        dict[3] = (rgba8){(uint8_t)((r0 + 2 * r1) / 3), (uint8_t)((g0 + 2 * g1) / 3), (uint8_t)((b0 + 2 * b1) / 3), 255};
    } else { /* three colours + transparent black */
        dict[2] = (rgba8){(uint8_t)((r0 + r1) / 2), (uint8_t)((g0 + g1) / 2), (uint8_t)((b0 + b1) / 2), 255};
        dict[3] = (rgba8){0, 0, 0, 0};
    }
    for (int i = 0; i < 16; i++) px[i] = dict[(idx >> (2 * i)) & 3];
}

/* BlockCase of normalize_blocks_impl: 0 = CannotNormalize, 1 = Transparent, 2 = SolidColorRoundtrippable (+ colour) */
static int classify_block(const uint8_t *src, uint16_t *color565) {
    rgba8 px[16];
    decode_bc1_block(src, px);
    for (int i = 1; i < 16; i++)
        if (!rgba_eq(px[i], px[0])) return 0;
    if (px[0].a == 0) return 1;
    const uint16_t c = (uint16_t)(((px[0].r & 0xF8u) << 8) | ((px[0].g & 0xFCu) << 3) | (px[0].b >> 3));
    *color565 = c;
    const rgba8 back = {red8(c), green8(c), blue8(c), 255};
    return rgba_eq(back, px[0]) ? 2 : 0;
}

static void write_normalized_solid(uint8_t *dst, const uint8_t *src, uint16_t c, int mode) {
    dst[0] = (uint8_t)c, dst[1] = (uint8_t)(c >> 8);
    if (mode == ORC_NORM_NONE) {
        memcpy(dst, src, 8);
    } else if (mode == ORC_NORM_COLOR0_ONLY) {
        memset(dst + 2, 0, 6);
    } else {
        dst[2] = (uint8_t)c, dst[3] = (uint8_t)(c >> 8);
        memset(dst + 4, 0, 4);
    }
}

void orc_bc1_normalize_blocks(const uint8_t *in, uint8_t *out, size_t len, int mode) {
    if (mode == ORC_NORM_NONE) {
        if (in != out) memcpy(out, in, len);
        return;
    }
    for (size_t i = 0; i + 8 <= len; i += 8) {
        uint16_t c = 0;
        uint8_t blk[8];
        memcpy(blk, in + i, 8); /* in-place use is allowed */
        const int kind = classify_block(blk, &c);
        if (kind == 1) memset(out + i, 0xFF, 8);
        else if (kind == 2) write_normalized_solid(out + i, blk, c, mode);
        else memcpy(out + i, blk, 8);
    }
}

int orc_bc1_normalize_blocks_all_modes(const uint8_t *in, uint8_t *out_none, uint8_t *out_color0, uint8_t *out_replicate,
                                       size_t len) {
    uint8_t *outs[3] = {out_none, out_color0, out_replicate};
    int any = 0;
    for (size_t i = 0; i + 8 <= len; i += 8) {
        uint16_t c = 0;
        const int kind = classify_block(in + i, &c);
        if (kind) any = 1;
        for (int m = 0; m < 3; m++) {
            if (kind == 1) memset(outs[m] + i, 0xFF, 8);
            else if (kind == 2) write_normalized_solid(outs[m] + i, in + i, c, m);
            else memcpy(outs[m] + i, in + i, 8);
        }
    }
    return any;
}

void orc_bc1_normalize_split_blocks_in_place(uint8_t *colors, uint8_t *indices, size_t num_blocks, int mode) {
    if (mode == ORC_NORM_NONE) return;
    for (size_t b = 0; b < num_blocks; b++) {
        uint8_t blk[8];
        memcpy(blk, colors + 4 * b, 4);
        memcpy(blk + 4, indices + 4 * b, 4);
        uint16_t c = 0;
        const int kind = classify_block(blk, &c);
        if (kind == 1) {
            memset(colors + 4 * b, 0xFF, 4);
            memset(indices + 4 * b, 0xFF, 4);
        } else if (kind == 2) {
            colors[4 * b] = (uint8_t)c, colors[4 * b + 1] = (uint8_t)(c >> 8);
            if (mode == ORC_NORM_COLOR0_ONLY) colors[4 * b + 2] = colors[4 * b + 3] = 0;
            else colors[4 * b + 2] = (uint8_t)c, colors[4 * b + 3] = (uint8_t)(c >> 8);
            memset(indices + 4 * b, 0, 4);
        }
    }
}

static void decorrelate_in_place(uint8_t *colours, size_t count, int variant) {
    if (variant == ORC_VARIANT_NONE) return;
    for (size_t i = 0; i < count; i++) wr16(colours + 2 * i, orc_decorrelate(rd16(colours + 2 * i), variant));
}

/* The four arms of transform_bc1_with_normalize_blocks, step by step as the reference runs them. */
void orc_bc1_transform_with_normalize_blocks(const uint8_t *in, uint8_t *out, size_t len, int norm, int variant, int split) {
    const size_t n = len / 8;
    uint8_t *work = (uint8_t *)malloc(len / 2 + 8);
    if (split) {
        for (size_t b = 0; b < n; b++) { /* colours to the work area, indices to their final place */
            memcpy(work + 4 * b, in + 8 * b, 4);
            memcpy(out + len / 2 + 4 * b, in + 8 * b + 4, 4);
        }
        if (norm != ORC_NORM_NONE) orc_bc1_normalize_split_blocks_in_place(work, out + len / 2, n, norm);
        orc_split_color_endpoints(work, out, len / 2);
    } else {
        for (size_t b = 0; b < n; b++) {
            memcpy(out + 4 * b, in + 8 * b, 4);
            memcpy(out + len / 2 + 4 * b, in + 8 * b + 4, 4);
        }
        if (norm != ORC_NORM_NONE) orc_bc1_normalize_split_blocks_in_place(out, out + len / 2, n, norm);
    }
    decorrelate_in_place(out, len / 4, variant);
    free(work);
}

/* transform_bc1_auto_with_normalization.  est == NULL: the LTU restatement.  Returns 0, or 1 when max_compressed_size
 * style failures would abort (not modelled: the callback has no such entry); a failing estimate SKIPS the variant. */
int orc_bc1_transform_auto_with_normalization(const uint8_t *in, uint8_t *out, size_t len, int use_all, orc_estimate_fn est,
                                              void *ctx, int *out_norm, int *out_variant, int *out_split) {
    if (!est) est = ltu_cb;
    uint8_t *nb[3], *sb[3];
    for (int m = 0; m < 3; m++) nb[m] = (uint8_t *)malloc(len + 8), sb[m] = (uint8_t *)malloc(len + 8);
    uint8_t *work = (uint8_t *)malloc(len / 2 + 8);
    const int any = orc_bc1_normalize_blocks_all_modes(in, nb[0], nb[1], nb[2], len);
    int rc = 0;
    if (any) {
        int best_norm = ORC_NORM_NONE, best_variant = ORC_VARIANT_1, best_split = 1; /* Default of the details struct */
        size_t best = (size_t)-1;
        for (int m = 0; m < 3; m++) orc_bc1_transform(nb[m], sb[m], len, ORC_VARIANT_NONE, 0); /* plain block split */
        const int k = use_all ? 8 : 4;
        for (int m = 0; m < 3; m++)
            for (int i = 0; i < k; i++) {
                const int variant = use_all ? BC12_ALL[i][0] : BC12_FAST[i][0], split = use_all ? BC12_ALL[i][1] : BC12_FAST[i][1];
                if (split) orc_split_color_endpoints(sb[m], work, len / 2);
                else memcpy(work, sb[m], len / 2);
                decorrelate_in_place(work, len / 4, variant);
                size_t size = 0;
                if (est(ctx, work, len / 2, &size) != 0) continue; /* skip this variant */
                if (size < best) best = size, best_norm = m, best_variant = variant, best_split = split;
            }
        orc_bc1_transform_with_normalize_blocks(in, out, len, best_norm, best_variant, best_split);
        *out_norm = best_norm, *out_variant = best_variant, *out_split = best_split;
    } else {
        *out_norm = ORC_NORM_NONE;
        rc = orc_bc1_transform_auto(in, out, len, use_all, est, ctx, out_variant, out_split);
    }
    for (int m = 0; m < 3; m++) free(nb[m]), free(sb[m]);
    free(work);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Multi-threaded driver (CPU baseline): contiguous block ranges, one per thread.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int format, direction, variant, split_a, split_c;
    const uint8_t *in;
    uint8_t *out;
    size_t n, b0, b1;
    int isa;   /* baseline_isa() */
} mt_job;

static void *mt_worker(void *arg) {
    mt_job *j = (mt_job *)arg;
    switch (j->format) {
    case 1:
        if (j->isa == 5) bc1_range_avx512(j->direction, j->in, j->out, j->n, j->b0, j->b1, j->variant, j->split_c);
        else if (j->isa == 2) bc1_range_avx2(j->direction, j->in, j->out, j->n, j->b0, j->b1, j->variant, j->split_c);
        else bc1_range_fast(j->direction, j->in, j->out, j->n, j->b0, j->b1, j->variant, j->split_c);
        break;
    case 2:
        if (j->isa == 5) bc2_range_avx512(j->direction, j->in, j->out, j->n, j->b0, j->b1, j->variant, j->split_c);
        else bc2_range(j->direction, j->in, j->out, j->n, j->b0, j->b1, j->variant, j->split_c);
        break;
    default:
        if (j->isa == 5) bc3_range_avx512(j->direction, j->in, j->out, j->n, j->b0, j->b1, j->variant, j->split_a, j->split_c);
        else bc3_range(j->direction, j->in, j->out, j->n, j->b0, j->b1, j->variant, j->split_a, j->split_c);
        break;
    }
    return NULL;
}

/* One contiguous block range [b0, b1) of a payload of len bytes (what one shard / one GPU owns). */
void orc_bcn_run_range(int format, int direction, const uint8_t *in, uint8_t *out, size_t len, int variant,
                       int split_a, int split_c, size_t b0, size_t b1) {
    mt_job j = {format, direction, variant, split_a, split_c, in, out, len / (format == 1 ? 8 : 16), b0, b1, baseline_isa()};
    mt_worker(&j);
}

void orc_bcn_run_mt(int format, int direction, const uint8_t *in, uint8_t *out, size_t len, int variant, int split_a,
                    int split_c, int threads) {
    size_t n = len / (format == 1 ? 8 : 16);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    mt_job jobs[256];
    for (int t = 0; t < threads; t++) {
        jobs[t] = (mt_job){format, direction, variant, split_a, split_c, in, out, n, n * (size_t)t / (size_t)threads,
                           n * (size_t)(t + 1) / (size_t)threads, baseline_isa()};
        if (t + 1 < threads) pthread_create(&tid[t], NULL, mt_worker, &jobs[t]);
    }
    mt_worker(&jobs[threads - 1]);
    for (int t = 0; t + 1 < threads; t++) pthread_join(tid[t], NULL);
}
