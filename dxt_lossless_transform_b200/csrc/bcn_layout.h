// bcn_layout.h — stream layout of the transformed BCn formats (host + device).
//
// A transformed payload of N blocks is a concatenation of per-field STREAMS.  Stream s holds one
// w(s)-byte element per block and, in the reference's single-buffer layout, starts at byte
// N * sum_{s' < s} w(s').  This is exactly what the reference dispatchers compute:
//   BC1  transform_with_settings.rs:31-72   (core/dxt-lossless-transform-bc1)
//   BC2  transform_with_settings.rs:30-74   (core/dxt-lossless-transform-bc2)
//   BC3  transform_with_settings.rs:32-141  (core/dxt-lossless-transform-bc3)
//
//   BC1          : [c0c1:4 | c0:2, c1:2]                                  idx:4
//   BC2          : alpha:8, [c0c1:4 | c0:2, c1:2]                         idx:4
//   BC3          : [a0a1:2 | a0:1, a1:1], aidx:6, [c0c1:4 | c0:2, c1:2]   idx:4
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
#define DLT_HD __host__ __device__
#else
#define DLT_HD
#endif

namespace dlt {

constexpr int kMaxStreams = 6;

// Internal numbering (common/src/color_565/decorrelate.rs:72-84).
enum Variant : int { kNone = 0, kVariant1 = 1, kVariant2 = 2, kVariant3 = 3 };

// experimental::normalize_blocks (BC1 only) — ColorNormalizationMode::all_values() order
// (core/dxt-lossless-transform-bc1/src/experimental/normalize_blocks/normalize.rs:487-500).
enum NormalizeMode : int {
    kNormNone = 0,
    kNormColor0Only = 1,
    kNormReplicateColor = 2,
    // internal: what normalize_blocks_all_modes writes into its `None` buffer — transparent blocks become 0xFF, solid
    // blocks stay as they are (normalize.rs:449-466).  transform_bc1_auto_with_normalization ESTIMATES its `None`
    // candidates on that buffer (transform.rs:250-291) while the final transform with mode None leaves every block
    // untouched (normalize_split_blocks_in_place returns early, normalize.rs:293).
    kNormAllModesNone = 3,
};

struct Settings {
    int format;        // 1, 2, 3
    int variant;       // Variant
    bool split_alpha;  // BC3 only
    bool split_colour;
    int normalize = kNormNone;  // BC1 transform only: blocks are normalized on their way into the transform
};

DLT_HD constexpr int block_bytes(int fmt) { return fmt == 1 ? 8 : 16; }

DLT_HD constexpr int num_streams(int fmt, bool sa, bool sc) {
    return fmt == 1 ? 2 + (sc ? 1 : 0) : fmt == 2 ? 3 + (sc ? 1 : 0) : 4 + (sa ? 1 : 0) + (sc ? 1 : 0);
}

// Width in bytes of stream s.
DLT_HD constexpr int stream_width(int fmt, bool sa, bool sc, int s) {
    if (fmt == 1) return sc ? (s < 2 ? 2 : 4) : 4;
    if (fmt == 2) return s == 0 ? 8 : (sc ? (s < 3 ? 2 : 4) : 4);
    // BC3
    int na = sa ? 2 : 1;
    if (s < na) return sa ? 1 : 2;
    if (s == na) return 6;
    int c = s - na - 1;  // index within the colour part
    return sc ? (c < 2 ? 2 : 4) : 4;
}

// Byte offset per block of stream s in the reference layout (multiply by N).
DLT_HD constexpr int stream_prefix(int fmt, bool sa, bool sc, int s) {
    int acc = 0;
    for (int i = 0; i < s; i++) acc += stream_width(fmt, sa, sc, i);
    return acc;
}

// Natural alignment a stream base must have for the tiled kernels (element-typed smem accesses).
DLT_HD constexpr int stream_align(int w) { return w == 6 ? 2 : w; }

// Per-stream base pointers, already advanced to the first block of the range a launch covers.
struct StreamPtrs {
    uint8_t* p[kMaxStreams];
};

// Reference single-buffer layout: stream pointers for blocks [first, ...) of a payload of n blocks.
inline StreamPtrs reference_layout(uint8_t* base, size_t n, size_t first, const Settings& st) {
    StreamPtrs r{};
    int ns = num_streams(st.format, st.split_alpha, st.split_colour);
    for (int s = 0; s < ns; s++) {
        size_t w = (size_t)stream_width(st.format, st.split_alpha, st.split_colour, s);
        r.p[s] = base + n * (size_t)stream_prefix(st.format, st.split_alpha, st.split_colour, s) + w * first;
    }
    return r;
}

}  // namespace dlt
