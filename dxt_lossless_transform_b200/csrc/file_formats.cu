// file_formats.cu — the callers on either side of the block transform:
//
//   * TransformHeader / TransformFormat and the BC1 / BC2 embeddable details
//       api/dxt-lossless-transform-file-formats-api/src/embed/{mod.rs,transform_format.rs,formats/bc1.rs,formats/bc2.rs}
//   * TransformBundle + dispatch_transform / dispatch_untransform
//       .../src/bundle/{mod.rs,bc1.rs,bc2.rs}, .../src/handlers/dispatch.rs
//   * the DDS container: is_dds / parse_dds (the reference's own C exports) and DdsHandler
//       extensions/file-formats/dxt-lossless-transform-dds/src/dds/{parse_dds.rs,likely_dds.rs,constants.rs,exports.rs}
//       .../src/handler/{file_format_handler.rs,format_conversion.rs,file_format_detection.rs,
//                        file_format_untransform_detection.rs}
//   * a batched DDS entry point: every texture payload of a directory goes through ONE pinned copy
//     pipeline per GPU (the reference's CLI runs one rayon task per file,
//     tools/dxt-lossless-transform-cli/src/commands/transform/mod.rs:154-176).
//
// All of this is host logic (a few dozen integer operations per file); the payload bytes only ever
// move through the CUDA path of cabi.cu / host_pipeline.cu.  Files transformed here carry the same 4-byte
// header as the reference's, so either side can untransform the other's output.
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "bcn_layout.h"
#include "cabi_internal.h"

using namespace dlt;
using namespace dlt::cabi;

#define DLT_EXPORT extern "C" __attribute__((visibility("default")))

extern "C" {

// embed/transform_format.rs:10-32
enum DltffTransformFormat : int32_t {
    kFmtBc1 = 0,
    kFmtBc2 = 1,
    kFmtBc3 = 2,
    kFmtBc7 = 3,
    kFmtBc6H = 4,
    kFmtRgba8888 = 5,
    kFmtBgra8888 = 6,
    kFmtBgr888 = 7,
    kFmtBc4 = 8,
    kFmtBc5 = 9,
};

// TransformError / FormatHandlerError / EmbedError flattened (error.rs:19-87, embed/embed_error.rs:7-15).
enum DltffErrorCode : int32_t {
    kFfSuccess = 0,
    kFfEmbedCorruptedEmbeddedData = 1,
    kFfEmbedUnknownFormat = 2,
    kFfUnknownFileFormat = 3,
    kFfInvalidInputFileHeader = 4,
    kFfInvalidRestoredFileHeader = 5,
    kFfFormatNotImplemented = 6,                // detail_a = TransformFormat
    kFfNoBuilderForFormat = 7,                  // detail_a = TransformFormat
    kFfOutputBufferTooSmall = 8,                // detail_a = required, detail_b = actual
    kFfInputTooShort = 9,                       // detail_a = required, detail_b = actual
    kFfInputTooShortForStatedTextureSize = 10,  // detail_a = required, detail_b = actual
    kFfBc1 = 11,                                // detail_a = Dltbc1ErrorCode (stable), detail_b = its payload
    kFfBc2 = 12,                                // detail_a = Dltbc2ErrorCode (stable), detail_b = its payload
    kFfUnknownTransformFormat = 13,
    kFfInvalidDataAlignment = 14,               // detail_a = size, detail_b = required_divisor
    kFfNoSupportedHandler = 15,
    kFfNullPointer = 16,                        // C ABI only
};

struct DltffResult {
    int32_t error_code;
    size_t detail_a;
    size_t detail_b;
};

// dds/parse_dds.rs:7-42 (repr(u8) enum, repr(C) struct)
enum DdsFormat : uint8_t {
    kDdsNotADds = 0,
    kDdsUnknown = 1,
    kDdsBC1 = 2,
    kDdsBC2 = 3,
    kDdsBC3 = 4,
    kDdsBC6H = 5,
    kDdsBC7 = 6,
    kDdsRGBA8888 = 7,
    kDdsBGRA8888 = 8,
    kDdsBGR888 = 9,
    kDdsBC4 = 10,
    kDdsBC5 = 11,
};
struct DdsInfo {
    uint8_t format;
    uint8_t data_offset;
    uint32_t data_length;
};

struct DltddsFile {
    const uint8_t* input;
    size_t input_len;
    uint8_t* output;
    size_t output_len;
};

int dltcuda_transform_batch(const DltcudaPayload* payloads, size_t count, bool untransform);
int dltcuda_transform_batch_multi_gpu(const DltcudaPayload* payloads, size_t count, bool untransform,
                                      const int* devices, int num_devices);
int dltcuda_transform_auto_batch_multi_gpu(DltcudaAutoJob* jobs, size_t count, bool use_all_modes, const int* devices,
                                           int num_devices);
}

namespace {

constexpr DltffResult kOk{kFfSuccess, 0, 0};
inline DltffResult err(int32_t code, size_t a = 0, size_t b = 0) { return DltffResult{code, a, b}; }

// ------------------------------------------------------------------------------------------------
// TransformHeader (embed/mod.rs:107-160): u32 little endian, bits 0-3 format, bits 4-31 data.
// ------------------------------------------------------------------------------------------------
constexpr size_t kTransformHeaderSize = 4;

inline uint32_t header_new(uint32_t format, uint32_t data) { return (format & 0xFu) | (data << 4); }
inline uint32_t header_format_raw(uint32_t h) { return h & 0xFu; }
inline uint32_t header_format_data(uint32_t h) { return h >> 4; }
inline bool format_known(uint32_t raw) { return raw <= kFmtBc5; }  // transform_format.rs:39-55

inline uint32_t read_le32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
inline void write_le32(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)v, p[1] = (uint8_t)(v >> 8), p[2] = (uint8_t)(v >> 16), p[3] = (uint8_t)(v >> 24);
}

// BC1 / BC2 format data (embed/formats/bc1.rs:33-56, bc2.rs:30-52): bits 0-1 version (0), bit 2 split
// colour endpoints, bits 3-4 decorrelation variant in the STABLE numbering, the rest reserved (0).
inline uint32_t pack_bc12(int variant_internal, bool split_colour) {
    return 0u | ((split_colour ? 1u : 0u) << 2) | ((uint32_t)internal_to_stable(variant_internal) << 3);
}
inline bool unpack_bc12(uint32_t data, int* variant_internal, bool* split_colour) {
    if ((data & 3u) != 0) return false;  // only InitialVersion = 0 is valid -> CorruptedEmbeddedData
    *split_colour = (data >> 2) & 1u;
    *variant_internal = stable_to_internal((uint8_t)((data >> 3) & 3u));
    return true;
}

// ------------------------------------------------------------------------------------------------
// TransformBundle (bundle/mod.rs:37-49): an optional manual-or-auto builder for BC1 and for BC2.
// ------------------------------------------------------------------------------------------------
struct BundleSlot {
    enum Kind { kNone, kManual, kAuto } kind = kNone;
    ManualBuilder manual{};
    AutoBuilder automatic{};
};
struct Bundle {
    BundleSlot bc1, bc2;
};

// Bc1Builder::transform_slice_with_details (bundle/bc1.rs:45-71): the settings used, or the builder's error.
DltffResult slot_transform(const BundleSlot& slot, int format, const uint8_t* in, size_t in_len, uint8_t* out,
                           size_t out_len, Settings* used) {
    const int32_t wrap = format == 1 ? kFfBc1 : kFfBc2;
    DltResult r;
    if (slot.kind == BundleSlot::kManual) {
        ManualBuilder b = slot.manual;
        *used = Settings{format, b.variant, false, b.split_colour};
        r = api_manual_run(format, false, in, in_len, out, out_len, &b);
    } else {
        r = api_auto_settings(format, &slot.automatic, in, in_len, out, out_len, used);
    }
    if (r.error_code == kApiSuccess) return kOk;
    // Bc1Error::InvalidLength(len) carries the length; OutputBufferTooSmall cannot happen here (checked above).
    return err(wrap, (size_t)r.error_code, r.error_code == kApiInvalidLength ? in_len : 0);
}

// TransformBundle::dispatch_transform (bundle/mod.rs:125-174)
DltffResult bundle_dispatch_transform(const Bundle& bundle, uint32_t format, const uint8_t* in, size_t in_len,
                                      uint8_t* out, size_t out_len, uint32_t* out_header) {
    if (out_len < in_len) return err(kFfOutputBufferTooSmall, in_len, out_len);
    if (format != kFmtBc1 && format != kFmtBc2) return err(kFfUnknownTransformFormat);
    const BundleSlot& slot = format == kFmtBc1 ? bundle.bc1 : bundle.bc2;
    if (slot.kind == BundleSlot::kNone) return err(kFfNoBuilderForFormat, format);
    Settings used{};
    const DltffResult r = slot_transform(slot, format == kFmtBc1 ? 1 : 2, in, in_len, out, out_len, &used);
    if (r.error_code != kFfSuccess) return r;
    *out_header = header_new(format, pack_bc12(used.variant, used.split_colour));
    return kOk;
}

// What dispatch_untransform (handlers/dispatch.rs:41-101) decides before it touches the data.
DltffResult plan_untransform(uint32_t header, size_t in_len, size_t out_len, Settings* st) {
    if (out_len < in_len) return err(kFfOutputBufferTooSmall, in_len, out_len);
    const uint32_t raw = header_format_raw(header);
    if (raw != kFmtBc1 && raw != kFmtBc2) return err(kFfUnknownTransformFormat);
    int variant = 0;
    bool split = false;
    if (!unpack_bc12(header_format_data(header), &variant, &split)) return err(kFfEmbedCorruptedEmbeddedData);
    const int format = raw == kFmtBc1 ? 1 : 2;
    const size_t div = (size_t)block_bytes(format);
    if (in_len % div) return err(kFfInvalidDataAlignment, in_len, div);
    *st = Settings{format, variant, false, split};
    return kOk;
}

DltffResult run_untransform(const Settings& st, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    ManualBuilder b{st.format, st.variant, st.split_colour};
    const DltResult r = api_manual_run(st.format, true, in, in_len, out, out_len, &b);
    if (r.error_code == kApiSuccess) return kOk;
    // The reference calls the unsafe untransform here, which cannot fail; a CUDA failure has to surface.
    return err(st.format == 1 ? kFfBc1 : kFfBc2, (size_t)r.error_code, 0);
}

// ------------------------------------------------------------------------------------------------
// DDS header parsing (dds/constants.rs, dds/parse_dds.rs)
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kDdsMagic = 0x20534444u;  // "DDS "
constexpr size_t kDdsHeaderSize = 0x80, kDx10HeaderSize = 20;
constexpr size_t kOffFlags = 0x08, kOffHeight = 0x0C, kOffWidth = 0x10, kOffMipCount = 0x1C;
constexpr size_t kOffPfFlags = 0x50, kOffFourCC = 0x54, kOffRgbBitCount = 0x58;
constexpr size_t kOffRMask = 0x5C, kOffGMask = 0x60, kOffBMask = 0x64, kOffAMask = 0x68, kOffDxgi = 0x80;
constexpr uint32_t kDdsdMipmapCount = 0x20000u;
constexpr uint32_t kDdpfAlphaPixels = 0x1u, kDdpfAlpha = 0x2u, kDdpfFourCC = 0x4u, kDdpfRgb = 0x40u, kDdpfYuv = 0x200u,
                   kDdpfLuminance = 0x20000u;

constexpr uint32_t fourcc(char a, char b, char c, char d) {
    return (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8) | ((uint32_t)(uint8_t)c << 16) | ((uint32_t)(uint8_t)d << 24);
}

bool likely_dds(const uint8_t* d, size_t len) {  // likely_dds.rs:9-13
    return len >= kDdsHeaderSize && read_le32(d) == kDdsMagic;
}

uint8_t dxgi_to_format(uint32_t f) {  // parse_dds.rs:99-134
    if (f >= 70 && f <= 72) return kDdsBC1;
    if (f >= 73 && f <= 75) return kDdsBC2;
    if (f >= 76 && f <= 78) return kDdsBC3;
    if (f >= 79 && f <= 81) return kDdsBC4;
    if (f >= 82 && f <= 84) return kDdsBC5;
    if (f >= 94 && f <= 96) return kDdsBC6H;
    if (f >= 97 && f <= 99) return kDdsBC7;
    if (f >= 27 && f <= 32) return kDdsRGBA8888;
    if (f == 87 || f == 90 || f == 91) return kDdsBGRA8888;
    return kDdsUnknown;
}

uint8_t detect_uncompressed(const uint8_t* d) {  // parse_dds.rs:178-237
    const uint32_t flags = read_le32(d + kOffPfFlags), bits = read_le32(d + kOffRgbBitCount);
    const uint32_t r = read_le32(d + kOffRMask), g = read_le32(d + kOffGMask), b = read_le32(d + kOffBMask),
                   a = read_le32(d + kOffAMask);
    if (bits == 24) return (r == 0x00FF0000u && g == 0x0000FF00u && b == 0x000000FFu && a == 0) ? kDdsBGR888 : kDdsUnknown;
    if (bits == 32 && (flags & kDdpfAlphaPixels)) {
        if (r == 0x000000FFu && g == 0x0000FF00u && b == 0x00FF0000u && a == 0xFF000000u) return kDdsRGBA8888;
        if (r == 0x00FF0000u && g == 0x0000FF00u && b == 0x000000FFu && a == 0xFF000000u) return kDdsBGRA8888;
    }
    return kDdsUnknown;
}

inline uint32_t sat_add(uint32_t a, uint32_t b) { return a + b < a ? 0xFFFFFFFFu : a + b; }
inline uint32_t div_ceil4(uint32_t v) { return v / 4 + (v % 4 != 0); }

// parse_dds.rs:283-340 / 374-400: u32 arithmetic (products wrap as in a release build, the sum saturates).
// `level(w, h)` is the size of one mip level.  Once the chain has reached 1x1 every further level adds the
// same amount, so the tail is closed-form (a header may claim 2^32-1 levels; the reference loops over them).
template <typename LevelFn>
uint32_t mip_chain_bytes(uint32_t w, uint32_t h, uint32_t mips, LevelFn level) {
    uint32_t total = 0;
    for (uint32_t i = 0; i < mips; i++) {
        if (w == 1 && h == 1) {
            const uint64_t tail = (uint64_t)level(1u, 1u) * (uint64_t)(mips - i) + total;
            return tail > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)tail;
        }
        total = sat_add(total, level(w, h));
        w = w / 2 > 1 ? w / 2 : 1;
        h = h / 2 > 1 ? h / 2 : 1;
    }
    return total;
}
uint32_t mip_chain_bytes_blocks(uint32_t w, uint32_t h, uint32_t mips, uint32_t block_size) {
    return mip_chain_bytes(w, h, mips, [=](uint32_t lw, uint32_t lh) { return div_ceil4(lw) * div_ceil4(lh) * block_size; });
}
uint32_t mip_chain_bytes_pixels(uint32_t w, uint32_t h, uint32_t mips, uint32_t bpp) {
    return mip_chain_bytes(w, h, mips, [=](uint32_t lw, uint32_t lh) { return lw * lh * bpp; });
}

uint32_t data_length_of(uint8_t format, const uint8_t* d) {  // calculate_data_length, parse_dds.rs:241-281
    const uint32_t flags = read_le32(d + kOffFlags), height = read_le32(d + kOffHeight), width = read_le32(d + kOffWidth);
    const uint32_t raw_mips = read_le32(d + kOffMipCount);
    const uint32_t mips = (flags & kDdsdMipmapCount) ? (raw_mips > 1 ? raw_mips : 1) : 1;
    switch (format) {
        case kDdsBC1:
        case kDdsBC4: return mip_chain_bytes_blocks(width, height, mips, 8);
        case kDdsBC2:
        case kDdsBC3:
        case kDdsBC5:
        case kDdsBC6H:
        case kDdsBC7: return mip_chain_bytes_blocks(width, height, mips, 16);
        case kDdsRGBA8888:
        case kDdsBGRA8888: return mip_chain_bytes_pixels(width, height, mips, 4);
        case kDdsBGR888: return mip_chain_bytes_pixels(width, height, mips, 3);
        case kDdsUnknown: {  // calculate_uncompressed_data_length, parse_dds.rs:343-371
            const uint32_t pf = read_le32(d + kOffPfFlags), bits = read_le32(d + kOffRgbBitCount);
            if ((pf & (kDdpfRgb | kDdpfLuminance | kDdpfYuv | kDdpfAlpha)) == 0) return 0;
            if (bits % 8 != 0 || bits / 8 == 0) return 0;
            return mip_chain_bytes_pixels(width, height, mips, bits / 8);
        }
        default: return 0;
    }
}

// parse_dds_ignore_magic (parse_dds.rs:78-172).  false = None.
bool parse_ignore_magic(const uint8_t* d, size_t len, DdsInfo* info) {
    if (len < kDdsHeaderSize) return false;
    const uint32_t cc = read_le32(d + kOffFourCC);
    uint8_t format;
    size_t offset;
    if (cc == fourcc('D', 'X', '1', '0')) {
        if (len < kDdsHeaderSize + kDx10HeaderSize) return false;
        format = dxgi_to_format(read_le32(d + kOffDxgi));
        offset = kDdsHeaderSize + kDx10HeaderSize;
    } else {
        const uint32_t pf = read_le32(d + kOffPfFlags);
        if (pf & kDdpfFourCC) {
            if (cc == fourcc('D', 'X', 'T', '1')) format = kDdsBC1;
            else if (cc == fourcc('D', 'X', 'T', '2') || cc == fourcc('D', 'X', 'T', '3')) format = kDdsBC2;
            else if (cc == fourcc('D', 'X', 'T', '4') || cc == fourcc('D', 'X', 'T', '5')) format = kDdsBC3;
            else if (cc == fourcc('B', 'C', '4', 'U') || cc == fourcc('B', 'C', '4', 'S') || cc == fourcc('A', 'T', 'I', '1'))
                format = kDdsBC4;
            else if (cc == fourcc('B', 'C', '5', 'U') || cc == fourcc('B', 'C', '5', 'S') || cc == fourcc('A', 'T', 'I', '2'))
                format = kDdsBC5;
            else format = kDdsUnknown;
        } else if (pf & kDdpfRgb) {
            format = detect_uncompressed(d);
        } else {
            format = kDdsUnknown;
        }
        offset = kDdsHeaderSize;
    }
    info->format = format;
    info->data_offset = (uint8_t)offset;
    info->data_length = data_length_of(format, d);
    return true;
}

bool parse_with_magic(const uint8_t* d, size_t len, DdsInfo* info) {  // parse_dds.rs:58-64
    return likely_dds(d, len) && parse_ignore_magic(d, len, info);
}

// dds_format_to_transform_format(format, allow_unimplemented = false) — handler/format_conversion.rs:45-107
DltffResult dds_to_transform_format(uint8_t dds_format, uint32_t* out) {
    switch (dds_format) {
        case kDdsBC1: *out = kFmtBc1; return kOk;
        case kDdsBC2: *out = kFmtBc2; return kOk;
        case kDdsBC3: return err(kFfFormatNotImplemented, kFmtBc3);
        case kDdsBC4: return err(kFfFormatNotImplemented, kFmtBc4);
        case kDdsBC5: return err(kFfFormatNotImplemented, kFmtBc5);
        case kDdsBC6H: return err(kFfFormatNotImplemented, kFmtBc6H);
        case kDdsBC7: return err(kFfFormatNotImplemented, kFmtBc7);
        case kDdsRGBA8888: *out = kFmtRgba8888; return kOk;  // known to the handler, rejected by the bundle
        case kDdsBGRA8888: *out = kFmtBgra8888; return kOk;
        case kDdsBGR888: *out = kFmtBgr888; return kOk;
        default: return err(kFfUnknownFileFormat);
    }
}

bool extension_ok(const char* ext) { return !ext || std::strcmp(ext, "dds") == 0; }

// ------------------------------------------------------------------------------------------------
// DdsHandler (handler/file_format_handler.rs): everything that happens before / after the payload call.
// ------------------------------------------------------------------------------------------------
struct DdsPlan {
    size_t data_offset = 0, data_length = 0;
    uint32_t format = 0;   // TransformFormat (transform direction)
    uint32_t header = 0;   // TransformHeader read from the file (untransform direction)
};

// transform_bundle :17-60 up to the dispatch call.
DltffResult dds_plan_transform(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, DdsPlan* plan) {
    if (out_len < in_len) return err(kFfOutputBufferTooSmall, in_len, out_len);
    DdsInfo info{};
    if (!parse_with_magic(in, in_len, &info)) return err(kFfInvalidInputFileHeader);
    plan->data_offset = info.data_offset;
    plan->data_length = info.data_length;
    const size_t total = plan->data_offset + plan->data_length;
    if (in_len < total) return err(kFfInputTooShortForStatedTextureSize, total, in_len);
    std::memcpy(out, in, plan->data_offset);
    return dds_to_transform_format(info.format, &plan->format);
}
// transform_bundle :62-84 after the dispatch call.
void dds_finish_transform(const uint8_t* in, size_t in_len, uint8_t* out, const DdsPlan& plan, uint32_t header) {
    const size_t rest = plan.data_offset + plan.data_length;
    if (in_len > rest) std::memcpy(out + rest, in + rest, in_len - rest);
    write_le32(out, header);
}

// untransform :88-131 up to the dispatch call.
DltffResult dds_plan_untransform(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, DdsPlan* plan) {
    if (in_len < kTransformHeaderSize) return err(kFfInputTooShort, kTransformHeaderSize, in_len);
    if (out_len < in_len) return err(kFfOutputBufferTooSmall, in_len, out_len);
    plan->header = read_le32(in);
    DdsInfo info{};
    if (!parse_ignore_magic(in, in_len, &info)) return err(kFfInvalidRestoredFileHeader);
    plan->data_offset = info.data_offset;
    plan->data_length = info.data_length;
    const size_t total = plan->data_offset + plan->data_length;
    if (in_len < total) return err(kFfInputTooShortForStatedTextureSize, total, in_len);
    write_le32(out, kDdsMagic);
    std::memcpy(out + 4, in + 4, plan->data_offset - 4);
    return kOk;
}
void dds_finish_untransform(const uint8_t* in, size_t in_len, uint8_t* out, const DdsPlan& plan) {
    const size_t rest = plan.data_offset + plan.data_length;
    if (in_len > rest) std::memcpy(out + rest, in + rest, in_len - rest);
}

DltcudaSettings to_cuda_settings(const Settings& st) {
    return DltcudaSettings{(uint8_t)st.format, (uint8_t)st.variant, st.split_alpha, st.split_colour};
}

int run_batch(const std::vector<DltcudaPayload>& payloads, bool untransform, const int* devices, int num_devices) {
    if (payloads.empty()) return 0;
    if (devices && num_devices > 0)
        return dltcuda_transform_batch_multi_gpu(payloads.data(), payloads.size(), untransform, devices, num_devices);
    return dltcuda_transform_batch(payloads.data(), payloads.size(), untransform);
}

}  // namespace

// =================================================================================================
// TransformHeader helpers
// =================================================================================================
DLT_EXPORT uint32_t dltff_TransformHeader_new(int32_t format, uint32_t data) { return header_new((uint32_t)format, data); }
// false when the 4-bit format value is not a known TransformFormat (TransformHeader::format() == None).
DLT_EXPORT bool dltff_TransformHeader_format(uint32_t header, int32_t* out_format) {
    const uint32_t raw = header_format_raw(header);
    if (out_format) *out_format = (int32_t)raw;
    return format_known(raw);
}
DLT_EXPORT uint32_t dltff_TransformHeader_format_data(uint32_t header) { return header_format_data(header); }
DLT_EXPORT uint32_t dltff_TransformHeader_read(const uint8_t* ptr) { return ptr ? read_le32(ptr) : 0; }
DLT_EXPORT void dltff_TransformHeader_write(uint32_t header, uint8_t* ptr) {
    if (ptr) write_le32(ptr, header);
}

// EmbeddableBc{1,2}Details::from_settings(..).to_header() — decorrelation_mode in the STABLE numbering.
DLT_EXPORT uint32_t dltff_bc1_header_from_settings(uint8_t decorrelation_mode, bool split_colour_endpoints) {
    return header_new(kFmtBc1, pack_bc12(stable_to_internal(decorrelation_mode & 3), split_colour_endpoints));
}
DLT_EXPORT uint32_t dltff_bc2_header_from_settings(uint8_t decorrelation_mode, bool split_colour_endpoints) {
    return header_new(kFmtBc2, pack_bc12(stable_to_internal(decorrelation_mode & 3), split_colour_endpoints));
}
// EmbeddableBc{1,2}Details::from_header (formats/mod.rs:43-49): wrong or unknown format -> UnknownFormat,
// bad version -> CorruptedEmbeddedData.
static DltffResult settings_from_header(uint32_t want, uint32_t header, uint8_t* out_mode, bool* out_split) {
    if (!out_mode || !out_split) return err(kFfNullPointer);
    if (header_format_raw(header) != want) return err(kFfEmbedUnknownFormat);
    int variant = 0;
    bool split = false;
    if (!unpack_bc12(header_format_data(header), &variant, &split)) return err(kFfEmbedCorruptedEmbeddedData);
    *out_mode = internal_to_stable(variant);
    *out_split = split;
    return kOk;
}
DLT_EXPORT DltffResult dltff_bc1_settings_from_header(uint32_t header, uint8_t* out_mode, bool* out_split) {
    return settings_from_header(kFmtBc1, header, out_mode, out_split);
}
DLT_EXPORT DltffResult dltff_bc2_settings_from_header(uint32_t header, uint8_t* out_mode, bool* out_split) {
    return settings_from_header(kFmtBc2, header, out_mode, out_split);
}

DLT_EXPORT const char* dltff_error_message(int32_t code) {  // the #[error(..)] strings of error.rs / embed_error.rs
    switch (code) {
        case kFfSuccess: return "Success";
        case kFfEmbedCorruptedEmbeddedData: return "Embed error: Corrupted embedded data. Info about the transform stored is invalid.";
        case kFfEmbedUnknownFormat: return "Embed error: Unknown transform format: header contains unrecognized format value";
        case kFfUnknownFileFormat: return "Format handler error: Unknown file format";
        case kFfInvalidInputFileHeader: return "Format handler error: Invalid input file header during transform";
        case kFfInvalidRestoredFileHeader:
            return "Format handler error: Invalid restored file header during untransform - file may be corrupted or wrong handler used";
        case kFfFormatNotImplemented: return "Format handler error: format not yet implemented";
        case kFfNoBuilderForFormat: return "Format handler error: No transform builder provided for format";
        case kFfOutputBufferTooSmall: return "Format handler error: Output buffer too small";
        case kFfInputTooShort: return "Format handler error: Input buffer too short";
        case kFfInputTooShortForStatedTextureSize:
            return "Format handler error: Input buffer too short for stated texture size in header";
        case kFfBc1: return "BC1 transform error";
        case kFfBc2: return "BC2 transform error";
        case kFfUnknownTransformFormat: return "Unrecognized or unsupported transform format in header";
        case kFfInvalidDataAlignment: return "Invalid data alignment";
        case kFfNoSupportedHandler: return "No file format handler can process the file";
        case kFfNullPointer: return "Null pointer provided";
        default: return "Unknown error";
    }
}

// =================================================================================================
// TransformBundle
// =================================================================================================
DLT_EXPORT void* dltff_new_TransformBundle(void) { return new (std::nothrow) Bundle; }  // TransformBundle::new, mod.rs:72
DLT_EXPORT void* dltff_TransformBundle_default_all(void) {                              // default_all, mod.rs:185-192
    Bundle* b = new (std::nothrow) Bundle;
    if (!b) return nullptr;
    b->bc1.kind = b->bc2.kind = BundleSlot::kManual;
    b->bc1.manual = ManualBuilder{1, kVariant1, true};
    b->bc2.manual = ManualBuilder{2, kVariant1, true};
    return b;
}
DLT_EXPORT void dltff_free_TransformBundle(void* bundle) { delete static_cast<Bundle*>(bundle); }

// with_bcN_manual / with_bcN_auto (mod.rs:77-111).  The Rust methods take the builder by value; here the
// builder's state is copied and the caller keeps (and later frees) its builder.
static DltffResult set_slot(void* bundle, const void* builder, int format, bool manual) {
    if (!bundle || !builder) return err(kFfNullPointer);
    Bundle* b = static_cast<Bundle*>(bundle);
    BundleSlot& slot = format == 1 ? b->bc1 : b->bc2;
    if (manual) {
        const ManualBuilder* m = static_cast<const ManualBuilder*>(builder);
        if (m->format != format) return err(kFfUnknownTransformFormat);
        slot.kind = BundleSlot::kManual, slot.manual = *m;
    } else {
        const AutoBuilder* a = static_cast<const AutoBuilder*>(builder);
        if (a->format != format) return err(kFfUnknownTransformFormat);
        slot.kind = BundleSlot::kAuto, slot.automatic = *a;
    }
    return kOk;
}
DLT_EXPORT DltffResult dltff_TransformBundle_with_bc1_manual(void* bundle, const void* builder) { return set_slot(bundle, builder, 1, true); }
DLT_EXPORT DltffResult dltff_TransformBundle_with_bc1_auto(void* bundle, const void* builder) { return set_slot(bundle, builder, 1, false); }
DLT_EXPORT DltffResult dltff_TransformBundle_with_bc2_manual(void* bundle, const void* builder) { return set_slot(bundle, builder, 2, true); }
DLT_EXPORT DltffResult dltff_TransformBundle_with_bc2_auto(void* bundle, const void* builder) { return set_slot(bundle, builder, 2, false); }

// dispatch_transform (handlers/dispatch.rs:131-143)
DLT_EXPORT DltffResult dltff_dispatch_transform(int32_t format, const uint8_t* input, size_t input_len, uint8_t* output,
                                                size_t output_len, const void* bundle, uint32_t* out_header) {
    if (!bundle || !out_header || (input_len && (!input || !output))) return err(kFfNullPointer);
    return bundle_dispatch_transform(*static_cast<const Bundle*>(bundle), (uint32_t)format, input, input_len, output,
                                     output_len, out_header);
}
// dispatch_untransform (handlers/dispatch.rs:41-101)
DLT_EXPORT DltffResult dltff_dispatch_untransform(uint32_t header, const uint8_t* input, size_t input_len, uint8_t* output,
                                                  size_t output_len) {
    if (input_len && (!input || !output)) return err(kFfNullPointer);
    Settings st{};
    const DltffResult r = plan_untransform(header, input_len, output_len, &st);
    if (r.error_code != kFfSuccess) return r;
    return run_untransform(st, input, input_len, output, output_len);
}

// =================================================================================================
// DDS: the reference's own exports (dds/exports.rs:12-64) ...
// =================================================================================================
DLT_EXPORT bool is_dds(const uint8_t* ptr, size_t len) { return ptr && len && likely_dds(ptr, len); }
DLT_EXPORT DdsInfo parse_dds(const uint8_t* ptr, size_t len) {
    DdsInfo info{kDdsNotADds, 0, 0};
    if (!ptr || !len || !parse_with_magic(ptr, len, &info)) return DdsInfo{kDdsNotADds, 0, 0};
    return info;
}
// ... and the Rust-only rest of the handler.
DLT_EXPORT DdsInfo dltdds_parse_dds_ignore_magic(const uint8_t* ptr, size_t len) {
    DdsInfo info{kDdsNotADds, 0, 0};
    if (!ptr || !len || !parse_ignore_magic(ptr, len, &info)) return DdsInfo{kDdsNotADds, 0, 0};
    return info;
}
DLT_EXPORT bool dltdds_can_handle(const uint8_t* input, size_t len, const char* file_extension) {
    DdsInfo info{};
    return extension_ok(file_extension) && input && parse_with_magic(input, len, &info);
}
DLT_EXPORT bool dltdds_can_handle_untransform(const uint8_t* input, size_t len, const char* file_extension) {
    DdsInfo info{};
    return extension_ok(file_extension) && input && len >= 4 && parse_ignore_magic(input, len, &info);
}

DLT_EXPORT DltffResult dltdds_transform_bundle(const uint8_t* input, size_t input_len, uint8_t* output, size_t output_len,
                                               const void* bundle) {
    if (!bundle || (input_len && (!input || !output))) return err(kFfNullPointer);
    DdsPlan plan;
    DltffResult r = dds_plan_transform(input, input_len, output, output_len, &plan);
    if (r.error_code != kFfSuccess) return r;
    uint32_t header = 0;
    r = bundle_dispatch_transform(*static_cast<const Bundle*>(bundle), plan.format, input + plan.data_offset,
                                  plan.data_length, output + plan.data_offset, plan.data_length, &header);
    if (r.error_code != kFfSuccess) return r;
    dds_finish_transform(input, input_len, output, plan, header);
    return kOk;
}

DLT_EXPORT DltffResult dltdds_untransform(const uint8_t* input, size_t input_len, uint8_t* output, size_t output_len) {
    if (input_len && (!input || !output)) return err(kFfNullPointer);
    DdsPlan plan;
    DltffResult r = dds_plan_untransform(input, input_len, output, output_len, &plan);
    if (r.error_code != kFfSuccess) return r;
    Settings st{};
    r = plan_untransform(plan.header, plan.data_length, plan.data_length, &st);
    if (r.error_code != kFfSuccess) return r;
    r = run_untransform(st, input + plan.data_offset, plan.data_length, output + plan.data_offset, plan.data_length);
    if (r.error_code != kFfSuccess) return r;
    dds_finish_untransform(input, input_len, output, plan);
    return kOk;
}

// =================================================================================================
// Batched DDS: per-file results identical to the single-file calls; the payloads of all files whose
// builder is manual (and every untransform) share one copy pipeline per device.
// devices == NULL / num_devices == 0: the calling thread's device.
// =================================================================================================
// C++ exceptions (vector growth) must not cross the C ABI: the batch bodies run inside try / catch, 2 = out of memory.
static int transform_bundle_batch_body(const DltddsFile* files, size_t count, const void* bundle, DltffResult* results,
                                       const int* devices, int num_devices);
static int untransform_batch_body(const DltddsFile* files, size_t count, DltffResult* results, const int* devices,
                                  int num_devices);
DLT_EXPORT int dltdds_transform_bundle_batch(const DltddsFile* files, size_t count, const void* bundle, DltffResult* results,
                                             const int* devices, int num_devices) {
    try {
        return transform_bundle_batch_body(files, count, bundle, results, devices, num_devices);
    } catch (...) {
        return 2;
    }
}
DLT_EXPORT int dltdds_untransform_batch(const DltddsFile* files, size_t count, DltffResult* results, const int* devices,
                                        int num_devices) {
    try {
        return untransform_batch_body(files, count, results, devices, num_devices);
    } catch (...) {
        return 2;
    }
}
static int transform_bundle_batch_body(const DltddsFile* files, size_t count, const void* bundle, DltffResult* results,
                                       const int* devices, int num_devices) {
    if (count == 0) return 0;
    if (!files || !bundle || !results) return 1;
    const Bundle& b = *static_cast<const Bundle*>(bundle);
    std::vector<DdsPlan> plans(count);
    std::vector<Settings> used(count);
    std::vector<DltcudaPayload> payloads;
    std::vector<size_t> owner;
    std::vector<DltcudaAutoJob> auto_fast, auto_all;   // files whose builder is auto + GPU LTU, by search depth
    std::vector<size_t> owner_fast, owner_all;
    for (size_t i = 0; i < count; i++) {
        const DltddsFile& f = files[i];
        if (f.input_len && (!f.input || !f.output)) {
            results[i] = err(kFfNullPointer);
            continue;
        }
        results[i] = dds_plan_transform(f.input, f.input_len, f.output, f.output_len, &plans[i]);
        if (results[i].error_code != kFfSuccess) continue;
        const DdsPlan& p = plans[i];
        if (p.format != kFmtBc1 && p.format != kFmtBc2) {
            results[i] = err(kFfUnknownTransformFormat);
            continue;
        }
        const BundleSlot& slot = p.format == kFmtBc1 ? b.bc1 : b.bc2;
        const int fmt = p.format == kFmtBc1 ? 1 : 2;
        if (slot.kind == BundleSlot::kNone) {
            results[i] = err(kFfNoBuilderForFormat, p.format);
        } else if (p.data_length % (size_t)block_bytes(fmt)) {
            results[i] = err(fmt == 1 ? kFfBc1 : kFfBc2, kApiInvalidLength, p.data_length);
        } else if (slot.kind == BundleSlot::kAuto && is_gpu_ltu_estimator(slot.automatic.estimator)) {
            // GPU LTU estimator: the searches of all such files share ONE batched search (below)
            (slot.automatic.use_all ? auto_all : auto_fast).push_back(
                DltcudaAutoJob{(uint8_t)fmt, f.input + p.data_offset, f.output + p.data_offset, p.data_length, {}, 0});
            (slot.automatic.use_all ? owner_all : owner_fast).push_back(i);
        } else if (slot.kind == BundleSlot::kAuto) {
            // a caller-supplied estimator sees host memory: one search per file, as in the reference
            results[i] = slot_transform(slot, fmt, f.input + p.data_offset, p.data_length, f.output + p.data_offset,
                                        p.data_length, &used[i]);
            if (results[i].error_code == kFfSuccess)
                dds_finish_transform(f.input, f.input_len, f.output, p,
                                     header_new(p.format, pack_bc12(used[i].variant, used[i].split_colour)));
        } else {
            used[i] = Settings{fmt, slot.manual.variant, false, slot.manual.split_colour};
            payloads.push_back(DltcudaPayload{f.input + p.data_offset, f.output + p.data_offset, p.data_length,
                                              to_cuda_settings(used[i])});
            owner.push_back(i);
        }
    }
    const int rc = run_batch(payloads, false, devices, num_devices);
    for (size_t k = 0; k < owner.size(); k++) {
        const size_t i = owner[k];
        if (rc != 0) {
            results[i] = err(used[i].format == 1 ? kFfBc1 : kFfBc2, kApiAllocationFailed, 0);
            continue;
        }
        dds_finish_transform(files[i].input, files[i].input_len, files[i].output, plans[i],
                             header_new(plans[i].format, pack_bc12(used[i].variant, used[i].split_colour)));
    }
    for (int depth = 0; depth < 2; depth++) {
        std::vector<DltcudaAutoJob>& jobs = depth ? auto_all : auto_fast;
        const std::vector<size_t>& own = depth ? owner_all : owner_fast;
        if (jobs.empty()) continue;
        const int arc = devices && num_devices > 0
                            ? dltcuda_transform_auto_batch_multi_gpu(jobs.data(), jobs.size(), depth != 0, devices, num_devices)
                            : auto_batch_host(jobs.data(), jobs.size(), depth != 0);
        for (size_t k = 0; k < own.size(); k++) {
            const size_t i = own[k];
            if (arc != 0 || jobs[k].status != 0) {
                results[i] = err(jobs[k].format == 1 ? kFfBc1 : kFfBc2, kApiAllocationFailed, 0);
                continue;
            }
            dds_finish_transform(files[i].input, files[i].input_len, files[i].output, plans[i],
                                 header_new(plans[i].format, pack_bc12(jobs[k].out_settings.decorrelation_mode,
                                                                       jobs[k].out_settings.split_colour_endpoints)));
        }
    }
    return 0;
}

static int untransform_batch_body(const DltddsFile* files, size_t count, DltffResult* results, const int* devices,
                                  int num_devices) {
    if (count == 0) return 0;
    if (!files || !results) return 1;
    std::vector<DdsPlan> plans(count);
    std::vector<Settings> used(count);
    std::vector<DltcudaPayload> payloads;
    std::vector<size_t> owner;
    for (size_t i = 0; i < count; i++) {
        const DltddsFile& f = files[i];
        if (f.input_len && (!f.input || !f.output)) {
            results[i] = err(kFfNullPointer);
            continue;
        }
        results[i] = dds_plan_untransform(f.input, f.input_len, f.output, f.output_len, &plans[i]);
        if (results[i].error_code != kFfSuccess) continue;
        const DdsPlan& p = plans[i];
        results[i] = plan_untransform(p.header, p.data_length, p.data_length, &used[i]);
        if (results[i].error_code != kFfSuccess) continue;
        payloads.push_back(DltcudaPayload{f.input + p.data_offset, f.output + p.data_offset, p.data_length,
                                          to_cuda_settings(used[i])});
        owner.push_back(i);
    }
    const int rc = run_batch(payloads, true, devices, num_devices);
    for (size_t k = 0; k < owner.size(); k++) {
        const size_t i = owner[k];
        if (rc != 0) {
            results[i] = err(used[i].format == 1 ? kFfBc1 : kFfBc2, kApiAllocationFailed, 0);
            continue;
        }
        dds_finish_untransform(files[i].input, files[i].input_len, files[i].output, plans[i]);
    }
    return 0;
}
