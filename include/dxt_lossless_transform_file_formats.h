/*
 * dxt_lossless_transform_file_formats.h — TransformHeader, TransformBundle and the dispatch functions.
 *
 * The reference crate dxt-lossless-transform-file-formats-api is Rust-only (no c-exports feature); these
 * ADDITIVE C entry points mirror its public items one to one so that the Rust crate can bind them
 * (paths relative to /root/reference/src/api/dxt-lossless-transform-file-formats-api/src):
 *   TransformFormat                         embed/transform_format.rs:10-32
 *   TransformHeader (u32 LE)                embed/mod.rs:107-160       bits 0-3 format, bits 4-31 format data
 *   EmbeddableBc1Details / Bc2              embed/formats/bc1.rs:33-118, bc2.rs:30-118
 *                                           data bits 0-1 version (0), bit 2 split_colour_endpoints,
 *                                           bits 3-4 YCoCgVariant (stable numbering), rest reserved (0)
 *   TransformBundle                         bundle/mod.rs:37-192, bundle/bc1.rs:45-71, bundle/bc2.rs
 *   dispatch_transform / _untransform       handlers/dispatch.rs:41-143
 *   TransformError / FormatHandlerError / EmbedError   error.rs:19-87, embed/embed_error.rs:7-15
 */
#ifndef DXT_LOSSLESS_TRANSFORM_FILE_FORMATS_H
#define DXT_LOSSLESS_TRANSFORM_FILE_FORMATS_H

#include "dxt_lossless_transform_api_common.h"
#include "dxt_lossless_transform_bc1_api.h"
#include "dxt_lossless_transform_bc2_api.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DLTFF_TRANSFORM_HEADER_SIZE 4 /* embed/mod.rs:87 */

typedef enum DltffTransformFormat {
  DltffTransformFormat_Bc1 = 0,
  DltffTransformFormat_Bc2 = 1,
  DltffTransformFormat_Bc3 = 2,
  DltffTransformFormat_Bc7 = 3,
  DltffTransformFormat_Bc6H = 4,
  DltffTransformFormat_Rgba8888 = 5,
  DltffTransformFormat_Bgra8888 = 6,
  DltffTransformFormat_Bgr888 = 7,
  DltffTransformFormat_Bc4 = 8,
  DltffTransformFormat_Bc5 = 9,
} DltffTransformFormat;

/* The Rust error enums, flattened.  detail_a / detail_b of DltffResult carry the variant's fields. */
typedef enum DltffErrorCode {
  DltffErrorCode_Success = 0,
  DltffErrorCode_Embed_CorruptedEmbeddedData = 1,
  DltffErrorCode_Embed_UnknownFormat = 2,
  DltffErrorCode_FormatHandler_UnknownFileFormat = 3,
  DltffErrorCode_FormatHandler_InvalidInputFileHeader = 4,
  DltffErrorCode_FormatHandler_InvalidRestoredFileHeader = 5,
  DltffErrorCode_FormatHandler_FormatNotImplemented = 6, /* a = DltffTransformFormat */
  DltffErrorCode_FormatHandler_NoBuilderForFormat = 7,   /* a = DltffTransformFormat */
  DltffErrorCode_FormatHandler_OutputBufferTooSmall = 8, /* a = required, b = actual */
  DltffErrorCode_FormatHandler_InputTooShort = 9,        /* a = required, b = actual */
  DltffErrorCode_FormatHandler_InputTooShortForStatedTextureSize = 10, /* a = required, b = actual */
  DltffErrorCode_Bc1 = 11, /* a = Dltbc1ErrorCode (dxt_lossless_transform_bc1_api.h), b = its payload (length) */
  DltffErrorCode_Bc2 = 12, /* a = Dltbc2ErrorCode, b = its payload */
  DltffErrorCode_UnknownTransformFormat = 13,
  DltffErrorCode_InvalidDataAlignment = 14, /* a = size, b = required_divisor */
  DltffErrorCode_NoSupportedHandler = 15,
  DltffErrorCode_NullPointer = 16, /* C ABI only */
} DltffErrorCode;

typedef struct DltffResult {
  DltffErrorCode error_code;
  size_t detail_a;
  size_t detail_b;
} DltffResult;

const char *dltff_error_message(DltffErrorCode code);

/* ---- TransformHeader ------------------------------------------------------------------------------- */
typedef uint32_t DltffTransformHeader;
/* TransformHeader::new (embed/mod.rs:122-128); `data` is truncated to 28 bits. */
DltffTransformHeader dltff_TransformHeader_new(DltffTransformFormat format, uint32_t data);
/* TransformHeader::format (:134-136): false = the 4-bit value is no known TransformFormat;
 * *out_format (optional) receives the raw value either way. */
bool dltff_TransformHeader_format(DltffTransformHeader header, DltffTransformFormat *out_format);
DltffTransformHeader dltff_TransformHeader_format_data(DltffTransformHeader header);
/* read_from_ptr / write_to_ptr (:145-159): unaligned little-endian u32. */
DltffTransformHeader dltff_TransformHeader_read(const uint8_t *ptr);
void dltff_TransformHeader_write(DltffTransformHeader header, uint8_t *ptr);

/* EmbeddableBc{1,2}Details::from_settings(..).to_header() and ::from_header(..).to_settings(). */
DltffTransformHeader dltff_bc1_header_from_settings(YCoCgVariant decorrelation_mode, bool split_colour_endpoints);
DltffTransformHeader dltff_bc2_header_from_settings(YCoCgVariant decorrelation_mode, bool split_colour_endpoints);
DltffResult dltff_bc1_settings_from_header(DltffTransformHeader header, YCoCgVariant *out_decorrelation_mode,
                                           bool *out_split_colour_endpoints);
DltffResult dltff_bc2_settings_from_header(DltffTransformHeader header, YCoCgVariant *out_decorrelation_mode,
                                           bool *out_split_colour_endpoints);

/* ---- TransformBundle ------------------------------------------------------------------------------- */
typedef struct DltffTransformBundle DltffTransformBundle; /* opaque */
DltffTransformBundle *dltff_new_TransformBundle(void);         /* TransformBundle::new: no builders */
DltffTransformBundle *dltff_TransformBundle_default_all(void); /* default manual builders for BC1 and BC2 */
void dltff_free_TransformBundle(DltffTransformBundle *bundle); /* NULL-safe */
/* with_bc1_manual / with_bc1_auto / with_bc2_manual / with_bc2_auto: the builder's state is COPIED; the
 * caller keeps ownership of its builder.  A builder of the other format -> UnknownTransformFormat. */
DltffResult dltff_TransformBundle_with_bc1_manual(DltffTransformBundle *bundle, const Dltbc1ManualTransformBuilder *builder);
DltffResult dltff_TransformBundle_with_bc1_auto(DltffTransformBundle *bundle, const Dltbc1AutoTransformBuilder *builder);
DltffResult dltff_TransformBundle_with_bc2_manual(DltffTransformBundle *bundle, const Dltbc2ManualTransformBuilder *builder);
DltffResult dltff_TransformBundle_with_bc2_auto(DltffTransformBundle *bundle, const Dltbc2AutoTransformBuilder *builder);

/* ---- dispatch --------------------------------------------------------------------------------------- */
/* dispatch_transform: OutputBufferTooSmall, then per format NoBuilderForFormat / the builder's error;
 * formats other than BC1 / BC2 -> UnknownTransformFormat.  *out_header = the header to embed. */
DltffResult dltff_dispatch_transform(DltffTransformFormat format, const uint8_t *input, size_t input_len,
                                     uint8_t *output, size_t output_len, const DltffTransformBundle *bundle,
                                     DltffTransformHeader *out_header);
/* dispatch_untransform: OutputBufferTooSmall, UnknownTransformFormat, CorruptedEmbeddedData,
 * InvalidDataAlignment (in this order), then the untransform. */
DltffResult dltff_dispatch_untransform(DltffTransformHeader header, const uint8_t *input, size_t input_len,
                                       uint8_t *output, size_t output_len);

#ifdef __cplusplus
}
#endif
#endif
