/*
 * dxt_lossless_transform_dds.h — the DDS container step around the block transform.
 *
 * Replaces (paths relative to /root/reference/src/extensions/file-formats/dxt-lossless-transform-dds/src):
 *   is_dds, parse_dds, DdsInfo, DdsFormat     dds/exports.rs:12-64, dds/parse_dds.rs:7-42   (the reference's OWN C exports,
 *                                             same names, layouts and values)
 *   dltdds_*                                  ADDITIVE C entry points for what is Rust-only in the reference:
 *     parse_dds_ignore_magic                  dds/parse_dds.rs:78-172
 *     DdsHandler::transform_bundle            handler/file_format_handler.rs:17-86
 *     DdsHandler::untransform                 handler/file_format_handler.rs:88-145
 *     DdsHandler::can_handle                  handler/file_format_detection.rs:7-17
 *     DdsHandler::can_handle_untransform      handler/file_format_untransform_detection.rs:7-22
 *   dltdds_*_batch                            no reference counterpart: the reference's CLI handles a directory one file per
 *                                             rayon task (tools/dxt-lossless-transform-cli/src/commands/transform/mod.rs:154-176);
 *                                             here all payloads of a batch share one pinned copy pipeline per GPU.
 *
 * Files written by dltdds_transform_bundle carry the reference's 4-byte TransformHeader in place of the
 * "DDS " magic, so the stock CPU reference can untransform them and vice versa.  BC3 and later formats are
 * FormatNotImplemented, exactly as in the reference (handler/format_conversion.rs:45-107).
 */
#ifndef DXT_LOSSLESS_TRANSFORM_DDS_H
#define DXT_LOSSLESS_TRANSFORM_DDS_H

#include "dxt_lossless_transform_file_formats.h"

#ifdef __cplusplus
extern "C" {
#endif

/* dds/parse_dds.rs:7-33, repr(u8). */
enum {
  DdsFormat_NotADds = 0,
  DdsFormat_Unknown = 1,
  DdsFormat_BC1 = 2,
  DdsFormat_BC2 = 3,
  DdsFormat_BC3 = 4,
  DdsFormat_BC6H = 5,
  DdsFormat_BC7 = 6,
  DdsFormat_RGBA8888 = 7,
  DdsFormat_BGRA8888 = 8,
  DdsFormat_BGR888 = 9,
  DdsFormat_BC4 = 10,
  DdsFormat_BC5 = 11,
};
typedef uint8_t DdsFormat;

/* dds/parse_dds.rs:36-42, repr(C): 8 bytes. */
typedef struct DdsInfo {
  DdsFormat format;
  uint8_t data_offset;  /* 128, or 148 with a DX10 header */
  uint32_t data_length; /* bytes of texture data incl. the mip chain; 0 if it cannot be derived */
} DdsInfo;

/* dds/exports.rs:12-21: 'DDS ' magic and at least 128 bytes.  NULL / 0 -> false. */
bool is_dds(const uint8_t *ptr, size_t len);
/* dds/exports.rs:41-64: format = DdsFormat_NotADds when `ptr` does not hold a DDS. */
DdsInfo parse_dds(const uint8_t *ptr, size_t len);

/* parse_dds_ignore_magic: for transformed files, whose first 4 bytes hold the TransformHeader. */
DdsInfo dltdds_parse_dds_ignore_magic(const uint8_t *ptr, size_t len);

/* file_extension: NULL = unknown (accepted), otherwise must equal "dds". */
bool dltdds_can_handle(const uint8_t *input, size_t len, const char *file_extension);
bool dltdds_can_handle_untransform(const uint8_t *input, size_t len, const char *file_extension);

/* DdsHandler::transform_bundle: header copied, texture data transformed with the bundle's builder for the
 * detected format (on the GPU), trailing bytes copied, magic replaced by the TransformHeader.
 * Check order: OutputBufferTooSmall, InvalidInputFileHeader, InputTooShortForStatedTextureSize,
 * FormatNotImplemented / UnknownFileFormat, then the bundle's errors. */
DltffResult dltdds_transform_bundle(const uint8_t *input, size_t input_len, uint8_t *output,
                                    size_t output_len, const DltffTransformBundle *bundle);
/* DdsHandler::untransform.  Check order: InputTooShort, OutputBufferTooSmall, InvalidRestoredFileHeader,
 * InputTooShortForStatedTextureSize, then dispatch_untransform's errors. */
DltffResult dltdds_untransform(const uint8_t *input, size_t input_len, uint8_t *output,
                               size_t output_len);

/* A directory of files at once.  results[i] is what the single-file call would have returned for files[i];
 * the return value is non-zero only for NULL arguments (1) or when the host ran out of memory (2).  devices == NULL: the calling thread's device
 * (dltcuda_set_device); otherwise the payloads are dealt out over `num_devices` GPUs of this box. */
typedef struct DltddsFile {
  const uint8_t *input;
  size_t input_len;
  uint8_t *output;
  size_t output_len;
} DltddsFile;
int dltdds_transform_bundle_batch(const DltddsFile *files, size_t count,
                                  const DltffTransformBundle *bundle, DltffResult *results,
                                  const int *devices, int num_devices);
int dltdds_untransform_batch(const DltddsFile *files, size_t count, DltffResult *results,
                             const int *devices, int num_devices);

#ifdef __cplusplus
}
#endif
#endif
