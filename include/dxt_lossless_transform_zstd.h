/*
 * dxt_lossless_transform_zstd.h — ZStandard size-estimator factory.
 *
 * Mirrors ZStandardSizeEstimation of crate dxt-lossless-transform-zstd
 * (/root/reference/src/extensions/compressors/dxt-lossless-transform-zstd/src/lib.rs:54-140): the estimate of a byte
 * range is the size ZSTD_compress2 produces for it with
 *   compressionLevel = level, format = ZSTD_f_zstd1_magicless, contentSizeFlag = checksumFlag = dictIDFlag = 0
 * (lib.rs:193-209); null or empty input -> 0; max_compressed_size(len) = ZSTD_compressBound(len), 0 for len == 0.
 * The reference crate has NO C exports — these three functions are additive, shaped like the dltltu_* pair
 * (extensions/estimators/dxt-lossless-transform-ltu/src/c_api.rs:74-150) so the estimator plugs into
 * dltbc{1,2}_new_AutoTransformBuilder / dltbc{1,2,3}core_transform_auto unchanged.
 *
 * zstd itself is a third-party library (the reference links zstd-sys 2.0.16+zstd.1.5.7); here it is bound at run time
 * from the system's libzstd.so.1 (override: environment variable DLTCUDA_LIBZSTD).  Sizes equal the reference's when the
 * zstd versions match.  With this estimator transform_bcN_auto transforms the candidates on the GPU and compresses all
 * of them concurrently on host threads.
 */
#ifndef DXT_LOSSLESS_TRANSFORM_ZSTD_H
#define DXT_LOSSLESS_TRANSFORM_ZSTD_H

#include "dxt_lossless_transform_api_common.h"

#ifdef __cplusplus
extern "C" {
#endif

/* lib.rs:60-69 — NULL when compression_level is outside 1..=22 (InvalidLevel), when no libzstd can be loaded, or on
 * allocation failure.  Levels of the reference's named constructors: new_fast = 1, new_default = 3, new_best = 22
 * (lib.rs:72-92).  Free with dltzstd_free_size_estimator. */
DltSizeEstimator *dltzstd_new_size_estimator(int compression_level);
/* null-safe */
void dltzstd_free_size_estimator(DltSizeEstimator *estimator);
/* ZSTD_versionNumber() of the library in use (10507 = 1.5.7); 0 when none could be loaded. */
unsigned dltzstd_version_number(void);

#ifdef __cplusplus
}
#endif
#endif
