/*
 * dxt_lossless_transform_ltu.h — LTU size-estimator factory (drop-in for the cbindgen header of
 * crate dxt-lossless-transform-ltu; /root/reference/src/extensions/estimators/dxt-lossless-transform-ltu).
 *
 * The returned DltSizeEstimator has the reference's semantics
 *   estimate = len.saturating_sub(estimate_num_lz_matches_fast(data)); null or empty -> 0;
 *   max_compressed_size -> 0                                                  (src/lib.rs:67-119)
 * and callback error codes 1 (null context/out), 2, 3 (src/c_api.rs:106-150).  The match count is
 * computed by the CUDA estimator; estimate_num_lz_matches_fast itself lives in the third-party
 * crate lossless-transform-utils 0.1.3 whose source is not in the reference tree — its algorithm
 * is restated and its exact values are PARITY-UNPINNED.
 */
#ifndef DXT_LOSSLESS_TRANSFORM_LTU_H
#define DXT_LOSSLESS_TRANSFORM_LTU_H

#include "dxt_lossless_transform_api_common.h"

#ifdef __cplusplus
extern "C" {
#endif

/* src/c_api.rs:74 — NULL on allocation failure; free with dltltu_free_size_estimator. */
DltSizeEstimator *dltltu_new_size_estimator(void);
/* src/c_api.rs:89 — null-safe. */
void dltltu_free_size_estimator(DltSizeEstimator *estimator);

#ifdef __cplusplus
}
#endif
#endif
