/*
 * ltu_params.h — the tunables of the restated LTU match estimator (TEST INFRASTRUCTURE).
 *
 * The reference calls exactly one function of the third-party crate
 * `lossless-transform-utils` 0.1.3 (crates.io; src/Cargo.lock:909-915, checksum
 * 5ac26ac93dad27151c040b8cf326018c25b6900e6e925eabbf53242d0ec938ba):
 *   lossless_transform_utils::match_estimator::estimate_num_lz_matches_fast(data)
 * (call site: extensions/estimators/dxt-lossless-transform-ltu/src/lib.rs:73).
 * Its source is NOT under /root/reference and there is no network, so the
 * algorithm below is a restatement of the crate's published algorithm:
 *
 *   table: [u32; 1 << HASH_BITS], zero-initialised, one per call
 *   for (i = 0; i < len.saturating_sub(7); i += 4):
 *       for k in 0..4:  d_k = read_u32_le(data + i + k) & 0x00FF_FFFF       (3-byte window)
 *                       idx_k = (d_k * GOLDEN_RATIO) >> (32 - HASH_BITS)
 *       matches += (table[idx_0] == d_0) + ... + (table[idx_3] == d_3)      (4 compares first)
 *       table[idx_0] = d_0; ...; table[idx_3] = d_3                         (then 4 updates, in order)
 *
 * **PARITY UNPINNED**: HASH_BITS, the index derivation, the 4-compare-then-4-update grouping and
 * the loop bound are the parts that could not be checked against the 0.1.3 source here.  The
 * product's CUDA estimator (csrc/estimator.cu) is built from the same five constants (it has its
 * own copy in include/dxt_lossless_transform_cuda.h), so fixing a constant is a two-line change.
 */
#ifndef LTU_PARAMS_H
#define LTU_PARAMS_H

#define LTU_HASH_BITS 16
#define LTU_GOLDEN_RATIO 0x9E3779B1u
#define LTU_KEY_MASK 0x00FFFFFFu
#define LTU_GROUP 4      /* positions per loop iteration (compare all, then update all) */
#define LTU_TAIL_GUARD 7 /* loop runs while i < len.saturating_sub(7) */

#endif
