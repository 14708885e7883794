// zstd_estimator.h — see zstd_estimator.cu.
#pragma once
#include <cstddef>
#include <cstdint>

#include "cabi_internal.h"

namespace dlt {

// True when `e` was made by dltzstd_new_size_estimator (recognised by its callback, like the LTU estimator).
bool is_zstd_estimator(const DltSizeEstimator& e);
int zstd_estimator_level(const DltSizeEstimator& e);
// One ZSTD_compress2 with the reference's parameters; 0 = success (callback error codes otherwise).  Thread-safe.
uint32_t zstd_compressed_size(int level, const uint8_t* input, size_t len, uint8_t* output, size_t output_len,
                              size_t* out_size);

}  // namespace dlt
