// bcn_kernels.h — host-side launch API of the BCn transform / untransform kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include "bcn_layout.h"

namespace dlt {

// Transform `nblocks` BCn blocks starting at `in` (device pointer) into the per-stream device
// pointers `out` (each already advanced to this range's first block).  Any pointer alignment is
// accepted; 16-byte aligned `in` and naturally aligned streams take the tiled kernels, anything
// else the byte-granular kernel.  Asynchronous on `stream`.
cudaError_t launch_transform(const Settings& st, const uint8_t* in, const StreamPtrs& out, uint64_t nblocks,
                             cudaStream_t stream);

// The same for MANY payloads that share one settings combination, in one launch (grid.y = payload).  `d_items` is a
// DEVICE array; `max_blocks` the largest nblocks among them.  Every item must pass transform_batch_item_ok (16-byte
// aligned blocks, naturally aligned streams — what the tiled kernels need); `ragged` is the OR it accumulates.
struct TransformBatchItem {
    const uint8_t* in;
    StreamPtrs out;
    uint64_t nblocks;
};
bool transform_batch_item_ok(const Settings& st, const TransformBatchItem& item, bool* ragged);
cudaError_t launch_transform_batch(const Settings& st, const TransformBatchItem* d_items, int nitems, uint64_t max_blocks,
                                   bool ragged, cudaStream_t stream);

// The ENDPOINT streams of up to eight transform candidates of one payload from ONE read of its blocks (the best-settings
// search only looks at those: transform_auto.rs estimates out[0, len/2) for BC1, out[len/2, +len/4) for BC2, the alpha-
// endpoint and colour ranges for BC3).  The bytes are exactly what transform_bcN_with_settings writes into those ranges:
//   colour : start of the colour range of the candidate's image — c0c1 words (4 B per block), or with split_colour the c0
//            stream followed by the c1 stream at colour + 2 N
//   alpha  : BC3 only, start of the alpha-endpoint range (nullptr: not wanted) — a0a1 pairs, or with split_alpha the a0
//            stream followed by the a1 stream at alpha + N
// `blocks` must be 16-byte aligned; the streams may start anywhere their element size allows.
struct EndpointCandidate {
    uint8_t* colour;
    uint8_t* alpha;
    int variant;
    bool split_colour, split_alpha;
};
constexpr int kMaxEndpointCandidates = 8;
cudaError_t launch_endpoint_candidates(int format, const uint8_t* blocks, uint64_t nblocks, const EndpointCandidate* cands, int count,
                                       cudaStream_t stream);

// ... and the inverse: item.in = the payload's streams, item.out = its blocks (16-byte aligned).
struct UntransformBatchItem {
    StreamPtrs in;
    uint8_t* out;
    uint64_t nblocks;
};
bool untransform_batch_item_ok(const Settings& st, const UntransformBatchItem& item);
cudaError_t launch_untransform_batch(const Settings& st, const UntransformBatchItem* d_items, int nitems, uint64_t max_blocks,
                                     cudaStream_t stream);

// Many small copies between device-visible buffers (mapped host memory included) in one launch; 4-byte aligned
// pointers (8-byte vectors when both are 8-byte aligned), sizes multiples of 8.  `d_items` is a DEVICE array, `max_bytes` the largest size among them.
struct CopyBatchItem {
    const uint8_t* src;
    uint8_t* dst;
    uint64_t bytes;
};
cudaError_t launch_copy_batch(const CopyBatchItem* d_items, int nitems, uint64_t max_bytes, cudaStream_t stream);

// Exact inverse: gathers the streams back into `nblocks` blocks at `out`.
cudaError_t launch_untransform(const Settings& st, const StreamPtrs& in, uint8_t* out, uint64_t nblocks,
                               cudaStream_t stream);

// split_color_endpoints: [c0 c1] x n (len_bytes = 4n) -> c0 x n | c1 x n at out and out + len_bytes/2.
cudaError_t launch_split_color_endpoints(const uint8_t* in, uint8_t* out, uint64_t len_bytes, cudaStream_t stream);

// experimental::normalize_blocks (BC1; core/dxt-lossless-transform-bc1/src/experimental/normalize_blocks/normalize.rs).
// One pass over `nblocks` blocks writes the normalized image of every mode whose output pointer is non-null
// (out_none receives a copy; an output may equal `in`); *d_any (optional, device) is OR-ed with 1 when at least one
// block is transparent or a round-trippable solid colour (normalize_blocks_all_modes' return value).
// transform_bc1_with_normalize_blocks needs no pass of its own: set Settings::normalize and call launch_transform.
cudaError_t launch_normalize_blocks(const uint8_t* in, uint8_t* out_none, uint8_t* out_color0, uint8_t* out_replicate,
                                    uint64_t nblocks, unsigned int* d_any, cudaStream_t stream);
// normalize_split_blocks_in_place: colours ([c0 c1] per block) and indices in separate device arrays.
cudaError_t launch_normalize_split_blocks(uint8_t* colors, uint8_t* indices, uint64_t nblocks, int mode, cudaStream_t stream);

// Number of kernel launches the two functions above have issued in this process (bench evidence).
uint64_t kernel_launch_count();

// Blocks per CTA tile of the tiled kernels (shard boundaries that are multiples of this keep every
// per-stream slice 16-byte aligned when the stream base is).
constexpr int kTileBytes = 16384;
inline constexpr int tile_blocks(int fmt) { return kTileBytes / block_bytes(fmt); }

}  // namespace dlt
