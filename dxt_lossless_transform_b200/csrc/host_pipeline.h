// host_pipeline.h — the host-pointer path behind the reference-compatible entry points.
//
// The reference's functions are synchronous calls on caller-owned HOST buffers
// (core/dxt-lossless-transform-bc1/src/transform/transform_with_settings.rs:31,92).  Here the same
// call streams the payload through one GPU: the block range is cut into chunks, and for each chunk
// H2D copy -> kernel -> per-stream D2H copies are queued on one of a few CUDA streams so that the
// copy engines and the SMs overlap.  Caller memory that is already page-locked is copied directly;
// pageable memory goes through a small ring of pinned staging slots.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "bcn_layout.h"

namespace dlt {

enum class Status : int { kOk = 0, kCudaError = 1, kOutOfMemory = 2 };

constexpr int kStages = 4;                      // chunk slots (chunks in flight)
constexpr size_t kChunkBytes = 64u << 20;       // largest chunk of the copy pipeline (multiple of the 16 KiB tile)
constexpr size_t kStagedChunkBytes = 16u << 20; // chunk when caller memory is pageable (size of a pinned staging slot)

// Host-path tuning knobs, read once from the environment (diagnostics / benchmarking only):
//   DLTCUDA_CHUNK_MIB  chunk size of the copy pipeline for page-locked buffers, 1..64 MiB (default 64: every copy of a chunk costs ~10 us of copy-engine turnaround, measured 42.2 / 44.7 / 45.0 GB/s per direction with 16 / 32 / 64 MiB chunks)
//   DLTCUDA_STAGES     chunks in flight, 1..4 (default 4)
//   DLTCUDA_ZEROCOPY   1 (default): SMALL page-locked caller buffers are read and written by the
//                      kernel directly over the host link (one launch, lowest latency); 0: always use
//                      the H2D -> kernel -> D2H copy pipeline
//   DLTCUDA_ZEROCOPY_MAX_KIB  largest payload that takes the zero-copy path (default 4096)
//   DLTCUDA_RAMP       1 (default): a large payload starts and ends with short chunks (1/8, 1/4, 1/2) so the
//                      un-overlapped first upload / last download are short; 0: equal chunks
//   DLTCUDA_STRIDED    1 (default): neighbouring streams of equal width travel as one strided (2D) copy; 0: one copy each
//   DLTCUDA_COPY_THREADS      threads used for staging copies of pageable caller memory (per pool; default by core count)
//   DLTCUDA_SPLIT / DLTCUDA_MIN_CHUNK_KIB   a call with ONE payload is cut into about SPLIT chunks (default 6) of at least
//                      MIN_CHUNK_KIB (default 4096) so that its own upload and download overlap
//   DLTCUDA_ZEROCOPY_PIECES   a single mapped-memory payload of >= 1 MiB is launched as up to this many block ranges on
//                      different streams (default 8; 1: one launch), and takes that path up to 8 MiB instead of 4
struct HostPathConfig {
    size_t chunk_bytes;
    int stages;
    bool zero_copy;
    size_t zero_copy_max_bytes;
    bool ramp;
    bool strided;
    size_t split;             // a single-payload call is cut into about this many chunks ...
    size_t min_chunk_bytes;   // ... but not into chunks smaller than this
    size_t zero_copy_pieces;  // a single mapped-memory payload is launched as up to this many block ranges on different streams
};
const HostPathConfig& host_path_config();

struct Context {
    int device = -1;
    cudaStream_t stream[kStages] = {};
    cudaEvent_t done[kStages] = {};
    uint8_t* d_in = nullptr;   // blocks (transform input / untransform output)
    uint8_t* d_out = nullptr;  // streams in the reference layout
    size_t d_cap = 0;
    uint8_t* h_in[kStages] = {};   // pinned staging, kChunkBytes each, allocated on first pageable use
    uint8_t* h_out[kStages] = {};
    uint8_t* d_scratch = nullptr;  // estimator scratch (see estimator.h)
    size_t d_scratch_cap = 0;
    uint8_t* h_scratch = nullptr;  // pinned host scratch kept between calls (zstd search: candidate images)
    size_t h_scratch_cap = 0;
    uint8_t* d_desc = nullptr;     // small device buffer for launch descriptors (batched search: gather lists)
    size_t d_desc_cap = 0;
    int prev_device = -1;          // the calling thread's CUDA device before acquire_context() switched it
    Context* next_free = nullptr;
};

// Borrow a context for `device` (-1 = the calling thread's current CUDA device).  Contexts are
// pooled per device, so concurrent host threads each get their own streams and buffers.
Context* acquire_context(int device, Status* st);
void release_context(Context* ctx);   // also restores the calling thread's previous CUDA device
// The same after waiting for everything queued on the context's streams: what every entry point that may leave
// through an error path uses, so that a pooled context never goes back with work still in flight (another thread
// could be handed its buffers, and page-locked caller memory could still be read after the call has returned).
void release_context_synced(Context* ctx);
// Frees the device / pinned buffers of every idle pooled context (they are retained between calls and sized to the
// largest payload seen); streams and events stay.  Returns the number of bytes released.
size_t release_cached_memory();

Status ensure_device_buffers(Context* ctx, size_t len);
Status ensure_scratch(Context* ctx, size_t bytes);
Status ensure_staging(Context* ctx);
Status ensure_host_scratch(Context* ctx, size_t bytes);
Status ensure_desc(Context* ctx, size_t bytes);

// transform (inverse=false) or untransform (inverse=true) `len` bytes of host memory; blocks until
// `out` holds the result.
Status run_host(const Settings& st, bool inverse, const uint8_t* in, uint8_t* out, size_t len, int device);

// A batch of independent payloads (a directory of textures) through ONE pipeline on one device: the
// chunks of consecutive payloads overlap, there is a single wait at the end.
struct HostJob {
    Settings st;
    bool inverse;
    const uint8_t* in;
    uint8_t* out;
    size_t len;
};
Status run_host_batch(const HostJob* jobs, size_t count, int device);

// The last CUDA error string seen by this thread (diagnostics only).
const char* last_error_string();
void note_cuda_error(cudaError_t e);

// Thread-local device override used by the C ABI (dltcuda_set_device).
void set_thread_device(int device);
int thread_device();

bool is_pinned_host(const void* p);

// Page-locked ranges the current call has seen: one driver query tells whether a pointer is page-locked host memory,
// where its allocation starts and ends and what its device alias is, so the thousands of payloads of a batch that
// were carved out of a few pinned pools cost one query per pool instead of four per payload.  An instance lives for
// ONE call (the caller cannot free memory it has handed to a running call), so it can never go stale.
class PinnedRanges {
public:
    // True when [p, p + len) is page-locked host memory; *dev (optional) receives the device alias of p.
    bool covers(const uint8_t* p, size_t len, uint8_t** dev);

private:
    struct Range {
        const uint8_t* host;
        size_t size;
        uint8_t* dev;
    };
    static bool query(const uint8_t* p, Range* out);
    static constexpr int kMax = 8;
    Range r_[kMax];
    int n_ = 0;
    unsigned next_ = 0;
};

}  // namespace dlt
