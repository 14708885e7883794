// host_pipeline.cu — see host_pipeline.h.
#include "host_pipeline.h"

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <algorithm>
#include <utility>
#include <vector>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif

#include <atomic>

#include <cuda.h>   // types of cuPointerGetAttributes only; the entry point comes from cudaGetDriverEntryPoint

#include "bcn_kernels.h"
#include "copy_pool.h"

static_assert(dlt::kCopyPoolMaxSegs >= dlt::kMaxStreams, "a copy job takes all streams of a chunk");

namespace dlt {
namespace {

constexpr int kMaxDevices = 64;
constexpr size_t kSlotPad = 256 * kMaxStreams;
constexpr size_t kStagingSlotBytes = kStagedChunkBytes + kSlotPad;

std::mutex g_pool_mutex;
Context* g_free[kMaxDevices] = {};

thread_local int t_device = -1;
thread_local char t_error[256] = "";

#define DLT_CUDA(expr)                                         \
    do {                                                       \
        cudaError_t e__ = (expr);                              \
        if (e__ != cudaSuccess) {                              \
            note_cuda_error(e__);                              \
            return e__ == cudaErrorMemoryAllocation ? Status::kOutOfMemory : Status::kCudaError; \
        }                                                      \
    } while (0)

inline void staged_copy(uint8_t* dst, const uint8_t* src, size_t n) { CopyPool::fill().copy(dst, src, n); }


Status create_context(int device, Context** out) {
    Context* c = new Context();
    c->device = device;
    for (int i = 0; i < kStages; i++) {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming);
        if (e != cudaSuccess) {
            note_cuda_error(e);
            delete c;
            return Status::kCudaError;
        }
    }
    *out = c;
    return Status::kOk;
}

}  // namespace

const HostPathConfig& host_path_config() {
    static const HostPathConfig cfg = [] {
        HostPathConfig c{kChunkBytes, kStages, true, (size_t)4 << 20, true, true, 6, (size_t)4 << 20, 8};   // split / min chunk: tools/latency_probe.py
        if (const char* v = std::getenv("DLTCUDA_ZEROCOPY_PIECES")) {
            const long n = std::atol(v);
            if (n >= 1 && n <= 64) c.zero_copy_pieces = (size_t)n;
        }
        if (const char* v = std::getenv("DLTCUDA_SPLIT")) {
            const long n = std::atol(v);
            if (n >= 1 && n <= 64) c.split = (size_t)n;
        }
        if (const char* v = std::getenv("DLTCUDA_MIN_CHUNK_KIB")) {
            const long kib = std::atol(v);
            if (kib >= 16 && (size_t)kib << 10 <= kChunkBytes) c.min_chunk_bytes = ((size_t)kib << 10) / kTileBytes * kTileBytes;
        }
        if (const char* v = std::getenv("DLTCUDA_CHUNK_MIB")) {
            const long mib = std::atol(v);
            if (mib >= 1 && (size_t)mib << 20 <= kChunkBytes) c.chunk_bytes = (size_t)mib << 20;
        }
        if (const char* v = std::getenv("DLTCUDA_STAGES")) {
            const long n = std::atol(v);
            if (n >= 1 && n <= kStages) c.stages = (int)n;
        }
        if (const char* v = std::getenv("DLTCUDA_ZEROCOPY")) c.zero_copy = std::atol(v) != 0;
        if (const char* v = std::getenv("DLTCUDA_RAMP")) c.ramp = std::atol(v) != 0;
        if (const char* v = std::getenv("DLTCUDA_STRIDED")) c.strided = std::atol(v) != 0;
        if (const char* v = std::getenv("DLTCUDA_ZEROCOPY_MAX_KIB")) {
            const long kib = std::atol(v);
            if (kib >= 0) c.zero_copy_max_bytes = (size_t)kib << 10;
        }
        return c;
    }();
    return cfg;
}

void note_cuda_error(cudaError_t e) {
    std::strncpy(t_error, cudaGetErrorString(e), sizeof(t_error) - 1);
    t_error[sizeof(t_error) - 1] = 0;
    (void)cudaGetLastError();  // clear the sticky-less error state
}
const char* last_error_string() { return t_error; }
void set_thread_device(int device) { t_device = device; }
int thread_device() { return t_device; }

Context* acquire_context(int device, Status* st) {
    *st = Status::kOk;
    if (device < 0) device = t_device;
    if (device < 0) {
        cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess) {
            note_cuda_error(e);
            *st = Status::kCudaError;
            return nullptr;
        }
    }
    if (device >= kMaxDevices) {
        *st = Status::kCudaError;
        return nullptr;
    }
    // the call runs on `device`; the thread's own current device (torch, another library) is put back on release
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        note_cuda_error(e);
        *st = Status::kCudaError;
        return nullptr;
    }
    Context* c = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        if ((c = g_free[device]) != nullptr) {
            g_free[device] = c->next_free;
            c->next_free = nullptr;
        }
    }
    if (!c) *st = create_context(device, &c);
    if (c) c->prev_device = prev;
    else if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return c;
}

void release_context(Context* ctx) {
    if (!ctx) return;
    if (ctx->prev_device >= 0 && ctx->prev_device != ctx->device) cudaSetDevice(ctx->prev_device);
    ctx->prev_device = -1;
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    ctx->next_free = g_free[ctx->device];
    g_free[ctx->device] = ctx;
}

void release_context_synced(Context* ctx) {
    if (!ctx) return;
    for (int i = 0; i < kStages; i++)
        if (ctx->stream[i]) cudaStreamSynchronize(ctx->stream[i]);   // errors were reported where they happened
    release_context(ctx);
}

size_t release_cached_memory() {
    size_t freed = 0;
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    for (int d = 0; d < kMaxDevices; d++) {
        if (!g_free[d] || cudaSetDevice(d) != cudaSuccess) continue;
        for (Context* c = g_free[d]; c; c = c->next_free) {
            if (c->d_in) cudaFree(c->d_in), freed += c->d_cap;
            if (c->d_out) cudaFree(c->d_out), freed += c->d_cap;
            c->d_in = c->d_out = nullptr, c->d_cap = 0;
            if (c->d_scratch) cudaFree(c->d_scratch), freed += c->d_scratch_cap;
            c->d_scratch = nullptr, c->d_scratch_cap = 0;
            if (c->h_scratch) cudaFreeHost(c->h_scratch), freed += c->h_scratch_cap;
            c->h_scratch = nullptr, c->h_scratch_cap = 0;
            if (c->d_desc) cudaFree(c->d_desc), freed += c->d_desc_cap;
            c->d_desc = nullptr, c->d_desc_cap = 0;
            for (int i = 0; i < kStages; i++) {
                if (c->h_in[i]) cudaFreeHost(c->h_in[i]), freed += kStagingSlotBytes;
                if (c->h_out[i]) cudaFreeHost(c->h_out[i]), freed += kStagingSlotBytes;
                c->h_in[i] = c->h_out[i] = nullptr;
            }
        }
    }
    if (prev >= 0) cudaSetDevice(prev);
    return freed;
}

Status ensure_device_buffers(Context* ctx, size_t len) {
    if (ctx->d_cap >= len && ctx->d_in) return Status::kOk;
    if (ctx->d_in) cudaFree(ctx->d_in);
    if (ctx->d_out) cudaFree(ctx->d_out);
    ctx->d_in = ctx->d_out = nullptr;
    ctx->d_cap = 0;
    const size_t cap = (len + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    DLT_CUDA(cudaMalloc(&ctx->d_in, cap));
    DLT_CUDA(cudaMalloc(&ctx->d_out, cap));
    ctx->d_cap = cap;
    return Status::kOk;
}

Status ensure_scratch(Context* ctx, size_t bytes) {
    if (ctx->d_scratch_cap >= bytes && ctx->d_scratch) return Status::kOk;
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    ctx->d_scratch = nullptr;
    ctx->d_scratch_cap = 0;
    const size_t cap = (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    DLT_CUDA(cudaMalloc(&ctx->d_scratch, cap));
    ctx->d_scratch_cap = cap;
    return Status::kOk;
}

Status ensure_host_scratch(Context* ctx, size_t bytes) {
    if (ctx->h_scratch_cap >= bytes && ctx->h_scratch) return Status::kOk;
    if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
    ctx->h_scratch = nullptr;
    ctx->h_scratch_cap = 0;
    const size_t cap = (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    DLT_CUDA(cudaMallocHost(&ctx->h_scratch, cap));
    ctx->h_scratch_cap = cap;
    return Status::kOk;
}

Status ensure_desc(Context* ctx, size_t bytes) {
    if (ctx->d_desc_cap >= bytes && ctx->d_desc) return Status::kOk;
    if (ctx->d_desc) cudaFree(ctx->d_desc);
    ctx->d_desc = nullptr;
    ctx->d_desc_cap = 0;
    const size_t cap = (bytes + 65535) & ~(size_t)65535;
    DLT_CUDA(cudaMalloc(&ctx->d_desc, cap));
    ctx->d_desc_cap = cap;
    return Status::kOk;
}

Status ensure_staging(Context* ctx) {
    for (int i = 0; i < kStages; i++) {
        if (!ctx->h_in[i]) DLT_CUDA(cudaMallocHost(&ctx->h_in[i], kStagingSlotBytes));
        if (!ctx->h_out[i]) DLT_CUDA(cudaMallocHost(&ctx->h_out[i], kStagingSlotBytes));
    }
    return Status::kOk;
}

bool is_pinned_host(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

namespace {

}  // namespace

bool PinnedRanges::covers(const uint8_t* p, size_t len, uint8_t** dev) {
    if (len == 0) return false;
    for (int i = 0; i < n_; i++) {
        const Range& r = r_[i];
        if (p >= r.host && p + len <= r.host + r.size) {
            if (dev) *dev = r.dev + (p - r.host);
            return true;
        }
    }
    Range r{};
    if (!query(p, &r)) return false;
    if (r.size > 1) {
        if (n_ < kMax) r_[n_++] = r;
        else r_[next_++ % kMax] = r;
    }
    if (p + len <= r.host + r.size) {
        if (dev) *dev = r.dev + (p - r.host);
        return true;
    }
    // no range information (fallback query): accept when the last byte is page-locked too
    if (r.size == 1 && is_pinned_host(p + len - 1)) {
        if (dev) *dev = r.dev;
        return true;
    }
    return false;
}

bool PinnedRanges::query(const uint8_t* p, Range* out) {
    using Fn = CUresult (*)(unsigned int, CUpointer_attribute*, void**, CUdeviceptr);
    static Fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q{};
        if (cudaGetDriverEntryPoint("cuPointerGetAttributes", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<Fn>(f);
    }();
    if (fn) {
        unsigned int type = 0;
        CUdeviceptr dptr = 0, start = 0;
        void* hptr = nullptr;
        size_t size = 0;
        CUpointer_attribute attrs[5] = {CU_POINTER_ATTRIBUTE_MEMORY_TYPE, CU_POINTER_ATTRIBUTE_DEVICE_POINTER,
                                        CU_POINTER_ATTRIBUTE_HOST_POINTER, CU_POINTER_ATTRIBUTE_RANGE_START_ADDR,
                                        CU_POINTER_ATTRIBUTE_RANGE_SIZE};
        void* data[5] = {&type, &dptr, &hptr, &start, &size};
        if (fn(5, attrs, data, reinterpret_cast<CUdeviceptr>(p)) == CUDA_SUCCESS && type == CU_MEMORYTYPE_HOST && dptr && size &&
            hptr == p) {
            // the range start is reported in the address space of the queried pointer (host) or of its device alias;
            // with unified addressing the two coincide.  Host and device aliases share offsets.
            const CUdeviceptr hp = reinterpret_cast<CUdeviceptr>(p);
            size_t before = SIZE_MAX;
            if (start <= hp && hp - start < size) before = (size_t)(hp - start);
            else if (start <= dptr && dptr - start < size) before = (size_t)(dptr - start);
            if (before != SIZE_MAX) {
                out->host = p - before;
                out->size = size;
                out->dev = reinterpret_cast<uint8_t*>(dptr) - before;
                return true;
            }
        }
        if (type != CU_MEMORYTYPE_HOST) return false;
    }
    // fallback: the runtime's per-pointer query, a range of one byte (the caller then checks the last byte too)
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
    out->host = p, out->size = 1, out->dev = static_cast<uint8_t*>(a.devicePointer);
    return true;
}

namespace {

// Device-side chunk slots: blocks in one buffer, the chunk's streams compacted in another with
// every stream 256-byte aligned (chunk block counts are powers of two), so the tiled kernels always
// run their aligned path no matter what N is.
struct Slots {
    uint8_t* blocks[kStages];
    uint8_t* streams[kStages];
};

Status ensure_slots(Context* ctx, Slots* s, size_t chunk_bytes) {
    // d_in / d_out double as the slot arena: kStages slots each.
    const size_t slot = chunk_bytes + kSlotPad;
    Status st = ensure_device_buffers(ctx, slot * kStages);
    if (st != Status::kOk) return st;
    for (int i = 0; i < kStages; i++) {
        s->blocks[i] = ctx->d_in + slot * i;
        s->streams[i] = ctx->d_out + slot * i;
    }
    return Status::kOk;
}

}  // namespace

namespace {

// One pass of host payloads through a context: every payload is cut into chunks, chunk i uses slot
// (= CUDA stream) i % stages, so slot reuse is ordered on the device; the host only waits where it
// has to touch a pinned staging slot again.
class HostPipeline {
public:
    explicit HostPipeline(Context* ctx) : ctx_(ctx), cfg_(host_path_config()) {}
    ~HostPipeline() { stop_drainer(); }

    // What prepare() learns about a job: where its buffers live and the largest chunk it will use.
    struct JobInfo {
        bool in_pinned = false, out_pinned = false, zero_copy = false;
        // zero_copy, but a side the tiled kernels cannot address as it lies in host memory (blocks not 16-byte aligned,
        // a stream without its natural alignment): that side passes through a piece of a device arena
        bool in_arena = false, out_arena = false;
        uint8_t *in_dev = nullptr, *out_dev = nullptr;   // device aliases of page-locked buffers
        size_t chunk_bytes = 0, in_off = 0, out_off = 0;
    };

    Status prepare(const HostJob* jobs, size_t count) {
        size_t max_chunk = 0;
        bool staged = false;
        info_.assign(count, JobInfo{});
        constexpr size_t kArenaBudget = (size_t)1 << 30;
        for (size_t i = 0; i < count; i++) {
            const HostJob& job = jobs[i];
            if (job.len == 0) continue;
            JobInfo& f = info_[i];
            f.in_pinned = pinned_.covers(job.in, job.len, &f.in_dev);
            f.out_pinned = pinned_.covers(job.out, job.len, &f.out_dev);
            size_t c = f.in_pinned && f.out_pinned ? cfg_.chunk_bytes : std::min(cfg_.chunk_bytes, kStagedChunkBytes);
            // A call with ONE mid-size payload has nothing else to overlap its upload and download with: cut it into
            // ~split pieces (not below min_chunk: every DMA costs ~10 us of turnaround) so that they overlap each other.
            // One chunk -> six: 16 MiB 642 -> 525 us, 64 MiB 2.44 -> 1.79 ms, 128 MiB 3.83 -> 3.30 ms (tools/latency_probe.py; 4, 5, 6, 8,
            // 12 pieces and 1 - 8 MiB minimum tried); a batch overlaps across payloads anyway.
            if (count == 1 && cfg_.split > 1) {
                const size_t piece = (job.len / cfg_.split + (size_t)kTileBytes - 1) / kTileBytes * kTileBytes;
                c = std::min(c, std::max(piece, cfg_.min_chunk_bytes));
            }
            f.chunk_bytes = std::min(c, (job.len + (size_t)kTileBytes - 1) / kTileBytes * kTileBytes);
            // (a call with one payload launches it as staggered block ranges, which beats the copy pipeline up to 8 MiB:
            // 282 against 301 us; a batch overlaps its payloads in the pipeline instead)
            const size_t zc_max = count == 1 && cfg_.zero_copy_pieces > 1 ? std::max(cfg_.zero_copy_max_bytes, (size_t)8 << 20) : cfg_.zero_copy_max_bytes;
            f.zero_copy = cfg_.zero_copy && job.len <= zc_max && cfg_.zero_copy_max_bytes > 0 && f.in_pinned && f.out_pinned;
            if (f.zero_copy) {
                // Mapped memory directly where the tiled kernels can take the pointers: the streams need their natural
                // alignment, the blocks 16 bytes.  A DDS payload behind a 148-byte DX10 header is 4-byte aligned: such a
                // side passes through a 256-byte aligned piece of a device arena — ONE gather kernel before and ONE scatter
                // kernel after the launches serve all such payloads of the call — while the other side still travels
                // over the mapped memory.  Sides that are not even 4-byte aligned take the copy pipeline.
                const Settings& st = job.st;
                const size_t n = job.len / block_bytes(st.format);
                alignas(16) static uint8_t aligned_dummy[16];
                bool ragged = false;
                auto streams_ok = [&](uint8_t* base) {
                    return transform_batch_item_ok(st, TransformBatchItem{aligned_dummy, reference_layout(base, n, 0, st), n}, &ragged);
                };
                const bool in_ok = job.inverse ? streams_ok(f.in_dev) : (reinterpret_cast<uintptr_t>(f.in_dev) & 15) == 0;
                const bool out_ok = job.inverse ? (reinterpret_cast<uintptr_t>(f.out_dev) & 15) == 0 : streams_ok(f.out_dev);
                const size_t piece = (job.len + 255) / 256 * 256;
                const size_t need = (in_ok ? 0 : piece) + (out_ok ? 0 : piece);
                if ((!in_ok && (reinterpret_cast<uintptr_t>(f.in_dev) & 3)) || (!out_ok && (reinterpret_cast<uintptr_t>(f.out_dev) & 3)) ||
                    arena_bytes_ + need > kArenaBudget) {
                    f.zero_copy = false;
                } else {
                    if (!in_ok) f.in_arena = true, f.in_off = arena_bytes_, arena_bytes_ += piece;
                    if (!out_ok) f.out_arena = true, f.out_off = arena_bytes_, arena_bytes_ += piece;
                }
            }
            if (f.zero_copy) continue;
            max_chunk = std::max(max_chunk, f.chunk_bytes);
            staged |= !f.in_pinned || !f.out_pinned;
        }
        if (max_chunk) {
            const Status st = ensure_slots(ctx_, &slots_, max_chunk);
            if (st != Status::kOk) return st;
        }
        return staged ? ensure_staging(ctx_) : Status::kOk;
    }

    Status submit(const HostJob& job, size_t index) {
        if (job.len == 0) return Status::kOk;
        const JobInfo& info = info_[index];
        const Settings& st = job.st;
        const int bpb = block_bytes(st.format);
        const int ns = num_streams(st.format, st.split_alpha, st.split_colour);
        const size_t n = job.len / bpb;
        const bool in_pinned = info.in_pinned, out_pinned = info.out_pinned;
        const size_t chunk_bytes = info.chunk_bytes;

        // Small page-locked payloads: the kernel reads the blocks and writes the streams straight over the host link
        // (mapped memory) — no staging copies, lowest latency.  They are not launched one by one: all payloads of the
        // call that share a settings combination go out as ONE launch (grid.y = payload) when the call drains.
        if (info.zero_copy) {
            ZeroCopyGroup* g = nullptr;
            for (ZeroCopyGroup& c : groups_)
                if (c.inverse == job.inverse && c.st.format == st.format && c.st.variant == st.variant &&
                    c.st.split_alpha == st.split_alpha && c.st.split_colour == st.split_colour && c.st.normalize == st.normalize &&
                    c.count() < 65535)
                    g = &c;
            if (!g) {
                groups_.emplace_back();
                g = &groups_.back();
                g->st = st, g->inverse = job.inverse;
            }
            // A side that goes through the arena is an OFFSET into it until flush_groups() knows where the arena is (the
            // arena is 256-byte aligned, so alignment-dependent decisions can be taken on the offsets).
            uint8_t* in_ptr = info.in_arena ? reinterpret_cast<uint8_t*>(info.in_off) : info.in_dev;
            uint8_t* out_ptr = info.out_arena ? reinterpret_cast<uint8_t*>(info.out_off) : info.out_dev;
            if (info.in_arena) gather_.push_back(CopyBatchItem{info.in_dev, in_ptr, job.len});
            if (info.out_arena) scatter_.push_back(CopyBatchItem{out_ptr, info.out_dev, job.len});
            const char flags = (char)((info.in_arena ? 1 : 0) | (info.out_arena ? 2 : 0));
            if (!job.inverse) {
                const TransformBatchItem item{in_ptr, reference_layout(out_ptr, n, 0, st), n};
                bool ragged = false;
                (void)transform_batch_item_ok(st, item, &ragged);
                if (st.normalize == kNormNone) {
                    g->fwd.push_back(item), g->fwd_arena.push_back(flags), g->ragged |= ragged;
                    g->max_blocks = std::max<uint64_t>(g->max_blocks, n);
                    return Status::kOk;
                }
                // (normalizing transforms have no batch instantiation: one launch each, at flush time)
                late_.push_back(LateTransform{st, item, flags});
            } else {
                g->inv.push_back(UntransformBatchItem{reference_layout(in_ptr, n, 0, st), out_ptr, n});
                g->inv_arena.push_back(flags);
                g->max_blocks = std::max<uint64_t>(g->max_blocks, n);
            }
            return Status::kOk;
        }

        const size_t chunk_blocks = chunk_bytes / bpb;
        // Chunk schedule.  The first upload and the last download of a call have nothing to overlap with, so a large
        // payload starts and ends with short chunks (1/8, 1/4, 1/2 of a chunk): the exposed copy shrinks from one
        // chunk to an eighth of one at either end (~5 % of a 1 GiB call with 32 MiB chunks).
        std::vector<std::pair<size_t, size_t>> chunks;   // (first block, blocks)
        {
            const size_t tile = (size_t)kTileBytes / bpb;
            std::vector<size_t> ramp;
            if (cfg_.ramp && n >= 8 * chunk_blocks)
                for (size_t c = chunk_blocks / 8; c < chunk_blocks; c *= 2) ramp.push_back(std::max(tile, c / tile * tile));
            size_t ramp_total = 0;
            for (size_t c : ramp) ramp_total += c;
            size_t b = 0;
            for (size_t c : ramp) chunks.emplace_back(b, c), b += c;
            const size_t middle_end = n - ramp_total;
            while (b < middle_end) {
                const size_t c = std::min(chunk_blocks, middle_end - b);
                chunks.emplace_back(b, c), b += c;
            }
            // descending tail; its last (smallest) chunk takes whatever is left
            for (size_t i = ramp.size(); i-- > 0;) {
                const size_t c = i == 0 ? n - b : ramp[i];
                chunks.emplace_back(b, c), b += c;
            }
        }
        int w[kMaxStreams], pre[kMaxStreams];
        for (int k = 0; k < ns; k++) {
            w[k] = stream_width(st.format, st.split_alpha, st.split_colour, k);
            pre[k] = stream_prefix(st.format, st.split_alpha, st.split_colour, k);
        }
        // Host offset of stream k, block b of the payload (reference layout); slot offset of the
        // same stream inside a chunk (streams 256-byte aligned: chunk_blocks is a tile multiple).
        auto host_off = [&](int k, size_t b) { return n * (size_t)pre[k] + (size_t)w[k] * b; };
        auto slot_off = [&](int k) { return chunk_blocks * (size_t)pre[k]; };
        // Consecutive streams of equal width (c0|c1, a0|a1, colours|indices) sit one pitch apart on both sides — N*w in
        // the payload, chunk_blocks*w in the slot — so a run of them is ONE strided copy with a row per stream; every
        // copy costs the copy engine 10-30 us of turnaround, whatever its size.  (2D copies need pitch < 2 GiB.)
        int run_of[kMaxStreams];
        for (int k = 0; k < ns;) {
            int r = 1;
            while (cfg_.strided && k + r < ns && w[k + r] == w[k] && pre[k + r] == pre[k] + r * w[k] &&
                   n * (size_t)w[k] < ((size_t)1 << 31))
                r++;
            run_of[k] = r;
            k += r;
        }

        // a pageable destination of more than a few chunks: downloads are copied out by the drain thread
        if (!out_pinned && chunks.size() > 2) start_drainer();
        for (const auto& chunk : chunks) {
            const int slot = (int)(seq_++ % cfg_.stages);
            Status f = finish(slot);
            if (f != Status::kOk) return f;
            cudaStream_t s = ctx_->stream[slot];
            const size_t b0 = chunk.first;
            const size_t nb = chunk.second;
            StreamPtrs sp{};
            for (int k = 0; k < ns; k++) sp.p[k] = slots_.streams[slot] + slot_off(k);
            Pending& pend = pending_[slot];
            pend = Pending{};
            if (!job.inverse) {
                const uint8_t* src = job.in + b0 * bpb;
                if (!in_pinned) {
                    staged_copy(ctx_->h_in[slot], src, nb * bpb);
                    src = ctx_->h_in[slot];
                }
                DLT_CUDA(cudaMemcpyAsync(slots_.blocks[slot], src, nb * bpb, cudaMemcpyHostToDevice, s));
                DLT_CUDA(launch_transform(st, slots_.blocks[slot], sp, nb, s));
                // a full chunk's streams are contiguous in the slot: a staged download is ONE copy (every DMA costs ~10 us)
                const bool whole_slot = nb == chunk_blocks;
                if (!out_pinned && whole_slot) {
                    DLT_CUDA(cudaMemcpyAsync(ctx_->h_out[slot], slots_.streams[slot], nb * bpb, cudaMemcpyDeviceToHost, s));
                    for (int j = 0; j < ns; j++) pend.copy[pend.ncopy++] = {job.out + host_off(j, b0), ctx_->h_out[slot] + slot_off(j), (size_t)w[j] * nb};
                }
                // ... and a payload that is exactly one chunk has the same layout in the slot as in the caller's buffer: ONE
                // copy whatever the number of streams (a BC3 payload of a batch: one download instead of four)
                const bool whole_payload = out_pinned && whole_slot && nb == n;
                if (whole_payload) DLT_CUDA(cudaMemcpyAsync(job.out, slots_.streams[slot], nb * bpb, cudaMemcpyDeviceToHost, s));
                for (int k = 0; k < ns && !(!out_pinned && whole_slot) && !whole_payload; k += run_of[k]) {
                    if (out_pinned && run_of[k] > 1) {   // equal-width neighbours: one strided copy (rows = streams)
                        DLT_CUDA(cudaMemcpy2DAsync(job.out + host_off(k, b0), n * (size_t)w[k], sp.p[k], chunk_blocks * (size_t)w[k],
                                                   (size_t)w[k] * nb, (size_t)run_of[k], cudaMemcpyDeviceToHost, s));
                        continue;
                    }
                    for (int j = k; j < k + run_of[k]; j++) {
                        uint8_t* dst = out_pinned ? job.out + host_off(j, b0) : ctx_->h_out[slot] + slot_off(j);
                        DLT_CUDA(cudaMemcpyAsync(dst, sp.p[j], (size_t)w[j] * nb, cudaMemcpyDeviceToHost, s));
                        if (!out_pinned) pend.copy[pend.ncopy++] = {job.out + host_off(j, b0), dst, (size_t)w[j] * nb};
                    }
                }
            } else {
                const bool whole_slot = nb == chunk_blocks;
                if (!in_pinned) {
                    CopyPool::Seg segs[kMaxStreams];
                    for (int j = 0; j < ns; j++) segs[j] = {ctx_->h_in[slot] + slot_off(j), job.in + host_off(j, b0), (size_t)w[j] * nb};
                    CopyPool::fill().copy_many(segs, ns);
                    if (whole_slot) {
                        DLT_CUDA(cudaMemcpyAsync(slots_.streams[slot], ctx_->h_in[slot], nb * bpb, cudaMemcpyHostToDevice, s));
                    } else {
                        for (int j = 0; j < ns; j++)
                            DLT_CUDA(cudaMemcpyAsync(sp.p[j], ctx_->h_in[slot] + slot_off(j), (size_t)w[j] * nb, cudaMemcpyHostToDevice, s));
                    }
                }
                const bool whole_payload = in_pinned && whole_slot && nb == n;
                if (whole_payload) DLT_CUDA(cudaMemcpyAsync(slots_.streams[slot], job.in, nb * bpb, cudaMemcpyHostToDevice, s));
                for (int k = 0; k < ns && in_pinned && !whole_payload; k += run_of[k]) {
                    if (in_pinned && run_of[k] > 1) {
                        DLT_CUDA(cudaMemcpy2DAsync(sp.p[k], chunk_blocks * (size_t)w[k], job.in + host_off(k, b0), n * (size_t)w[k],
                                                   (size_t)w[k] * nb, (size_t)run_of[k], cudaMemcpyHostToDevice, s));
                        continue;
                    }
                    for (int j = k; j < k + run_of[k]; j++)
                        DLT_CUDA(cudaMemcpyAsync(sp.p[j], job.in + host_off(j, b0), (size_t)w[j] * nb, cudaMemcpyHostToDevice, s));
                }
                DLT_CUDA(launch_untransform(st, sp, slots_.blocks[slot], nb, s));
                uint8_t* dst = out_pinned ? job.out + b0 * bpb : ctx_->h_out[slot];
                DLT_CUDA(cudaMemcpyAsync(dst, slots_.blocks[slot], nb * bpb, cudaMemcpyDeviceToHost, s));
                if (!out_pinned) pend.copy[pend.ncopy++] = {job.out + b0 * bpb, dst, nb * bpb};
            }
            // The host has to wait for this chunk only if it must touch the slot's staging memory again.
            pend.host_wait = !in_pinned || !out_pinned;
            if (pend.host_wait) DLT_CUDA(cudaEventRecord(ctx_->done[slot], s));
            if (pend.ncopy && drainer_.joinable()) queue_drain(slot);
        }
        return Status::kOk;
    }

    // The grouped zero-copy payloads: descriptors to the device, one launch per settings combination; the payloads whose
    // blocks are not 16-byte aligned in host memory are gathered into / scattered from the device arena by one kernel each.
    Status flush_groups() {
        size_t desc = (gather_.size() + scatter_.size()) * sizeof(CopyBatchItem);
        for (const ZeroCopyGroup& g : groups_) desc += g.fwd.size() * sizeof(TransformBatchItem) + g.inv.size() * sizeof(UntransformBatchItem);
        if (desc == 0 && late_.empty()) return Status::kOk;
        if (groups_.size() == 1 && groups_[0].count() == 1 && gather_.empty() && scatter_.empty() && late_.empty()) {
            // a single small payload (the plain synchronous call): no descriptor upload, the ordinary launch
            // One launch over mapped memory first READS the whole payload over the link and then WRITES it (every tile is
            // resident at once and in the same phase): the two directions of the link take turns.  From 1 MiB on the block
            // range is cut into up to eight pieces on the context's streams, so that one piece reads while another writes
            // (block-range launches compose: the ragged kernel never touches bytes outside its range).
            const ZeroCopyGroup& g = groups_[0];
            const bool fwd = !g.fwd.empty();
            const uint64_t n = fwd ? g.fwd[0].nblocks : g.inv[0].nblocks;
            const int bpb = block_bytes(g.st.format);
            const uint64_t tile = (uint64_t)kTileBytes / bpb, tiles = (n + tile - 1) / tile;
            const uint64_t pieces = std::min<uint64_t>(cfg_.zero_copy_pieces, tiles / 32);   // >= 512 KiB each
            if (pieces <= 1) {
                if (fwd) DLT_CUDA(launch_transform(g.st, g.fwd[0].in, g.fwd[0].out, n, ctx_->stream[0]));
                else DLT_CUDA(launch_untransform(g.st, g.inv[0].in, g.inv[0].out, n, ctx_->stream[0]));
                return Status::kOk;
            }
            const int ns = num_streams(g.st.format, g.st.split_alpha, g.st.split_colour);
            const uint64_t per = (tiles + pieces - 1) / pieces * tile;
            int k = 0;
            for (uint64_t b0 = 0; b0 < n; b0 += per, k++) {
                const uint64_t nb = std::min(per, n - b0);
                StreamPtrs sp = fwd ? g.fwd[0].out : g.inv[0].in;
                for (int j = 0; j < ns; j++) sp.p[j] += (uint64_t)stream_width(g.st.format, g.st.split_alpha, g.st.split_colour, j) * b0;
                cudaStream_t s = ctx_->stream[k % cfg_.stages];
                if (fwd) DLT_CUDA(launch_transform(g.st, g.fwd[0].in + b0 * bpb, sp, nb, s));
                else DLT_CUDA(launch_untransform(g.st, sp, g.inv[0].out + b0 * bpb, nb, s));
            }
            return Status::kOk;
        }
        desc = (desc + 255) / 256 * 256;
        const Status st = ensure_scratch(ctx_, desc + arena_bytes_);
        if (st != Status::kOk) return st;
        cudaStream_t s = ctx_->stream[0];
        uint8_t* d = ctx_->d_scratch;
        uint8_t* arena = ctx_->d_scratch + desc;
        auto in_arena = [arena](const uint8_t* off) { return arena + reinterpret_cast<uintptr_t>(off); };
        auto streams_in_arena = [&](StreamPtrs& sp, const Settings& st) {
            for (int k = 0; k < num_streams(st.format, st.split_alpha, st.split_colour); k++) sp.p[k] = in_arena(sp.p[k]);
        };
        auto run_copies = [&](std::vector<CopyBatchItem>& list, bool src_in_arena) -> cudaError_t {
            if (list.empty()) return cudaSuccess;
            uint64_t max_bytes = 0;
            for (CopyBatchItem& c : list) {
                if (src_in_arena) c.src = in_arena(c.src);
                else c.dst = in_arena(c.dst);
                max_bytes = std::max<uint64_t>(max_bytes, c.bytes);
            }
            const size_t nb = list.size() * sizeof(CopyBatchItem);
            cudaError_t e = cudaMemcpyAsync(d, list.data(), nb, cudaMemcpyHostToDevice, s);
            for (size_t at = 0; e == cudaSuccess && at < list.size(); at += 65535)
                e = launch_copy_batch(reinterpret_cast<const CopyBatchItem*>(d) + at, (int)std::min<size_t>(65535, list.size() - at), max_bytes, s);
            d += nb;
            return e;
        };
        DLT_CUDA(run_copies(gather_, false));
        for (ZeroCopyGroup& g : groups_) {
            if (!g.fwd.empty()) {
                for (size_t i = 0; i < g.fwd.size(); i++) {
                    if (g.fwd_arena[i] & 1) g.fwd[i].in = in_arena(g.fwd[i].in);
                    if (g.fwd_arena[i] & 2) streams_in_arena(g.fwd[i].out, g.st);
                }
                const size_t nb = g.fwd.size() * sizeof(TransformBatchItem);
                DLT_CUDA(cudaMemcpyAsync(d, g.fwd.data(), nb, cudaMemcpyHostToDevice, s));
                DLT_CUDA(launch_transform_batch(g.st, reinterpret_cast<const TransformBatchItem*>(d), (int)g.fwd.size(), g.max_blocks,
                                                g.ragged, s));
                d += nb;
            }
            if (!g.inv.empty()) {
                for (size_t i = 0; i < g.inv.size(); i++) {
                    if (g.inv_arena[i] & 1) streams_in_arena(g.inv[i].in, g.st);
                    if (g.inv_arena[i] & 2) g.inv[i].out = in_arena(g.inv[i].out);
                }
                const size_t nb = g.inv.size() * sizeof(UntransformBatchItem);
                DLT_CUDA(cudaMemcpyAsync(d, g.inv.data(), nb, cudaMemcpyHostToDevice, s));
                DLT_CUDA(launch_untransform_batch(g.st, reinterpret_cast<const UntransformBatchItem*>(d), (int)g.inv.size(),
                                                  g.max_blocks, s));
                d += nb;
            }
        }
        for (LateTransform& t : late_) {
            if (t.flags & 2) streams_in_arena(t.item.out, t.st);
            DLT_CUDA(launch_transform(t.st, (t.flags & 1) ? in_arena(t.item.in) : t.item.in, t.item.out, t.item.nblocks, s));
        }
        DLT_CUDA(run_copies(scatter_, true));
        return Status::kOk;
    }

    Status drain() {
        Status result = flush_groups();
        for (int i = 0; i < cfg_.stages; i++) {
            // oldest first: slots are used round-robin starting at seq_ % stages
            const Status f = finish((int)((seq_ + i) % cfg_.stages));
            if (f != Status::kOk && result == Status::kOk) result = f;
        }
        for (int i = 0; i < kStages; i++) {
            const cudaError_t e = cudaStreamSynchronize(ctx_->stream[i]);
            if (e != cudaSuccess && result == Status::kOk) note_cuda_error(e), result = Status::kCudaError;
        }
        stop_drainer();   // idle by now: every slot has been finished
        return result;
    }

private:
    struct Pending {
        bool host_wait = false;
        bool queued = false;   // the drain thread owns the copies (guarded by dm_)
        int ncopy = 0;
        struct {
            uint8_t* dst;
            const uint8_t* src;
            size_t n;
        } copy[kMaxStreams] = {};
    };

    // The host is done with a slot once the chunk that used it has left the device AND its staged output has been
    // copied to the caller's buffer.  Small calls do that copy right here; a large pageable call hands it to a drain
    // thread (start_drainer), so that this thread can already fill the next chunk's input slot: uploads and downloads
    // then overlap on the host side too, instead of taking turns on one thread.
    Status finish(int slot) {
        Pending& p = pending_[slot];
        if (!p.host_wait) return Status::kOk;
        p.host_wait = false;
        if (drainer_.joinable()) {
            std::unique_lock<std::mutex> lk(dm_);
            if (p.queued) {
                dcv_done_.wait(lk, [&] { return !pending_[slot].queued; });
                return drain_status_;
            }
        }
        DLT_CUDA(cudaEventSynchronize(ctx_->done[slot]));
        {
            CopyPool::Seg segs[kMaxStreams];
            for (int i = 0; i < p.ncopy; i++) segs[i] = {p.copy[i].dst, p.copy[i].src, p.copy[i].n};
            CopyPool::fill().copy_many(segs, p.ncopy);
        }
        p.ncopy = 0;
        return Status::kOk;
    }

    void start_drainer() {
        if (drainer_.joinable()) return;
        try {
            drainer_ = std::thread([this] {
                cudaSetDevice(ctx_->device);
                for (;;) {
                    int slot;
                    {
                        std::unique_lock<std::mutex> lk(dm_);
                        dcv_work_.wait(lk, [&] { return dstop_ || dhead_ != dtail_; });
                        if (dhead_ == dtail_) return;   // stop requested and nothing left
                        slot = dqueue_[dhead_ % kDrainQueue];
                    }
                    Pending& p = pending_[slot];
                    Status st = Status::kOk;
                    const cudaError_t e = cudaEventSynchronize(ctx_->done[slot]);
                    if (e != cudaSuccess) note_cuda_error(e), st = Status::kCudaError;
                    else
                    {
                        CopyPool::Seg segs[kMaxStreams];
                        for (int i = 0; i < p.ncopy; i++) segs[i] = {p.copy[i].dst, p.copy[i].src, p.copy[i].n};
                        CopyPool::drain().copy_many(segs, p.ncopy);
                    }
                    {
                        std::lock_guard<std::mutex> lk(dm_);
                        p.ncopy = 0, p.queued = false, dhead_++;
                        if (st != Status::kOk) drain_status_ = st;
                    }
                    dcv_done_.notify_all();
                }
            });
        } catch (...) {
            // no thread: the copies stay on the submitting thread (finish)
        }
    }
    void stop_drainer() {
        if (!drainer_.joinable()) return;
        {
            std::lock_guard<std::mutex> lk(dm_);
            dstop_ = true;
        }
        dcv_work_.notify_all();
        drainer_.join();
    }
    // queue the staged output of `slot` for the drain thread (the slot's event has been recorded)
    void queue_drain(int slot) {
        {
            std::lock_guard<std::mutex> lk(dm_);
            pending_[slot].queued = true;
            dqueue_[dtail_ % kDrainQueue] = slot, dtail_++;
        }
        dcv_work_.notify_all();
    }

    struct ZeroCopyGroup {
        Settings st{};
        bool inverse = false, ragged = false;
        uint64_t max_blocks = 0;
        std::vector<TransformBatchItem> fwd;
        std::vector<UntransformBatchItem> inv;
        std::vector<char> fwd_arena, inv_arena;   // per item: bit 0 = input side, bit 1 = output side is an arena offset
        size_t count() const { return fwd.size() + inv.size(); }
    };
    struct LateTransform {
        Settings st;
        TransformBatchItem item;
        char flags;
    };

    Context* ctx_;
    const HostPathConfig& cfg_;
    Slots slots_{};
    Pending pending_[kStages];
    // drain thread (large pageable calls only)
    static constexpr int kDrainQueue = 2 * kStages;
    std::thread drainer_;
    std::mutex dm_;
    std::condition_variable dcv_work_, dcv_done_;
    int dqueue_[kDrainQueue] = {};
    size_t dhead_ = 0, dtail_ = 0;
    bool dstop_ = false;
    Status drain_status_ = Status::kOk;
    size_t seq_ = 0;
    PinnedRanges pinned_;
    std::vector<JobInfo> info_;
    std::vector<ZeroCopyGroup> groups_;
    std::vector<CopyBatchItem> gather_, scatter_;
    std::vector<LateTransform> late_;
    size_t arena_bytes_ = 0;
};

}  // namespace

Status run_host_batch(const HostJob* jobs, size_t count, int device) {
    bool any = false;
    for (size_t i = 0; i < count; i++) any |= jobs[i].len != 0;
    if (!any) return Status::kOk;
    Status status;
    Context* ctx = acquire_context(device, &status);
    if (!ctx) return status;
    HostPipeline pipe(ctx);
    status = pipe.prepare(jobs, count);
    for (size_t i = 0; i < count && status == Status::kOk; i++) status = pipe.submit(jobs[i], i);
    const Status d = pipe.drain();
    if (status != Status::kOk || d != Status::kOk) release_context_synced(ctx);   // nothing may be in flight in a pooled context
    else release_context(ctx);
    return status != Status::kOk ? status : d;
}

Status run_host(const Settings& st, bool inverse, const uint8_t* in, uint8_t* out, size_t len, int device) {
    const HostJob job{st, inverse, in, out, len};
    return run_host_batch(&job, 1, device);
}

}  // namespace dlt
