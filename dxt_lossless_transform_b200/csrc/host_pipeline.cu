// host_pipeline.cu — see host_pipeline.h.
#include "host_pipeline.h"

#include <cstring>
#include <mutex>

#include "bcn_kernels.h"

namespace dlt {
namespace {

constexpr int kMaxDevices = 64;
constexpr size_t kSlotBytes = kChunkBytes + 256 * kMaxStreams;

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
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        note_cuda_error(e);
        *st = Status::kCudaError;
        return nullptr;
    }
    {
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        if (Context* c = g_free[device]) {
            g_free[device] = c->next_free;
            c->next_free = nullptr;
            return c;
        }
    }
    Context* c = nullptr;
    *st = create_context(device, &c);
    return c;
}

void release_context(Context* ctx) {
    if (!ctx) return;
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    ctx->next_free = g_free[ctx->device];
    g_free[ctx->device] = ctx;
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

Status ensure_staging(Context* ctx) {
    for (int i = 0; i < kStages; i++) {
        if (!ctx->h_in[i]) DLT_CUDA(cudaMallocHost(&ctx->h_in[i], kSlotBytes));
        if (!ctx->h_out[i]) DLT_CUDA(cudaMallocHost(&ctx->h_out[i], kSlotBytes));
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

// Device-side chunk slots: blocks in one buffer, the chunk's streams compacted in another with
// every stream 256-byte aligned (chunk block counts are powers of two), so the tiled kernels always
// run their aligned path no matter what N is.
struct Slots {
    uint8_t* blocks[kStages];
    uint8_t* streams[kStages];
};

Status ensure_slots(Context* ctx, Slots* s) {
    // d_in / d_out double as the slot arena: kStages slots each.
    Status st = ensure_device_buffers(ctx, kSlotBytes * kStages);
    if (st != Status::kOk) return st;
    for (int i = 0; i < kStages; i++) {
        s->blocks[i] = ctx->d_in + kSlotBytes * i;
        s->streams[i] = ctx->d_out + kSlotBytes * i;
    }
    return Status::kOk;
}

}  // namespace

Status run_host(const Settings& st, bool inverse, const uint8_t* in, uint8_t* out, size_t len, int device) {
    if (len == 0) return Status::kOk;
    Status status;
    Context* ctx = acquire_context(device, &status);
    if (!ctx) return status;
    struct Releaser {
        Context* c;
        ~Releaser() { release_context(c); }
    } releaser{ctx};

    Slots slots;
    if ((status = ensure_slots(ctx, &slots)) != Status::kOk) return status;

    const int bpb = block_bytes(st.format);
    const int ns = num_streams(st.format, st.split_alpha, st.split_colour);
    const size_t n = len / bpb;
    const size_t chunk_blocks = kChunkBytes / bpb;
    const size_t nchunks = (n + chunk_blocks - 1) / chunk_blocks;
    int w[kMaxStreams], pre[kMaxStreams];
    for (int k = 0; k < ns; k++) {
        w[k] = stream_width(st.format, st.split_alpha, st.split_colour, k);
        pre[k] = stream_prefix(st.format, st.split_alpha, st.split_colour, k);
    }

    const bool in_pinned = is_pinned_host(in) && is_pinned_host(in + len - 1);
    const bool out_pinned = is_pinned_host(out) && is_pinned_host(out + len - 1);
    if (!in_pinned || !out_pinned)
        if ((status = ensure_staging(ctx)) != Status::kOk) return status;

    // Host offset of stream k, block b of the payload (reference layout); slot offset of the same
    // element inside a chunk that starts at block b0.
    auto host_off = [&](int k, size_t b) { return n * (size_t)pre[k] + (size_t)w[k] * b; };
    auto slot_off = [&](int k) { return chunk_blocks * (size_t)pre[k]; };

    auto issue = [&](size_t c) -> Status {
        const int slot = (int)(c % kStages);
        cudaStream_t s = ctx->stream[slot];
        const size_t b0 = c * chunk_blocks;
        const size_t nb = n - b0 < chunk_blocks ? n - b0 : chunk_blocks;
        StreamPtrs sp{};
        for (int k = 0; k < ns; k++) sp.p[k] = slots.streams[slot] + slot_off(k);
        if (!inverse) {
            const uint8_t* src = in + b0 * bpb;
            if (!in_pinned) {
                std::memcpy(ctx->h_in[slot], src, nb * bpb);
                src = ctx->h_in[slot];
            }
            DLT_CUDA(cudaMemcpyAsync(slots.blocks[slot], src, nb * bpb, cudaMemcpyHostToDevice, s));
            DLT_CUDA(launch_transform(st, slots.blocks[slot], sp, nb, s));
            for (int k = 0; k < ns; k++) {
                uint8_t* dst = out_pinned ? out + host_off(k, b0) : ctx->h_out[slot] + slot_off(k);
                DLT_CUDA(cudaMemcpyAsync(dst, sp.p[k], (size_t)w[k] * nb, cudaMemcpyDeviceToHost, s));
            }
        } else {
            for (int k = 0; k < ns; k++) {
                const uint8_t* src = in + host_off(k, b0);
                if (!in_pinned) {
                    std::memcpy(ctx->h_in[slot] + slot_off(k), src, (size_t)w[k] * nb);
                    src = ctx->h_in[slot] + slot_off(k);
                }
                DLT_CUDA(cudaMemcpyAsync(sp.p[k], src, (size_t)w[k] * nb, cudaMemcpyHostToDevice, s));
            }
            DLT_CUDA(launch_untransform(st, sp, slots.blocks[slot], nb, s));
            uint8_t* dst = out_pinned ? out + b0 * bpb : ctx->h_out[slot];
            DLT_CUDA(cudaMemcpyAsync(dst, slots.blocks[slot], nb * bpb, cudaMemcpyDeviceToHost, s));
        }
        DLT_CUDA(cudaEventRecord(ctx->done[slot], s));
        return Status::kOk;
    };

    auto finish = [&](size_t c) -> Status {
        const int slot = (int)(c % kStages);
        DLT_CUDA(cudaEventSynchronize(ctx->done[slot]));
        if (out_pinned) return Status::kOk;
        const size_t b0 = c * chunk_blocks;
        const size_t nb = n - b0 < chunk_blocks ? n - b0 : chunk_blocks;
        if (!inverse) {
            for (int k = 0; k < ns; k++)
                std::memcpy(out + host_off(k, b0), ctx->h_out[slot] + slot_off(k), (size_t)w[k] * nb);
        } else {
            std::memcpy(out + b0 * bpb, ctx->h_out[slot], nb * bpb);
        }
        return Status::kOk;
    };

    Status result = Status::kOk;
    for (size_t c = 0; c < nchunks + kStages; c++) {
        if (c >= (size_t)kStages && c - kStages < nchunks) {
            Status f = finish(c - kStages);
            if (f != Status::kOk && result == Status::kOk) result = f;
        }
        if (c < nchunks && result == Status::kOk) {
            Status f = issue(c);
            if (f != Status::kOk) result = f;
        }
    }
    if (result != Status::kOk) {
        for (int i = 0; i < kStages; i++) (void)cudaStreamSynchronize(ctx->stream[i]);
    }
    return result;
}

}  // namespace dlt
