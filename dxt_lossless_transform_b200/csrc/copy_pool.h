// copy_pool.h — the host-side staging copies of the pageable path (plain C++, no CUDA: tests/native/copy_pool_stress.cpp
// exercises it on the CPU).  See host_pipeline.cu for how the two pools are used.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif
#if defined(__linux__)
#include <sched.h>
#endif

namespace dlt {

constexpr int kCopyPoolMaxSegs = 6;   // = kMaxStreams (bcn_layout.h): the streams of one chunk

// Copy with non-temporal stores: neither side of a staging copy is read again by this core (the pinned slot is read by
// the DMA engine, the caller's buffer by whoever comes after the call), so the destination lines need not be fetched
// first (a plain store reads the line it is about to overwrite: 3 bytes of DRAM traffic per byte copied instead of 2)
// and must not evict the caches.  glibc only switches to such stores for copies far larger than the 2 MiB parts here.
inline void stream_copy(uint8_t* dst, const uint8_t* src, size_t n) {
#if defined(__x86_64__) || defined(_M_X64)
    if (n < 4096) {
        std::memcpy(dst, src, n);
        return;
    }
    const size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;
    std::memcpy(dst, src, head);
    dst += head, src += head, n -= head;
    const size_t lines = n / 64;
    for (size_t i = 0; i < lines; i++) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 64 * i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 64 * i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 64 * i + 32));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 64 * i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 64 * i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 64 * i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 64 * i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 64 * i + 48), d);
    }
    _mm_sfence();
    std::memcpy(dst + 64 * lines, src + 64 * lines, n - 64 * lines);
#else
    std::memcpy(dst, src, n);
#endif
}

// Staging copies for pageable caller memory: a few worker threads split each large copy so the host side of the
// pipeline keeps up with the link (one thread tops out well below PCIe Gen5).  Two pools: one FILLS the pinned input
// slots (driven by the submitting thread), one DRAINS the pinned output slots (driven by the pipeline's drain thread),
// so uploads and downloads of a pageable call overlap on the host as they do on the link.
// A part is 0.5-3 MiB — 50-300 us of copying — so how fast a worker STARTS matters as much as how fast it copies: waking
// a thread that sleeps on a condition variable costs 50-100 us in a VM.  Workers therefore keep polling for ~200 us after
// their last part (the next chunk of a running pipeline arrives sooner than that) before they go to sleep, and parts are
// handed out through one atomic ticket {generation, parts of the job, next part}: no lock on the copy path.
class CopyPool {
public:
    static CopyPool& fill() {
        static CopyPool* pool = new CopyPool;  // leaked on purpose: workers outlive static destruction
        return *pool;
    }
    static CopyPool& drain() {
        static CopyPool* pool = new CopyPool;
        return *pool;
    }
    struct Seg {
        uint8_t* dst;
        const uint8_t* src;
        size_t n;
    };
    void copy(uint8_t* dst, const uint8_t* src, size_t n) {
        const Seg one{dst, src, n};
        copy_many(&one, 1);
    }
    // Several copies as ONE parallel job (the streams of a chunk): one hand-over to the workers instead of one per stream.
    void copy_many(const Seg* segs, int nseg) {
        constexpr size_t kMinPart = 512u << 10;
        size_t total = 0;
        for (int i = 0; i < nseg; i++) total += segs[i].n;
        if (nseg > kMaxSegs || total / kMinPart <= 1 || workers_ == 0) {
            for (int i = 0; i < nseg; i++) stream_copy(segs[i].dst, segs[i].src, segs[i].n);
            return;
        }
        std::lock_guard<std::mutex> serial(call_mutex_);  // one parallel copy at a time per pool
        // about one part per thread, never below kMinPart, every segment cut into whole parts of that size
        per_ = std::max(kMinPart, (total / (workers_ + 1) + 63) & ~(size_t)63);
        size_t parts = 0;
        nseg_ = nseg;
        for (int i = 0; i < nseg; i++) {
            seg_[i] = segs[i];
            first_part_[i] = parts;
            parts += (segs[i].n + per_ - 1) / per_;
        }
        first_part_[nseg] = parts;
        if (parts == 0) return;
        pending_.store(parts, std::memory_order_relaxed);
        const uint64_t gen = (ticket_.load(std::memory_order_relaxed) >> (2 * kFieldBits)) + 1;
        ticket_.store(gen << (2 * kFieldBits) | (uint64_t)parts << kFieldBits, std::memory_order_seq_cst);   // publishes the job
        if (sleepers_.load(std::memory_order_seq_cst) > 0) {
            { std::lock_guard<std::mutex> lk(m_); }
            cv_.notify_all();
        }
        work();   // the caller copies parts too
        while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();
    }

private:
    static constexpr int kFieldBits = 20;
    static constexpr uint64_t kFieldMask = (1u << kFieldBits) - 1;
    static void cpu_relax() {
#if defined(__x86_64__) || defined(_M_X64)
        _mm_pause();
#endif
    }
    CopyPool() {
        // each pool gets a bit under half of the cores: the two run at the same time
        unsigned hw = std::thread::hardware_concurrency();
#if defined(__linux__)
        {   // the cores this process may use, shared with the other ranks of a one-process-per-GPU job on the same box
            cpu_set_t set;
            if (sched_getaffinity(0, sizeof(set), &set) == 0) hw = std::min<unsigned>(hw, (unsigned)CPU_COUNT(&set));
            if (const char* v = std::getenv("LOCAL_WORLD_SIZE")) {
                const long ranks = std::atol(v);
                if (ranks > 1) hw = std::max(1u, hw / (unsigned)ranks);
            }
        }
#endif
        // (16 cores, threads per pool incl. the caller: 4 / 6 / 8 / 12 -> 22.9 / 24.2 / 28.2 / 19.2 GB/s per direction for a pageable
        // 1 GiB transform: half of the cores each, never more — the workers poll)
        unsigned n = hw >= 16 ? 7 : hw >= 8 ? 3 : hw >= 4 ? 1 : 0;
        if (const char* v = std::getenv("DLTCUDA_COPY_THREADS")) {
            const long t = std::atol(v);
            if (t >= 1 && t <= 64) n = (unsigned)t - 1;
        }
        workers_ = n;
        for (unsigned i = 0; i < n; i++) std::thread([this] { run(); }).detach();
    }
    // claims and copies parts of the current job until none is left; returns the generation it worked on
    uint64_t work() {
        for (;;) {
            const uint64_t t = ticket_.fetch_add(1, std::memory_order_acq_rel);
            const uint64_t next = t & kFieldMask, total = (t >> kFieldBits) & kFieldMask;
            if (next >= total) {
                // nothing left: undo the overshoot's effect on nobody (the field only counts up until the next job resets it;
                // at most one overshoot per thread and job, far from the 2^20 that would carry)
                return t >> (2 * kFieldBits);
            }
            // a valid part: the job is unfinished, so its parameters are the ones published with this ticket
            int i = 0;
            while (next >= first_part_[i + 1]) i++;
            const size_t off = (next - first_part_[i]) * per_;
            stream_copy(seg_[i].dst + off, seg_[i].src + off, std::min(per_, seg_[i].n - off));
            pending_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    void run() {
        uint64_t seen = 0;
        for (;;) {
            // wait for a job of a generation this thread has not finished yet: poll first, then sleep
            uint32_t spins = 0;
            while ((ticket_.load(std::memory_order_acquire) >> (2 * kFieldBits)) == seen) {
                if (++spins < 40000) {   // ~200 us
                    cpu_relax();
                    continue;
                }
                std::unique_lock<std::mutex> lk(m_);
                sleepers_.fetch_add(1, std::memory_order_seq_cst);
                cv_.wait(lk, [&] { return (ticket_.load(std::memory_order_seq_cst) >> (2 * kFieldBits)) != seen; });
                sleepers_.fetch_sub(1, std::memory_order_seq_cst);
                spins = 0;
            }
            seen = work();
        }
    }
    size_t workers_ = 0;
    std::mutex m_, call_mutex_;
    std::condition_variable cv_;
    std::atomic<uint64_t> ticket_{0};   // generation << 40 | parts << 20 | next part
    std::atomic<size_t> pending_{0};
    std::atomic<int> sleepers_{0};
    static constexpr int kMaxSegs = kCopyPoolMaxSegs;
    Seg seg_[kMaxSegs] = {};
    size_t first_part_[kMaxSegs + 1] = {};
    int nseg_ = 0;
    size_t per_ = 0;
};

}  // namespace dlt
