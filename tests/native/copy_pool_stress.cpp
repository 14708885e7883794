// CPU stress test of csrc/copy_pool.h (the staging-copy pools of the pageable host path): many jobs of random shapes,
// from two caller threads on two pools at once, every byte checked.  Built and run by tests/test_copy_pool.py.
#include <cstdio>
#include <random>
#include <vector>

#include "copy_pool.h"

using dlt::CopyPool;

static int run(CopyPool& pool, unsigned seed, int jobs) {
    std::mt19937_64 rng(seed);
    const size_t cap = 24u << 20;
    std::vector<uint8_t> src(cap), dst(cap);
    for (size_t i = 0; i < cap; i++) src[i] = (uint8_t)(rng() >> 56);
    for (int j = 0; j < jobs; j++) {
        const int nseg = 1 + (int)(rng() % 6);
        CopyPool::Seg segs[6];
        size_t at = 0;
        std::fill(dst.begin(), dst.end(), (uint8_t)0xEE);
        std::vector<std::pair<size_t, size_t>> spans;
        for (int s = 0; s < nseg; s++) {
            // sizes from a few bytes to a few MiB, odd offsets on both sides
            const size_t kind = rng() % 4;
            size_t n = kind == 0 ? rng() % 5000 : kind == 1 ? (512u << 10) + rng() % 4096 : kind == 2 ? rng() % (6u << 20) : 0;
            const size_t gap = rng() % 131;
            if (at + gap + n > cap) break;
            at += gap;
            segs[s] = CopyPool::Seg{dst.data() + at, src.data() + at, n};
            spans.emplace_back(at, n);
            at += n;
        }
        pool.copy_many(segs, (int)spans.size());
        size_t pos = 0;
        for (auto [o, n] : spans) {
            for (size_t i = pos; i < o; i++)
                if (dst[i] != 0xEE) return std::printf("job %d: byte %zu outside the segments was written\n", j, i), 1;
            if (std::memcmp(dst.data() + o, src.data() + o, n) != 0) return std::printf("job %d: segment at %zu (%zu bytes) differs\n", j, o, n), 1;
            pos = o + n;
        }
        for (size_t i = pos; i < std::min(cap, pos + 4096); i++)
            if (dst[i] != 0xEE) return std::printf("job %d: byte %zu after the segments was written\n", j, i), 1;
        if (j % 64 == 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % 600));   // let the workers fall asleep now and then
    }
    return 0;
}

int main(int argc, char** argv) {
    const int jobs = argc > 1 ? std::atoi(argv[1]) : 300;
    int rc_a = 0, rc_b = 0;
    std::thread a([&] { rc_a = run(CopyPool::fill(), 1, jobs); });
    std::thread b([&] { rc_b = run(CopyPool::drain(), 2, jobs); });
    a.join(), b.join();
    if (rc_a || rc_b) return 1;
    // one pool, two callers (concurrent host threads share the pools)
    std::thread c([&] { rc_a = run(CopyPool::fill(), 3, jobs / 2); });
    std::thread d([&] { rc_b = run(CopyPool::fill(), 4, jobs / 2); });
    c.join(), d.join();
    std::puts(rc_a || rc_b ? "FAILED" : "copy pool ok");
    return rc_a || rc_b;
}
