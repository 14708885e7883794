// Microbenchmark: MATCH.ANY latency / throughput on sm_100a (design input for the estimator's window resolution).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o match_bench match_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>   // 0: dependent chain (latency), 1: 8 independent per iteration (throughput)
__global__ void k(uint32_t* out, int iters, uint32_t distinct_mask, unsigned member_mask_kind) {
    const unsigned lane = threadIdx.x & 31;
    unsigned mm = 0xffffffffu;
    if (member_mask_kind == 1) mm = 0xffu << (lane & ~7u);   // 8-lane groups
    uint32_t v[8];
    for (int r = 0; r < 8; r++) v[r] = (lane * 2654435761u + r * 40503u + blockIdx.x) & distinct_mask;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {
            uint32_t m = __match_any_sync(mm, v[0]);
            v[0] = (v[0] + (m & 1)) & distinct_mask;
            acc += m;
        } else {
            uint32_t m[8];
#pragma unroll
            for (int r = 0; r < 8; r++) m[r] = __match_any_sync(mm, v[r]);
#pragma unroll
            for (int r = 0; r < 8; r++) { acc += m[r]; v[r] = (v[r] + (m[r] & 1)) & distinct_mask; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (uint32_t)(t1 - t0); }
    if (acc == 0x12345) out[1] = acc;
}

int main() {
    uint32_t* d; cudaMalloc(&d, 64);
    const int iters = 2000;
    for (unsigned kind = 0; kind < 2; kind++)
    for (uint32_t dm : {0xffffffffu, 0x3u, 0x0u}) {
        for (int warps : {1, 4, 8, 12, 16}) {
            uint32_t h[2];
            k<0><<<148, warps * 32>>>(d, iters, dm, kind); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
            double lat = (double)h[0] / iters;
            k<1><<<148, warps * 32>>>(d, iters, dm, kind); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
            double thr = (double)h[0] / (iters * 8);
            printf("mask_kind=%u distinct_mask=%08x warps/SM=%2d  chain: %.1f cyc/match(warp)   indep: %.1f cyc/match(warp) -> SM-wide %.1f cyc/match\n",
                   kind, dm, warps, lat, thr, thr / warps);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
