// auto_search.cu — see auto_search.h.
#include "auto_search.h"

#include "bcn_kernels.h"
#include "estimator.h"

namespace dlt {

int candidate_order(int format, bool use_all, Settings out[kMaxCandidates]) {
    // (variant, split_colour) — bc1/src/transform/settings.rs:81-98; bc2 identical.
    static const int fast12[4][2] = {{kNone, 0}, {kNone, 1}, {kVariant1, 0}, {kVariant1, 1}};
    static const int all12[8][2] = {{kVariant2, 0}, {kNone, 0},     {kNone, 1},     {kVariant3, 0},
                                    {kVariant3, 1}, {kVariant2, 1}, {kVariant1, 0}, {kVariant1, 1}};
    // (variant, split_alpha, split_colour) — bc3/src/transform/settings.rs:91-121.
    static const int fast3[8][3] = {{kVariant1, 1, 0}, {kVariant1, 1, 1}, {kNone, 1, 0}, {kNone, 0, 1},
                                    {kNone, 1, 1},     {kVariant1, 0, 1}, {kNone, 0, 0}, {kVariant1, 0, 0}};
    static const int all3[16][3] = {{kVariant2, 1, 0}, {kVariant2, 1, 1}, {kVariant3, 1, 1}, {kVariant3, 1, 0},
                                    {kVariant1, 1, 0}, {kVariant3, 0, 1}, {kVariant1, 1, 1}, {kVariant2, 0, 1},
                                    {kVariant2, 0, 0}, {kVariant3, 0, 0}, {kNone, 1, 0},     {kNone, 0, 1},
                                    {kNone, 1, 1},     {kVariant1, 0, 1}, {kNone, 0, 0},     {kVariant1, 0, 0}};
    if (format == 3) {
        const int k = use_all ? 16 : 8;
        for (int i = 0; i < k; i++) {
            const int* c = use_all ? all3[i] : fast3[i];
            out[i] = Settings{3, c[0], c[1] != 0, c[2] != 0};
        }
        return k;
    }
    const int k = use_all ? 8 : 4;
    for (int i = 0; i < k; i++) {
        const int* c = use_all ? all12[i] : fast12[i];
        out[i] = Settings{format, c[0], false, c[1] != 0};
    }
    return k;
}

Settings default_settings(int format) { return Settings{format, kVariant1, format == 3, true}; }

int estimate_ranges(int format, size_t len, EstimateRange out[2]) {
    if (format == 1) {
        out[0] = {0, len / 2};
        return 1;
    }
    if (format == 2) {
        out[0] = {len / 2, len / 4};
        return 1;
    }
    const size_t n = len / 16;
    out[0] = {0, n * 2};
    out[1] = {len / 2, n * 4};
    return 2;
}

Status auto_ltu_device(Context* ctx, int format, const uint8_t* d_in, uint8_t* d_out, size_t len, bool use_all,
                       Settings* best, size_t* sizes, cudaStream_t stream) {
    Settings order[kMaxCandidates];
    const int k = candidate_order(format, use_all, order);
    const size_t n = len / block_bytes(format);
    EstimateRange ranges[2];
    const int nr = estimate_ranges(format, len, ranges);

    Settings best_s = default_settings(format);
    size_t best_size = SIZE_MAX;
    int best_i = -1;
    for (int i = 0; i < k; i++) {
        size_t total = 0;
        if (len != 0) {
            cudaError_t e = launch_transform(order[i], d_in, reference_layout(d_out, n, 0, order[i]), n, stream);
            if (e != cudaSuccess) {
                note_cuda_error(e);
                return Status::kCudaError;
            }
            LtuSegment segs[2];
            uint64_t matches[2] = {0, 0};
            for (int r = 0; r < nr; r++) segs[r] = LtuSegment{d_out + ranges[r].offset, ranges[r].len};
            Status st = ltu_matches_device(ctx, segs, nr, matches, stream);
            if (st != Status::kOk) return st;
            for (int r = 0; r < nr; r++) total += ltu_estimate_from_matches(ranges[r].len, matches[r]);
        }
        if (sizes) sizes[i] = total;
        if (total < best_size) {
            best_size = total;
            best_s = order[i];
            best_i = i;
        }
    }
    if (len != 0 && best_i != k - 1) {
        cudaError_t e = launch_transform(best_s, d_in, reference_layout(d_out, n, 0, best_s), n, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
            note_cuda_error(e);
            return Status::kCudaError;
        }
    }
    *best = best_s;
    return Status::kOk;
}

}  // namespace dlt
