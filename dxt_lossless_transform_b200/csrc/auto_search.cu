// auto_search.cu — see auto_search.h.
#include "auto_search.h"

#include "bcn_kernels.h"
#include "estimator.h"

#include <algorithm>
#include <vector>

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

// Which candidates really differ in an estimated range?  The bytes of a range depend on part of the settings only:
// the BC3 alpha-endpoint range [0, 2N) on split_alpha alone, every colour range on (variant, split_colour, normalize).
// The reference transforms and estimates every candidate (BC3: 8 / 16 x two ranges); estimates are a function of the
// bytes, so estimating each DISTINCT range once and adding the shared results per candidate gives the same totals:
// BC3 needs 2 alpha + 4 / 8 colour estimates and 4-5 / 9 transforms instead of 16 / 32 and 8 / 16.
DistinctPlan plan_distinct(int format, const Settings* order, int k, int nr) {
    DistinctPlan p;
    p.seg_of.assign((size_t)k * nr, -1);
    auto key_of = [format](const Settings& s, int r) {
        if (format == 3 && r == 0) return (int)s.split_alpha;
        return ((s.variant * 2 + (int)s.split_colour) << 3) | s.normalize;
    };
    std::vector<char> need((size_t)k, 0);
    for (int i = 0; i < k; i++)
        for (int r = 0; r < nr; r++) {
            int found = -1;
            for (size_t j = 0; j < p.segs.size() && found < 0; j++)
                if (p.segs[j].range == r && key_of(order[p.segs[j].cand], r) == key_of(order[i], r)) found = (int)j;
            if (found < 0) {
                found = (int)p.segs.size();
                p.segs.push_back({i, r});
                need[i] = 1;
            }
            p.seg_of[(size_t)i * nr + r] = found;
        }
    for (int i = 0; i < k; i++)
        if (need[i]) p.images.push_back(i);
    return p;
}

// A candidate's image is only looked at in its estimated ranges: the colour-endpoint streams (and BC3's alpha-endpoint
// streams).  The tiled kernels skip a stream whose pointer is null, so the candidates do not write their index streams
// (half of a BC1 image, ten of BC3's sixteen bytes per block).  Only where the tiled kernels will run: the bytewise
// fall-back for misaligned pointers writes every stream.
static StreamPtrs estimated_streams_only(const Settings& st, const uint8_t* in, const StreamPtrs& all, uint64_t nblocks) {
    bool ragged = false;
    if (!transform_batch_item_ok(st, TransformBatchItem{in, all, nblocks}, &ragged)) return all;
    StreamPtrs sp = all;
    const int ns = num_streams(st.format, st.split_alpha, st.split_colour);
    sp.p[ns - 1] = nullptr;                                        // the colour indices
    if (st.format == 2) sp.p[0] = nullptr;                         // BC2: the explicit alpha
    if (st.format == 3) sp.p[st.split_alpha ? 2 : 1] = nullptr;    // BC3: the alpha indices
    return sp;
}

// Transforms the candidates of one device-resident payload that hold a distinct estimated range into scratch images and
// estimates those ranges (in batches that fit the scratch budget): totals[i] = estimate of candidate i.
static Status estimate_candidates(Context* ctx, int format, const uint8_t* d_in, size_t len, const Settings* order, int k,
                                  size_t* totals, cudaStream_t stream) {
    const size_t n = len / block_bytes(format);
    EstimateRange ranges[2];
    const int nr = estimate_ranges(format, len, ranges);
    for (int i = 0; i < k; i++) totals[i] = 0;
    if (len == 0) return Status::kOk;
    const DistinctPlan plan = plan_distinct(format, order, k, nr);
    const int nimg = (int)plan.images.size();
    // scratch = [m images][estimator buffers for their distinct ranges]; all images at once when they fit
    const size_t img = (len + 255) / 256 * 256;
    constexpr size_t kScratchBudget = (size_t)12 << 30;
    std::vector<LtuSegment> segs;
    auto scratch_for = [&](int m) {   // worst case: every range of m images is distinct
        segs.clear();
        for (int c = 0; c < m; c++)
            for (int r = 0; r < nr; r++) segs.push_back(LtuSegment{nullptr, ranges[r].len});
        return (size_t)m * img + ltu_scratch_bytes(segs.data(), (int)segs.size());
    };
    int m = nimg;
    while (m > 1 && scratch_for(m) > kScratchBudget) m--;
    Status st;
    while ((st = ensure_scratch(ctx, scratch_for(m))) == Status::kOutOfMemory && m > 1) m = (m + 1) / 2;
    if (st != Status::kOk) return st;
    uint8_t* est_scratch = ctx->d_scratch + (size_t)m * img;
    const size_t est_bytes = ctx->d_scratch_cap - (size_t)m * img;

    std::vector<size_t> seg_estimate(plan.segs.size(), 0);
    for (int c0 = 0; c0 < nimg; c0 += m) {
        const int mb = nimg - c0 < m ? nimg - c0 : m;
        segs.clear();
        std::vector<int> which;   // index into plan.segs of every segment of this batch
        // The endpoint streams of all candidates of the batch from ONE read of the blocks (eight candidates per launch)
        // where that kernel applies: no normalization, 16-byte aligned blocks.  Otherwise one transform per candidate.
        bool fused = (reinterpret_cast<uintptr_t>(d_in) & 15) == 0;
        for (int c = 0; c < mb; c++) fused = fused && order[plan.images[c0 + c]].normalize == kNormNone;
        if (fused) {
            EndpointCandidate ec[kMaxEndpointCandidates];
            int nec = 0;
            for (int c = 0; c <= mb; c++) {
                if (nec == kMaxEndpointCandidates || (c == mb && nec)) {
                    const cudaError_t e = launch_endpoint_candidates(format, d_in, n, ec, nec, stream);
                    if (e != cudaSuccess) {
                        note_cuda_error(e);
                        return Status::kCudaError;
                    }
                    nec = 0;
                }
                if (c == mb) break;
                const Settings& st_c = order[plan.images[c0 + c]];
                uint8_t* image = ctx->d_scratch + (size_t)c * img;
                // the colour range is the last estimated range of every format; BC3's first one holds the alpha endpoints
                ec[nec++] = EndpointCandidate{image + ranges[nr - 1].offset, format == 3 ? image + ranges[0].offset : nullptr, st_c.variant,
                                              st_c.split_colour, st_c.split_alpha};
            }
        }
        for (int c = 0; c < mb; c++) {
            const int cand = plan.images[c0 + c];
            uint8_t* image = ctx->d_scratch + (size_t)c * img;
            if (!fused) {
                cudaError_t e = launch_transform(order[cand], d_in, estimated_streams_only(order[cand], d_in, reference_layout(image, n, 0, order[cand]), n), n, stream);
                if (e != cudaSuccess) {
                    note_cuda_error(e);
                    return Status::kCudaError;
                }
            }
            for (size_t j = 0; j < plan.segs.size(); j++)
                if (plan.segs[j].cand == cand) {
                    const EstimateRange& rg = ranges[plan.segs[j].range];
                    segs.push_back(LtuSegment{image + rg.offset, rg.len});
                    which.push_back((int)j);
                }
        }
        std::vector<uint64_t> matches(segs.size(), 0);
        st = ltu_matches_device(segs.data(), (int)segs.size(), matches.data(), stream, est_scratch, est_bytes);
        if (st != Status::kOk) return st;
        for (size_t q = 0; q < which.size(); q++) seg_estimate[which[q]] = ltu_estimate_from_matches(segs[q].len, matches[q]);
    }
    for (int i = 0; i < k; i++)
        for (int r = 0; r < nr; r++) totals[i] += seg_estimate[plan.seg_of[(size_t)i * nr + r]];
    return Status::kOk;
}

static Status final_transform(const Settings& best, const uint8_t* d_in, uint8_t* d_out, size_t len, cudaStream_t stream) {
    if (len == 0) return Status::kOk;
    const size_t n = len / block_bytes(best.format);
    cudaError_t e = launch_transform(best, d_in, reference_layout(d_out, n, 0, best), n, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        note_cuda_error(e);
        return Status::kCudaError;
    }
    return Status::kOk;
}

Status auto_ltu_device(Context* ctx, int format, const uint8_t* d_in, uint8_t* d_out, size_t len, bool use_all,
                       Settings* best, size_t* sizes, cudaStream_t stream) {
    Settings order[kMaxCandidates];
    const int k = candidate_order(format, use_all, order);
    size_t totals[kMaxCandidates] = {};
    const Status st = estimate_candidates(ctx, format, d_in, len, order, k, totals, stream);
    if (st != Status::kOk) return st;
    // strict '<': the first candidate in test order wins ties (transform_auto.rs:257-260)
    Settings best_s = default_settings(format);
    size_t best_size = SIZE_MAX;
    for (int i = 0; i < k; i++) {
        if (sizes) sizes[i] = totals[i];
        if (totals[i] < best_size) best_size = totals[i], best_s = order[i];
    }
    *best = best_s;
    return final_transform(best_s, d_in, d_out, len, stream);
}

// experimental::transform_bc1_auto_with_normalization (normalize_blocks/transform.rs:222-340): if no block is
// normalizable the plain search runs; otherwise every (normalization mode, settings) pair — modes outermost, in
// ColorNormalizationMode::all_values() order, settings in the usual test order — is estimated over the colour half
// and the first minimum wins.  Normalization is fused into the candidate transforms: there is no normalized copy.
Status auto_ltu_norm_device(Context* ctx, const uint8_t* d_in, uint8_t* d_out, size_t len, bool use_all, Settings* best,
                            size_t* sizes, cudaStream_t stream) {
    const size_t n = len / 8;
    Status st = ensure_scratch(ctx, 256);
    if (st != Status::kOk) return st;
    unsigned int* d_any = reinterpret_cast<unsigned int*>(ctx->d_scratch);
    unsigned int any = 0;
    cudaError_t e = cudaMemsetAsync(d_any, 0, sizeof(unsigned int), stream);
    if (e == cudaSuccess) e = launch_normalize_blocks(d_in, nullptr, nullptr, nullptr, n, d_any, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&any, d_any, sizeof(any), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        note_cuda_error(e);
        return Status::kCudaError;
    }
    if (!any) return auto_ltu_device(ctx, 1, d_in, d_out, len, use_all, best, sizes, stream);

    Settings base[kMaxCandidates], order[3 * kMaxCandidates];
    const int k = candidate_order(1, use_all, base);
    for (int m = 0; m < 3; m++)
        for (int i = 0; i < k; i++) {
            order[m * k + i] = base[i];
            // the `None` candidates are estimated on normalize_blocks_all_modes' None buffer (transparent blocks -> 0xFF)
            order[m * k + i].normalize = m == kNormNone ? (int)kNormAllModesNone : m;
        }
    size_t totals[3 * kMaxCandidates] = {};
    if ((st = estimate_candidates(ctx, 1, d_in, len, order, 3 * k, totals, stream)) != Status::kOk) return st;
    Settings best_s = default_settings(1);   // Bc1TransformDetailsWithNormalization::default(): (None, Variant1, split)
    size_t best_size = SIZE_MAX;
    for (int i = 0; i < 3 * k; i++) {
        if (sizes) sizes[i] = totals[i];
        if (totals[i] < best_size) best_size = totals[i], best_s = order[i];
    }
    if (best_s.normalize == kNormAllModesNone) best_s.normalize = kNormNone;   // ... but transformed without normalization
    *best = best_s;
    return final_transform(best_s, d_in, d_out, len, stream);
}

// Queues the transforms of many (payload, settings) pairs with ONE launch per distinct settings combination
// (launch_transform_batch); the item descriptors go through `d_desc` (device, room for `n` items).  Pairs the tiled
// kernels cannot take (misaligned pointers) are launched one by one.
struct BatchedTransform {
    Settings st;
    const uint8_t* in;
    uint8_t* image;   // reference layout for `len` bytes
    size_t len;
    bool estimate_only = false;   // a candidate: only its estimated streams are needed
};
static cudaError_t queue_transforms(const BatchedTransform* work, int n, TransformBatchItem* d_desc, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    struct Group {
        Settings st;
        std::vector<TransformBatchItem> items;
        uint64_t max_blocks = 0;
        bool ragged = false;
    };
    std::vector<Group> groups;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < n && e == cudaSuccess; i++) {
        const BatchedTransform& w = work[i];
        const size_t nb = w.len / block_bytes(w.st.format);
        if (nb == 0) continue;
        TransformBatchItem item{w.in, reference_layout(w.image, nb, 0, w.st), nb};
        if (w.estimate_only) item.out = estimated_streams_only(w.st, w.in, item.out, nb);
        bool ragged = false;
        if (!transform_batch_item_ok(w.st, item, &ragged)) {
            e = launch_transform(w.st, w.in, item.out, nb, stream);
            continue;
        }
        Group* g = nullptr;
        for (Group& c : groups)
            if (c.st.format == w.st.format && c.st.variant == w.st.variant && c.st.split_alpha == w.st.split_alpha &&
                c.st.split_colour == w.st.split_colour && c.items.size() < 65535)
                g = &c;
        if (!g) groups.emplace_back(), g = &groups.back(), g->st = w.st;
        g->items.push_back(item);
        g->max_blocks = std::max<uint64_t>(g->max_blocks, nb);
        g->ragged |= ragged;
    }
    if (e != cudaSuccess) return e;
    std::vector<TransformBatchItem> all;
    for (const Group& g : groups) all.insert(all.end(), g.items.begin(), g.items.end());
    if (all.empty()) return cudaSuccess;
    e = cudaMemcpyAsync(d_desc, all.data(), all.size() * sizeof(TransformBatchItem), cudaMemcpyHostToDevice, stream);
    size_t at = 0;
    for (const Group& g : groups) {
        if (e != cudaSuccess) break;
        e = launch_transform_batch(g.st, d_desc + at, (int)g.items.size(), g.max_blocks, g.ragged, stream);
        at += g.items.size();
    }
    return e;
}

Status auto_ltu_device_batch(Context* ctx, AutoJob* jobs, int njobs, bool use_all, cudaStream_t stream) {
    constexpr size_t kScratchBudget = (size_t)12 << 30;
    auto desc_bytes = [](size_t ncands) { return (ncands * sizeof(TransformBatchItem) + 255) / 256 * 256; };
    auto cuda_fail = [](cudaError_t e) {
        note_cuda_error(e);
        return Status::kCudaError;
    };
    struct Cand {
        int job, index;      // candidate `index` of job `job` (only candidates that hold a distinct estimated range)
        uint8_t* image;      // its transformed image in the scratch
    };
    struct JobSegs {
        DistinctPlan plan;   // which ranges of which candidates are distinct (plan_distinct)
        int first_seg = 0;   // the job's distinct ranges are segs[first_seg ...)
        int nr = 0;
    };
    int j0 = 0;
    while (j0 < njobs) {
        // ---- the largest group [j0, j1) whose images + estimator scratch fit the budget
        std::vector<LtuSegment> segs;
        std::vector<Cand> cands;
        std::vector<JobSegs> plans;   // one per job of the group, index j - j0
        LtuScratchMeter meter;
        size_t image_bytes = 0;
        int j1 = j0;
        for (; j1 < njobs; j1++) {
            const AutoJob& job = jobs[j1];
            JobSegs js;
            if (job.len == 0) {
                plans.push_back(js);
                continue;
            }
            Settings order[kMaxCandidates];
            const int k = candidate_order(job.format, use_all, order);
            EstimateRange ranges[2];
            js.nr = estimate_ranges(job.format, job.len, ranges);
            js.plan = plan_distinct(job.format, order, k, js.nr);
            js.first_seg = (int)segs.size();
            const size_t nimg = js.plan.images.size();
            const size_t img = (job.len + 255) / 256 * 256;
            LtuScratchMeter with_job = meter;
            for (const DistinctPlan::Seg& sg : js.plan.segs) with_job.add(ranges[sg.range].len);
            const size_t need = image_bytes + nimg * img + desc_bytes(cands.size() + nimg) + with_job.bytes();
            if (need > kScratchBudget && j1 > j0) break;   // this job starts the next group
            meter = with_job;
            for (int c : js.plan.images) cands.push_back(Cand{j1, c, nullptr});
            for (const DistinctPlan::Seg& sg : js.plan.segs) segs.push_back(LtuSegment{nullptr, ranges[sg.range].len});
            image_bytes += nimg * img;
            plans.push_back(std::move(js));
        }
        if (!cands.empty() && image_bytes + desc_bytes(cands.size()) + meter.bytes() > kScratchBudget) {
            // a single job that is too large for one group: the single-payload path batches its candidates itself
            AutoJob& job = jobs[j0];
            const Status st = auto_ltu_device(ctx, job.format, job.d_in, job.d_out, job.len, use_all, &job.best, job.sizes, stream);
            if (st != Status::kOk) return st;
            j0 = j0 + 1;
            continue;
        }
        const size_t est_offset = image_bytes + desc_bytes(std::max<size_t>(cands.size(), (size_t)(j1 - j0)));
        Status st = ensure_scratch(ctx, est_offset + meter.bytes());
        if (st != Status::kOk) return st;
        TransformBatchItem* d_desc = reinterpret_cast<TransformBatchItem*>(ctx->d_scratch + image_bytes);

        // ---- transform the candidates that hold a distinct range, point the segments at those ranges
        uint8_t* next_image = ctx->d_scratch;
        std::vector<BatchedTransform> work;
        work.reserve(cands.size());
        for (Cand& cd : cands) {
            const AutoJob& job = jobs[cd.job];
            const JobSegs& js = plans[(size_t)(cd.job - j0)];
            Settings order[kMaxCandidates];
            candidate_order(job.format, use_all, order);
            EstimateRange ranges[2];
            estimate_ranges(job.format, job.len, ranges);
            cd.image = next_image;
            next_image += (job.len + 255) / 256 * 256;
            work.push_back(BatchedTransform{order[cd.index], job.d_in, cd.image, job.len, /*estimate_only=*/true});
            for (size_t q = 0; q < js.plan.segs.size(); q++)
                if (js.plan.segs[q].cand == cd.index) segs[(size_t)js.first_seg + q].d_ptr = cd.image + ranges[js.plan.segs[q].range].offset;
        }
        cudaError_t qe = queue_transforms(work.data(), (int)work.size(), d_desc, stream);
        if (qe != cudaSuccess) return cuda_fail(qe);
        std::vector<uint64_t> matches(segs.size(), 0);
        if (!segs.empty()) {
            st = ltu_matches_device(segs.data(), (int)segs.size(), matches.data(), stream, ctx->d_scratch + est_offset,
                                    ctx->d_scratch_cap - est_offset);
            if (st != Status::kOk) return st;
        }

        // ---- winners (strict '<': the first candidate in test order wins ties) and their final transforms
        for (int j = j0; j < j1; j++) {
            AutoJob& job = jobs[j];
            const JobSegs& js = plans[(size_t)(j - j0)];
            Settings order[kMaxCandidates];
            const int k = candidate_order(job.format, use_all, order);
            // an empty payload: every estimate is 0, the first candidate in test order wins (as in the reference)
            job.best = order[0];
            for (int c = 0; c < kMaxCandidates; c++) job.sizes[c] = 0;
            if (job.len == 0) continue;
            size_t best_size = SIZE_MAX;
            for (int c = 0; c < k; c++) {
                size_t total = 0;
                for (int r = 0; r < js.nr; r++) {
                    const size_t q = (size_t)js.first_seg + (size_t)js.plan.seg_of[(size_t)c * js.nr + r];
                    total += ltu_estimate_from_matches(segs[q].len, matches[q]);
                }
                job.sizes[c] = total;
                if (total < best_size) best_size = total, job.best = order[c];
            }
        }
        work.clear();
        for (int j = j0; j < j1; j++)
            if (jobs[j].len) work.push_back(BatchedTransform{jobs[j].best, jobs[j].d_in, jobs[j].d_out, jobs[j].len});
        // (the descriptor area holds at least one entry per candidate, hence per job; stream order protects its reuse)
        qe = queue_transforms(work.data(), (int)work.size(), d_desc, stream);
        if (qe != cudaSuccess) return cuda_fail(qe);
        j0 = j1;
    }
    const cudaError_t e = cudaStreamSynchronize(stream);
    return e == cudaSuccess ? Status::kOk : cuda_fail(e);
}

}  // namespace dlt
