// Device side of the target assignment, shared by match_kernel (matcher.cu) and the fused training-step kernel
// (train_step.cu).  Replaces detector/training_target_creation.py (match_boxes :48-130, create_targets :133-176,
// get_training_targets :5-45) and the per-image tf.map_fn of detector/ssd.py:165-199.
//
// The reference materialises the [G,A] IoU matrix plus a [G,A] int32 one-hot per image.  Here nothing of size G*A ever
// reaches memory: one thread owns one anchor, GT boxes are staged in shared memory, the per-anchor argmax is a register scan
// in GT order (strict '>' == tf.argmax's first maximum), and the per-GT argmax over anchors is a (value, lowest index)
// reduction: warp REDUX.MAX on the IoU bit pattern (IoU >= 0, so the uint order is the float order) -> shared-memory
// atomicMax -> one global atomicMax per (work item, GT) on the packed key
//     key = iou_bits << 32 | (0xFFFFFFFF - anchor_index)      (ties -> LOWEST anchor index, as tf.argmax axis=1).
// The forced matches (:105-126), including the reference's row-id quirk (:117), are applied by whichever CTA finishes the
// image last (ticket counter).
//
// A `Hook` observes the result: hook.anchors(...) is called warp-collectively for every 32 anchors with their final
// threshold match (before forced matching), hook.forced(...) by the single thread that overrides an anchor (for ground-truth
// box g; hook.forced_init(g) was called for every g before), hook.image_done(...) by all threads once an image is complete.  The fused
// training step uses it to add the matched / ignored anchors' loss contributions in the same pass; match_kernel passes a
// no-op.
#pragma once
#include "common.cuh"

#define MATCH_THREADS 256
#ifndef GT_CHUNK
#define GT_CHUNK 512
#endif

struct MatchSmem {
    float4 box[GT_CHUNK];
    float area[GT_CHUNK];
    unsigned long long best[GT_CHUNK];
    int last, cnt;
};

struct MatchArgs {
    const float4* anchors; int A;
    const float4* gt_boxes; const int* gt_labels; const int* num_boxes; int Gmax;
    float pos_thr, neg_thr; int same_thr;
    unsigned long long* gt_best;   // [B,Gmax], zero before the launch; may be NULL (no forced matching)
    int* tickets;                  // [B], zero before the launch; non-NULL: the last work item of an image applies the forced matches (Gmax <= GT_CHUNK)
    int* img_count;                // [B], zero before the launch; with out_count: matched anchors per image before forced matching
    double* out_count;             // zeroed or NULL: + number of matched anchors (ssd.py:89,121-122), added once per image
    int* matches; float4* reg; int* cls;
    int self_clean;                // 1: the CTA that finishes an image re-zeroes its gt_best / tickets / img_count entries
};

struct MatchNoHook {
    __device__ __forceinline__ void anchors(int, int, bool, int, const float4&, const float4&, int) {}
    __device__ __forceinline__ void forced_init(int) {}
    __device__ __forceinline__ void forced(int, int, int, int, const float4&) {}
    __device__ __forceinline__ void image_done(int, int) {}
};

// matches value from the thresholds: training_target_creation.py:92-100
__device__ __forceinline__ int threshold_match(int best_g, float best_v, float pos_thr, float neg_thr, bool same_thr) {
    if (best_v >= pos_thr) return best_g;
    if (same_thr) return -1;
    return (neg_thr > best_v) ? -1 : -2;
}

// Forced matches: training_target_creation.py:105-126.  For GT g: fid[g] = first anchor with the row maximum,
// ok[g] = (row maximum >= 0.1).  Anchor a is overridden iff some ok GT picked it; the value written is the
// LOWEST GT index among all GTs that picked a, ok or not (argmax over the unmasked one-hot, :117).
// Runs in one CTA per image: either force_match_kernel or the CTA that finishes the image last.
template <bool WRITE_TARGETS, class Hook>
__device__ __forceinline__ void force_match_image(
    int b, int N, int* s_fid, unsigned char* s_ok, const float4* __restrict__ anchors, int A,
    const float4* __restrict__ gt_boxes, const int* __restrict__ gt_labels, int Gmax,
    const unsigned long long* gt_best, int* matches, float4* reg, int* cls, int* s_new_matched, Hook& hook) {
    for (int g = threadIdx.x; g < N; g += blockDim.x) {
        const unsigned long long key = __ldcg(&gt_best[(size_t)b * Gmax + g]);
        // key == 0: the whole IoU row is 0 -> argmax is anchor 0, value 0
        s_fid[g] = key ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : 0;
        s_ok[g] = __uint_as_float((unsigned)(key >> 32)) >= 0.1f;
        hook.forced_init(g);
    }
    __syncthreads();
    for (int g = threadIdx.x; g < N; g += blockDim.x) {
        const int a = s_fid[g];
        bool first = true, any_ok = false;
        for (int h = 0; h < N; ++h) {
            if (s_fid[h] == a) {
                if (h < g) first = false;
                any_ok |= (s_ok[h] != 0);
            }
        }
        if (first && any_ok) {
            const size_t o = (size_t)b * A + a;
            const int m_old = __ldcg(&matches[o]);
            if (s_new_matched && m_old < 0) atomicAdd(s_new_matched, 1);   // a forced match of a so far unmatched anchor
            matches[o] = g;
            const float4 anc = anchors[a];
            hook.forced(b, a, m_old, g, anc);
            if (WRITE_TARGETS) {
                reg[o] = box_encode(gt_boxes[(size_t)b * Gmax + g], anc);
                cls[o] = gt_labels[(size_t)b * Gmax + g] + 1;
            }
        }
    }
}

// One work item: image b, the 256-anchor chunks cx, cx + gx, ... of that image.  All MATCH_THREADS threads of the CTA take
// part (barriers inside).  When all GT boxes fit one staging chunk (the normal case) they are staged once and the per-GT
// maxima are accumulated in shared memory over all the item's anchors, so that the barriers, the global atomics and the final
// fence are paid once per item instead of once per 256 anchors.
template <bool WRITE_TARGETS, class Hook>
__device__ __forceinline__ void match_work_item(const MatchArgs& M, MatchSmem& sm, int b, int cx, int gx, Hook& hook) {
    const int lane = threadIdx.x & 31;
    const int A = M.A, Gmax = M.Gmax;
    const int N = M.num_boxes ? min(max(M.num_boxes[b], 0), Gmax) : Gmax;
    const float4* gtb = M.gt_boxes + (size_t)b * Gmax;
    const int nchunks = (A + MATCH_THREADS - 1) / MATCH_THREADS;
    unsigned long long* gt_best = M.gt_best;

    const bool single = N <= GT_CHUNK;
    auto stage = [&](int g0, int n) {
        for (int t = threadIdx.x; t < n; t += MATCH_THREADS) {
            const float4 gb = gtb[g0 + t];
            sm.box[t] = gb;
            sm.area[t] = box_area(gb);
            sm.best[t] = 0ull;
        }
    };
    auto flush = [&](int g0, int n) {
        for (int t = threadIdx.x; t < n; t += MATCH_THREADS)
            if (sm.best[t] != 0ull) atomicMax(&gt_best[(size_t)b * Gmax + g0 + t], sm.best[t]);
    };
    __syncthreads();                                                  // the previous item of this CTA is done with the shared memory
    if (single) {
        stage(0, N);
        __syncthreads();
    }

    int my_matched = 0;
    for (int chunk = cx; chunk < nchunks; chunk += gx) {
        const int a = chunk * MATCH_THREADS + threadIdx.x;
        const bool valid = a < A;
        float4 anc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) anc = M.anchors[a];
        const float area_a = box_area(anc);

        // Bounding box of the warp's 32 consecutive anchors (neighbouring cells of one FPN level): a GT box that does
        // not overlap it has intersection 0 -- hence IoU exactly 0 -- with every lane, and is skipped warp-uniformly.
        float wy0 = valid ? anc.x : INFINITY, wx0 = valid ? anc.y : INFINITY;
        float wy1 = valid ? anc.z : -INFINITY, wx1 = valid ? anc.w : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wy0 = fminf(wy0, __shfl_xor_sync(0xffffffffu, wy0, o));
            wx0 = fminf(wx0, __shfl_xor_sync(0xffffffffu, wx0, o));
            wy1 = fmaxf(wy1, __shfl_xor_sync(0xffffffffu, wy1, o));
            wx1 = fmaxf(wx1, __shfl_xor_sync(0xffffffffu, wx1, o));
        }

        float best_v = 0.0f;   // IoU is clipped to [0,1]: starting from (0, index 0) with strict '>' is tf.argmax
        int best_g = 0;

        for (int g0 = 0; g0 < N; g0 += GT_CHUNK) {
            const int n = min(GT_CHUNK, N - g0);
            if (!single) {
                __syncthreads();
                stage(g0, n);
                __syncthreads();
            }
            // 32 GT boxes at a time: lane j tests box t0+j against the warp's bounding box, the ballot is the set of
            // boxes that can have a non-zero IoU with some lane; only those are visited (in index order, as tf.argmax needs)
            for (int t0 = 0; t0 < n; t0 += 32) {
                bool near = false;
                if (t0 + lane < n) {
                    const float4 gb = sm.box[t0 + lane];
                    near = !(gb.z <= wy0 || gb.x >= wy1 || gb.w <= wx0 || gb.y >= wx1);   // otherwise every lane's IoU is exactly 0
                }
                unsigned todo = __ballot_sync(0xffffffffu, near);
                while (todo) {
                    const int t = t0 + __ffs(todo) - 1;
                    todo &= todo - 1;
                    const float4 gb = sm.box[t];
                    // iou(groundtruth_boxes, anchors): box_utils.py:14-27.  inter == 0 -> 0 / (union + eps) == 0 exactly.
                    const float inter = box_intersection(gb, anc);
                    float v = 0.0f;
                    if (valid && inter > 0.0f) {
                        const float uni = f_sub(f_add(sm.area[t], area_a), inter);
                        v = fminf(fmaxf(f_div(inter, f_add(uni, SSDK_EPS)), 0.0f), 1.0f);
                    }
                    if (v > best_v) { best_v = v; best_g = g0 + t; }           // :90-91 (first max over GT)
                    if (gt_best) {                                              // :112,120 (first max over anchors)
                        const unsigned bits = __float_as_uint(v);
                        const unsigned wmax = __reduce_max_sync(0xffffffffu, bits);
                        if (wmax != 0u) {
                            const unsigned ball = __ballot_sync(0xffffffffu, bits == wmax);
                            if (lane == __ffs(ball) - 1)
                                atomicMax(&sm.best[t], ((unsigned long long)wmax << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)a));
                        }
                    }
                }
            }
            if (!single && gt_best) {
                __syncthreads();
                flush(g0, n);
            }
        }
        int m = -1;
        float4 target = make_float4(0.f, 0.f, 0.f, 0.f);
        int label1 = 0;
        if (valid) {
            m = (N > 0) ? threshold_match(best_g, best_v, M.pos_thr, M.neg_thr, M.same_thr != 0) : -1;   // :24-37
            const size_t o = (size_t)b * A + a;
            M.matches[o] = m;
            my_matched += (m >= 0);
            if (m >= 0) {                                                      // create_targets :133-176
                target = box_encode(gtb[m], anc);
                label1 = M.gt_labels ? M.gt_labels[(size_t)b * Gmax + m] + 1 : 0;
            }
            if (WRITE_TARGETS) {
                M.reg[o] = target;
                M.cls[o] = label1;
            }
        }
        hook.anchors(b, a, valid, m, anc, target, label1);
    }
    if (single && gt_best && N > 0) {
        __syncthreads();
        flush(0, N);
    }
    if (M.tickets && N > 0) {
        // Fused forced matching: every item publishes its writes and takes a ticket; the item that draws the last ticket
        // of the image sees all threshold results and all per-GT maxima, and overrides the forced anchors.
        if (threadIdx.x == 0) sm.cnt = 0;
        __syncthreads();
        if (M.out_count) {                                            // this item's matched anchors -> the image's counter
            my_matched = __reduce_add_sync(0xffffffffu, my_matched);
            if (lane == 0 && my_matched) atomicAdd(&sm.cnt, my_matched);
            __syncthreads();
            if (threadIdx.x == 0 && sm.cnt) atomicAdd(&M.img_count[b], sm.cnt);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) { sm.last = (atomicAdd(&M.tickets[b], 1) == gx - 1); sm.cnt = 0; }
        __syncthreads();
        if (sm.last) {
            __threadfence();
            force_match_image<WRITE_TARGETS>(b, N, (int*)sm.area, (unsigned char*)sm.box, M.anchors, A, M.gt_boxes, M.gt_labels, Gmax,
                                             gt_best, M.matches, M.reg, M.cls, M.out_count ? &sm.cnt : nullptr, hook);
            __syncthreads();
            if (M.out_count && threadIdx.x == 0) {
                const int total = __ldcg(&M.img_count[b]) + sm.cnt;
                if (total) atomicAdd(M.out_count, (double)total);        // integers: exact and order independent
            }
            if (M.self_clean) {                                           // leave the workspace zeroed for the next launch
                for (int g = threadIdx.x; g < N; g += MATCH_THREADS) gt_best[(size_t)b * Gmax + g] = 0ull;
                if (threadIdx.x == 0) { M.tickets[b] = 0; if (M.img_count) M.img_count[b] = 0; }
            }
            hook.image_done(b, N);
        }
    }
}
