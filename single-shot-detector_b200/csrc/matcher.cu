// Target assignment: GT x anchor IoU, per-anchor argmax + thresholds, per-GT forced match, box-coder targets.
// Replaces detector/training_target_creation.py (match_boxes :48-130, create_targets :133-176,
// get_training_targets :5-45) and the per-image tf.map_fn of detector/ssd.py:165-199.
//
// The reference materialises the [G,A] IoU matrix plus a [G,A] int32 one-hot per image.  Here nothing of size
// G*A ever reaches memory: one thread owns one anchor, GT boxes are staged in shared memory, the per-anchor
// argmax is a register scan in GT order (strict '>' == tf.argmax's first maximum), and the per-GT argmax over
// anchors is a (value, lowest index) reduction: warp REDUX.MAX on the IoU bit pattern (IoU >= 0, so the uint
// order is the float order) -> shared-memory atomicMax -> one global atomicMax per (CTA, GT) on the packed key
//     key = iou_bits << 32 | (0xFFFFFFFF - anchor_index)      (ties -> LOWEST anchor index, as tf.argmax axis=1).
// A second, tiny kernel applies the forced matches, including the reference's row-id quirk (:117).
#include "common.cuh"

#define MATCH_THREADS 256
#define GT_CHUNK 512

// matches value from the thresholds: training_target_creation.py:92-100
__device__ __forceinline__ int threshold_match(int best_g, float best_v, float pos_thr, float neg_thr, bool same_thr) {
    if (best_v >= pos_thr) return best_g;
    if (same_thr) return -1;
    return (neg_thr > best_v) ? -1 : -2;
}


// Forced matches: training_target_creation.py:105-126.  For GT g: fid[g] = first anchor with the row maximum,
// ok[g] = (row maximum >= 0.1).  Anchor a is overridden iff some ok GT picked it; the value written is the
// LOWEST GT index among all GTs that picked a, ok or not (argmax over the unmasked one-hot, :117).
// Runs in one CTA per image: either force_match_kernel or the last match_kernel CTA of the image.
template <bool WRITE_TARGETS>
__device__ __forceinline__ void force_match_image(
    int b, int N, int* s_fid, unsigned char* s_ok, const float4* __restrict__ anchors, int A,
    const float4* __restrict__ gt_boxes, const int* __restrict__ gt_labels, int Gmax,
    const unsigned long long* gt_best, int* matches, float4* reg, int* cls, int* s_new_matched = nullptr) {
    for (int g = threadIdx.x; g < N; g += blockDim.x) {
        const unsigned long long key = __ldcg(&gt_best[(size_t)b * Gmax + g]);
        // key == 0: the whole IoU row is 0 -> argmax is anchor 0, value 0
        s_fid[g] = key ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : 0;
        s_ok[g] = __uint_as_float((unsigned)(key >> 32)) >= 0.1f;
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
            if (s_new_matched && __ldcg(&matches[o]) < 0) atomicAdd(s_new_matched, 1);   // a forced match of a so far unmatched anchor
            matches[o] = g;
            if (WRITE_TARGETS) {
                reg[o] = box_encode(gt_boxes[(size_t)b * Gmax + g], anchors[a]);
                cls[o] = gt_labels[(size_t)b * Gmax + g] + 1;
            }
        }
    }
}

template <bool WRITE_TARGETS>
__global__ void __launch_bounds__(MATCH_THREADS) match_kernel(
    const float4* __restrict__ anchors, int A, const float4* __restrict__ gt_boxes, const int* __restrict__ gt_labels,
    const int* __restrict__ num_boxes, int Gmax, float pos_thr, float neg_thr, int same_thr,
    unsigned long long* __restrict__ gt_best /*[B,Gmax], zeroed; may be NULL (no forced matching)*/,
    int* __restrict__ tickets /*[B], zeroed; non-NULL: the last CTA of an image applies the forced matches (Gmax <= GT_CHUNK)*/,
    int* __restrict__ img_count /*[B], zeroed; with out_count: matched anchors per image before forced matching*/,
    double* __restrict__ out_count /*zeroed or NULL: + number of matched anchors (ssd.py:89,121-122), added once per image*/,
    int* __restrict__ matches, float4* __restrict__ reg, int* __restrict__ cls) {
    __shared__ float4 s_box[GT_CHUNK];
    __shared__ float s_area[GT_CHUNK];
    __shared__ unsigned long long s_best[GT_CHUNK];

    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int N = num_boxes ? min(max(num_boxes[b], 0), Gmax) : Gmax;
    const float4* gtb = gt_boxes + (size_t)b * Gmax;
    const int nchunks = (A + MATCH_THREADS - 1) / MATCH_THREADS;

    // A CTA walks several 256-anchor chunks of its image (grid.x is sized to fill the GPU once): when all GT boxes fit
    // one staging chunk (the normal case) they are staged once and the per-GT maxima are accumulated in shared memory
    // over all the CTA's anchors, so that the barriers, the global atomics and the final fence are paid once per CTA.
    const bool single = N <= GT_CHUNK;
    auto stage = [&](int g0, int n) {
        for (int t = threadIdx.x; t < n; t += MATCH_THREADS) {
            const float4 gb = gtb[g0 + t];
            s_box[t] = gb;
            s_area[t] = box_area(gb);
            s_best[t] = 0ull;
        }
    };
    auto flush = [&](int g0, int n) {
        for (int t = threadIdx.x; t < n; t += MATCH_THREADS)
            if (s_best[t] != 0ull) atomicMax(&gt_best[(size_t)b * Gmax + g0 + t], s_best[t]);
    };
    if (single) {
        stage(0, N);
        __syncthreads();
    }

    int my_matched = 0;
    for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int a = chunk * MATCH_THREADS + threadIdx.x;
        const bool valid = a < A;
        float4 anc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) anc = anchors[a];
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
                    const float4 gb = s_box[t0 + lane];
                    near = !(gb.z <= wy0 || gb.x >= wy1 || gb.w <= wx0 || gb.y >= wx1);   // otherwise every lane's IoU is exactly 0
                }
                unsigned todo = __ballot_sync(0xffffffffu, near);
                while (todo) {
                    const int t = t0 + __ffs(todo) - 1;
                    todo &= todo - 1;
                    const float4 gb = s_box[t];
                    // iou(groundtruth_boxes, anchors): box_utils.py:14-27.  inter == 0 -> 0 / (union + eps) == 0 exactly.
                    const float inter = box_intersection(gb, anc);
                    float v = 0.0f;
                    if (valid && inter > 0.0f) {
                        const float uni = f_sub(f_add(s_area[t], area_a), inter);
                        v = fminf(fmaxf(f_div(inter, f_add(uni, SSDK_EPS)), 0.0f), 1.0f);
                    }
                    if (v > best_v) { best_v = v; best_g = g0 + t; }           // :90-91 (first max over GT)
                    if (gt_best) {                                              // :112,120 (first max over anchors)
                        const unsigned bits = __float_as_uint(v);
                        const unsigned wmax = __reduce_max_sync(0xffffffffu, bits);
                        if (wmax != 0u) {
                            const unsigned ball = __ballot_sync(0xffffffffu, bits == wmax);
                            if (lane == __ffs(ball) - 1)
                                atomicMax(&s_best[t], ((unsigned long long)wmax << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)a));
                        }
                    }
                }
            }
            if (!single && gt_best) {
                __syncthreads();
                flush(g0, n);
            }
        }
        if (valid) {
            const int m = (N > 0) ? threshold_match(best_g, best_v, pos_thr, neg_thr, same_thr != 0) : -1;   // :24-37
            const size_t o = (size_t)b * A + a;
            matches[o] = m;
            my_matched += (m >= 0);
            if (WRITE_TARGETS) {                                               // create_targets :133-176
                if (m >= 0) {
                    reg[o] = box_encode(gtb[m], anc);
                    cls[o] = gt_labels[(size_t)b * Gmax + m] + 1;
                } else {
                    reg[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                    cls[o] = 0;
                }
            }
        }
    }
    if (single && gt_best && N > 0) {
        __syncthreads();
        flush(0, N);
    }
    if (tickets && N > 0) {
        // Fused forced matching: every CTA publishes its writes and takes a ticket; the CTA that draws the last ticket
        // of the image sees all threshold results and all per-GT maxima, and overrides the forced anchors.
        __shared__ int s_last, s_cnt;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        if (out_count) {                                              // this CTA's matched anchors -> the image's counter
            my_matched = __reduce_add_sync(0xffffffffu, my_matched);
            if (lane == 0 && my_matched) atomicAdd(&s_cnt, my_matched);
            __syncthreads();
            if (threadIdx.x == 0 && s_cnt) atomicAdd(&img_count[b], s_cnt);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) { s_last = (atomicAdd(&tickets[b], 1) == (int)gridDim.x - 1); s_cnt = 0; }
        __syncthreads();
        if (s_last) {
            __threadfence();
            force_match_image<WRITE_TARGETS>(b, N, (int*)s_area, (unsigned char*)s_box, anchors, A, gt_boxes, gt_labels, Gmax,
                                             gt_best, matches, reg, cls, out_count ? &s_cnt : nullptr);
            if (out_count) {
                __syncthreads();
                if (threadIdx.x == 0) {
                    const int total = __ldcg(&img_count[b]) + s_cnt;
                    if (total) atomicAdd(out_count, (double)total);      // integers: exact and order independent
                }
            }
        }
    }
}

// Stand-alone forced matching for images with more than GT_CHUNK boxes (one CTA per image).
template <bool WRITE_TARGETS>
__global__ void __launch_bounds__(256) force_match_kernel(
    const float4* __restrict__ anchors, int A, const float4* __restrict__ gt_boxes, const int* __restrict__ gt_labels,
    const int* __restrict__ num_boxes, int Gmax, const unsigned long long* __restrict__ gt_best,
    int* __restrict__ matches, float4* __restrict__ reg, int* __restrict__ cls) {
    extern __shared__ int s_dyn[];
    int* s_fid = s_dyn;                                  // [Gmax]
    unsigned char* s_ok = (unsigned char*)(s_dyn + Gmax);  // [Gmax]
    const int b = blockIdx.x;
    const int N = num_boxes ? min(max(num_boxes[b], 0), Gmax) : Gmax;
    force_match_image<WRITE_TARGETS>(b, N, s_fid, s_ok, anchors, A, gt_boxes, gt_labels, Gmax, gt_best, matches, reg, cls);
}

// create_targets alone (:133-176), for callers that bring their own matches.
__global__ void __launch_bounds__(256) create_targets_kernel(
    const float4* __restrict__ anchors, int A, const float4* __restrict__ gt_boxes, const int* __restrict__ gt_labels,
    int Gmax, const int* __restrict__ matches, float4* __restrict__ reg, int* __restrict__ cls) {
    const int b = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    const size_t o = (size_t)b * A + a;
    const int m = matches[o];
    if (m >= 0 && m < Gmax) {
        reg[o] = box_encode(gt_boxes[(size_t)b * Gmax + m], anchors[a]);
        cls[o] = gt_labels[(size_t)b * Gmax + m] + 1;
    } else {
        reg[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        cls[o] = 0;
    }
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

int ssdk_count_impl(ssdk_ctx* ctx, const int32_t* matches, int64_t n, double* out_count);

int ssdk_match_impl(ssdk_ctx* ctx, const float* anchors, int64_t A, const float* gt_boxes, const int32_t* gt_labels,
                    const int32_t* num_boxes, int B, int Gmax, double pos_thr, double neg_thr, int force,
                    float* out_reg, int32_t* out_cls, int32_t* out_matches, double* out_count = nullptr) {
    SSDK_TRY(ssdk_ctx_enter(ctx));
    SSDK_REQUIRE(pos_thr >= neg_thr, SSDK_ERR_ARG,
                 "positives_threshold (%g) must be >= negatives_threshold (%g)", pos_thr, neg_thr);  // :86
    SSDK_REQUIRE(B >= 0 && A >= 0 && Gmax >= 0, SSDK_ERR_ARG, "match: negative size");
    SSDK_REQUIRE(A < (1ll << 31) && B <= 65535, SSDK_ERR_SHAPE, "match: A must be < 2^31 and B <= 65535");
    SSDK_REQUIRE(Gmax <= 4096, SSDK_ERR_SHAPE, "match: at most 4096 ground-truth boxes per image (got %d)", Gmax);
    if (out_count) SSDK_CHECK_CUDA(cudaMemsetAsync(out_count, 0, sizeof(double), ctx->stream));
    if (B == 0 || A == 0) return SSDK_OK;
    SSDK_REQUIRE(anchors && out_matches && (Gmax == 0 || gt_boxes), SSDK_ERR_ARG, "match: null pointer");
    const bool targets = out_reg != nullptr || out_cls != nullptr;
    if (targets) SSDK_REQUIRE(out_reg && out_cls && (Gmax == 0 || gt_labels), SSDK_ERR_ARG, "match: targets need reg, cls and labels");
    SSDK_REQUIRE(aligned16(anchors) && aligned16(gt_boxes) && aligned16(out_reg), SSDK_ERR_SHAPE,
                 "match: box arrays must be 16-byte aligned");
    unsigned long long* best = nullptr;
    int* tickets = nullptr;
    int* img_count = nullptr;
    if (force && Gmax > 0) {
        const size_t bytes = (size_t)B * Gmax * sizeof(unsigned long long) + 2 * (size_t)B * sizeof(int);   // keys + tickets + counts
        SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_gtbest, bytes));
        best = (unsigned long long*)ctx->ws_gtbest.p;
        if (Gmax <= GT_CHUNK) {
            tickets = (int*)(best + (size_t)B * Gmax);
            img_count = tickets + B;
        }
        SSDK_CHECK_CUDA(cudaMemsetAsync(best, 0, bytes, ctx->stream));
    }
    // the matched count is folded into the kernel when the forced matches are (one CTA per image sees the final state);
    // otherwise a separate pass over `matches` counts afterwards
    double* fused_count = tickets ? out_count : nullptr;
    // about six resident CTAs per SM in total; each CTA walks ceil(nchunks / grid.x) chunks of its image
    const int nchunks = ceil_div_i(A, MATCH_THREADS);
    int gx = (ctx->num_sms * 6 + B - 1) / B;
    if (gx > nchunks) gx = nchunks;
    if (gx < 1) gx = 1;
    const dim3 grid(gx, B);
    const int same = (pos_thr == neg_thr) ? 1 : 0;   // compared as Python floats in the reference (:94)
    SSDK_KERNEL(ctx, SSDK_K_MATCH,
        if (targets)
            match_kernel<true><<<grid, MATCH_THREADS, 0, ctx->stream>>>(
                (const float4*)anchors, (int)A, (const float4*)gt_boxes, gt_labels, num_boxes, Gmax, (float)pos_thr,
                (float)neg_thr, same, best, tickets, img_count, fused_count, out_matches, (float4*)out_reg, out_cls);
        else
            match_kernel<false><<<grid, MATCH_THREADS, 0, ctx->stream>>>(
                (const float4*)anchors, (int)A, (const float4*)gt_boxes, gt_labels, num_boxes, Gmax, (float)pos_thr,
                (float)neg_thr, same, best, tickets, img_count, fused_count, out_matches, nullptr, nullptr));
    if (best && !tickets) {
        const size_t smem = (size_t)Gmax * 5 + 16;
        SSDK_KERNEL(ctx, SSDK_K_FORCE_MATCH,
            if (targets)
                force_match_kernel<true><<<B, 256, smem, ctx->stream>>>((const float4*)anchors, (int)A, (const float4*)gt_boxes,
                                                                       gt_labels, num_boxes, Gmax, best, out_matches,
                                                                       (float4*)out_reg, out_cls);
            else
                force_match_kernel<false><<<B, 256, smem, ctx->stream>>>((const float4*)anchors, (int)A, (const float4*)gt_boxes,
                                                                        gt_labels, num_boxes, Gmax, best, out_matches, nullptr,
                                                                        nullptr));
    }
    if (out_count && !fused_count) SSDK_TRY(ssdk_count_impl(ctx, out_matches, (int64_t)B * A, out_count));
    return SSDK_OK;
}

extern "C" {

int ssdk_training_targets_count(ssdk_ctx* ctx, const float* anchors, int64_t A, const float* gt_boxes, const int32_t* gt_labels,
                                const int32_t* num_boxes, int B, int Gmax, double pos_thr, double neg_thr, float* out_reg,
                                int32_t* out_cls, int32_t* out_matches, double* out_count) {
    SSDK_REQUIRE(out_reg && out_cls && out_count, SSDK_ERR_ARG, "ssdk_training_targets_count: out_reg/out_cls/out_count are required");
    return ssdk_match_impl(ctx, anchors, A, gt_boxes, gt_labels, num_boxes, B, Gmax, pos_thr, neg_thr, 1, out_reg, out_cls,
                           out_matches, out_count);
}

int ssdk_match_boxes(ssdk_ctx* ctx, const float* anchors, int64_t A, const float* gt_boxes, const int32_t* num_boxes,
                     int B, int Gmax, double pos_thr, double neg_thr, int force, int32_t* out_matches) {
    return ssdk_match_impl(ctx, anchors, A, gt_boxes, nullptr, num_boxes, B, Gmax, pos_thr, neg_thr, force, nullptr,
                           nullptr, out_matches);
}

int ssdk_training_targets(ssdk_ctx* ctx, const float* anchors, int64_t A, const float* gt_boxes, const int32_t* gt_labels,
                          const int32_t* num_boxes, int B, int Gmax, double pos_thr, double neg_thr, float* out_reg,
                          int32_t* out_cls, int32_t* out_matches) {
    SSDK_REQUIRE(out_reg && out_cls, SSDK_ERR_ARG, "ssdk_training_targets: out_reg/out_cls are required");
    return ssdk_match_impl(ctx, anchors, A, gt_boxes, gt_labels, num_boxes, B, Gmax, pos_thr, neg_thr, 1, out_reg, out_cls,
                           out_matches);
}

int ssdk_create_targets(ssdk_ctx* ctx, const float* anchors, int64_t A, const float* gt_boxes, const int32_t* gt_labels,
                        int B, int Gmax, const int32_t* matches, float* out_reg, int32_t* out_cls) {
    SSDK_TRY(ssdk_ctx_enter(ctx));
    SSDK_REQUIRE(B >= 0 && A >= 0 && Gmax >= 0 && A < (1ll << 31) && B <= 65535, SSDK_ERR_ARG, "create_targets: bad sizes");
    if (B == 0 || A == 0) return SSDK_OK;
    SSDK_REQUIRE(anchors && matches && out_reg && out_cls && (Gmax == 0 || (gt_boxes && gt_labels)), SSDK_ERR_ARG,
                 "create_targets: null pointer");
    SSDK_REQUIRE(aligned16(anchors) && aligned16(gt_boxes) && aligned16(out_reg), SSDK_ERR_SHAPE,
                 "create_targets: box arrays must be 16-byte aligned");
    create_targets_kernel<<<dim3(ceil_div_i(A, 256), B), 256, 0, ctx->stream>>>(
        (const float4*)anchors, (int)A, (const float4*)gt_boxes, gt_labels, Gmax, matches, (float4*)out_reg, out_cls);
    SSDK_CHECK_LAUNCH(ctx);
    return SSDK_OK;
}

}  // extern "C"
