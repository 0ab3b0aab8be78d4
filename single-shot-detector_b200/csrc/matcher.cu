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
// The device code lives in matcher.cuh (shared with the fused training-step kernel of train_step.cu).
#include "matcher.cuh"

#ifndef MATCH_MIN_CTAS
#define MATCH_MIN_CTAS 6               // 40 registers
#endif
#ifndef MATCH_GRID_PER_SM
#define MATCH_GRID_PER_SM 6
#endif
template <bool WRITE_TARGETS>
__global__ void __launch_bounds__(MATCH_THREADS, MATCH_MIN_CTAS) match_kernel(const MatchArgs M) {
    __shared__ MatchSmem sm;
    MatchNoHook hook;
    match_work_item<WRITE_TARGETS>(M, sm, blockIdx.y, blockIdx.x, gridDim.x, hook);
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
    MatchNoHook hook;
    force_match_image<WRITE_TARGETS>(b, N, s_fid, s_ok, anchors, A, gt_boxes, gt_labels, Gmax, gt_best, matches, reg, cls, nullptr, hook);
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
    SSDK_ENTER(ctx);
    SsdkWsGuard ws_guard(ctx, SSDK_WS_MATCH);
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
    int gx = (ctx->num_sms * MATCH_GRID_PER_SM + B - 1) / B;
    if (gx > nchunks) gx = nchunks;
    if (gx < 1) gx = 1;
    const dim3 grid(gx, B);
    const int same = (pos_thr == neg_thr) ? 1 : 0;   // compared as Python floats in the reference (:94)
    MatchArgs M;
    M.anchors = (const float4*)anchors; M.A = (int)A;
    M.gt_boxes = (const float4*)gt_boxes; M.gt_labels = gt_labels; M.num_boxes = num_boxes; M.Gmax = Gmax;
    M.pos_thr = (float)pos_thr; M.neg_thr = (float)neg_thr; M.same_thr = same;
    M.gt_best = best; M.tickets = tickets; M.img_count = img_count; M.out_count = fused_count;
    M.matches = out_matches; M.reg = (float4*)out_reg; M.cls = out_cls; M.self_clean = 0;
    SSDK_KERNEL(ctx, SSDK_K_MATCH,
        if (targets) match_kernel<true><<<grid, MATCH_THREADS, 0, ctx->stream>>>(M);
        else match_kernel<false><<<grid, MATCH_THREADS, 0, ctx->stream>>>(M));
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
    SSDK_ENTER(ctx);
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
