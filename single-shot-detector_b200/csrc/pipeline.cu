// Composite entry points: the whole training-side hot path in one call, and the *_host variants that take host
// buffers (H2D copy, kernels, D2H copy inside the call).  See include/ssdk.h.
#include "common.cuh"

int ssdk_match_impl(ssdk_ctx* ctx, const float* anchors, int64_t A, const float* gt_boxes, const int32_t* gt_labels,
                    const int32_t* num_boxes, int B, int Gmax, double pos_thr, double neg_thr, int force,
                    float* out_reg, int32_t* out_cls, int32_t* out_matches, double* out_count = nullptr);

struct HeadGeom;
HeadGeom ssdk_flat_geom(const float* logits, const float* codes, int64_t A, int C);
int ssdk_train_step_impl(ssdk_ctx* ctx, const HeadGeom& G, const float* anchors, const float* gt_boxes, const int32_t* gt_labels,
                         const int32_t* num_boxes, int B, int64_t A, int C, int Gmax, double pos_thr, double neg_thr, double gamma,
                         double alpha, int flags, double* out_sums, float* out_losses, float* out_reg, int32_t* out_cls,
                         int32_t* out_matches);

// the anchor-major tensors qualify for the flat pass (one channels_last "level", 16-byte aligned, sizes within the geometry's ints)
static bool flat_path_ok(const float* logits, const float* codes, int64_t A, int C) {
    return C > 0 && A < (1ll << 31) && logits && codes && (((uintptr_t)logits | (uintptr_t)codes) & 15) == 0 &&
           (long long)A * (C > 4 ? C : 4) < (1ll << 31);
}

template <typename T>
static int stage_h2d(ssdk_ctx* ctx, int slot, const T* host, size_t count, T** dev) {
    *dev = nullptr;
    if (count == 0) return SSDK_OK;
    SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_stage[slot], count * sizeof(T)));
    *dev = (T*)ctx->ws_stage[slot].p;
    if (host) SSDK_CHECK_CUDA(cudaMemcpyAsync(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return SSDK_OK;
}

extern "C" {

int ssdk_ssd_targets_and_loss(ssdk_ctx* ctx, const float* anchors, const float* logits, const float* codes,
                              const float* gt_boxes, const int32_t* gt_labels, const int32_t* num_boxes, int B, int64_t A,
                              int C, int Gmax, double pos_thr, double neg_thr, double gamma, double alpha, double* out_sums,
                              float* out_reg, int32_t* out_cls, int32_t* out_matches, float* out_cls_losses,
                              float* out_loc_losses) {
    SSDK_ENTER(ctx);
    SsdkWsGuard ws_guard(ctx, SSDK_WS_TRAIN);
    SSDK_REQUIRE(B >= 0 && A >= 0, SSDK_ERR_ARG, "ssdk_ssd_targets_and_loss: bad sizes");
    if (!out_cls_losses && !out_loc_losses && flat_path_ok(logits, codes, A, C)) {
        // no per-anchor outputs wanted: the fused training step (csrc/train_step.cu) -- matching, the flat pass over the logits and
        // the matched-anchor corrections in one launch.  The anchor-major tensors are just a one-level channels_last head.
        const HeadGeom G = ssdk_flat_geom(logits, codes, A, C);
        return ssdk_train_step_impl(ctx, G, anchors, gt_boxes, gt_labels, num_boxes, B, A, C, Gmax, pos_thr, neg_thr, gamma, alpha, 0,
                                    out_sums, nullptr, out_reg, out_cls, out_matches);
    }
    const size_t NA = (size_t)B * (size_t)A;
    if (!out_reg) { SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_reg, NA * 16 + 16)); out_reg = (float*)ctx->ws_reg.p; }
    if (!out_cls) { SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_cls, NA * 4 + 16)); out_cls = (int32_t*)ctx->ws_cls.p; }
    if (!out_matches) { SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_matches, NA * 4 + 16)); out_matches = (int32_t*)ctx->ws_matches.p; }
    SSDK_TRY(ssdk_match_impl(ctx, anchors, A, gt_boxes, gt_labels, num_boxes, B, Gmax, pos_thr, neg_thr, 1, out_reg, out_cls,
                             out_matches));                                                   // ssd.py:84
    return ssdk_ssd_loss(ctx, logits, codes, out_reg, out_cls, out_matches, B, A, C, gamma, alpha, out_sums, out_cls_losses,
                         out_loc_losses);                                                     // ssd.py:89-133
}

int ssdk_ssd_loss_step(ssdk_ctx* ctx, const float* anchors, const float* logits, const float* codes, const float* gt_boxes,
                       const int32_t* gt_labels, const int32_t* num_boxes, int B, int64_t A, int C, int Gmax, double pos_thr,
                       double neg_thr, double gamma, double alpha, int flags, double* out_sums, float* out_losses, float* out_reg,
                       int32_t* out_cls, int32_t* out_matches) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(B >= 0 && A >= 0 && C > 0, SSDK_ERR_ARG, "ssdk_ssd_loss_step: bad sizes");
    SSDK_REQUIRE(out_sums && out_losses, SSDK_ERR_ARG, "ssdk_ssd_loss_step: out_sums / out_losses are required");
    if (flat_path_ok(logits, codes, A, C)) {
        const HeadGeom G = ssdk_flat_geom(logits, codes, A, C);
        return ssdk_train_step_impl(ctx, G, anchors, gt_boxes, gt_labels, num_boxes, B, A, C, Gmax, pos_thr, neg_thr, gamma, alpha, flags,
                                    out_sums, out_losses, out_reg, out_cls, out_matches);
    }
    // unaligned or oversized tensors: the row-tiled kernel, then the exchange / finalisation as separate launches
    SSDK_TRY(ssdk_ssd_targets_and_loss(ctx, anchors, logits, codes, gt_boxes, gt_labels, num_boxes, B, A, C, Gmax, pos_thr, neg_thr, gamma,
                                       alpha, out_sums, out_reg, out_cls, out_matches, nullptr, nullptr));
    if (flags & SSDK_STEP_ALL_REDUCE) return ssdk_comm_loss_finalize(ctx, out_sums, out_losses);
    return ssdk_loss_finalize(ctx, out_sums, out_losses);
}

int ssdk_ssd_targets_and_loss_host(ssdk_ctx* ctx, const float* anchors, const float* logits, const float* codes,
                                   const float* gt_boxes, const int32_t* gt_labels, const int32_t* num_boxes, int B, int64_t A,
                                   int C, int Gmax, double pos_thr, double neg_thr, double gamma, double alpha,
                                   double* out_sums, float* out_losses) {
    SSDK_ENTER(ctx);
    SsdkWsGuard ws_guard(ctx, SSDK_WS_STAGE);
    SSDK_REQUIRE(B >= 0 && A >= 0 && C > 0 && Gmax >= 0, SSDK_ERR_ARG, "ssdk_ssd_targets_and_loss_host: bad sizes");
    SSDK_REQUIRE(out_sums || out_losses, SSDK_ERR_ARG, "ssdk_ssd_targets_and_loss_host: no output requested");
    const size_t NA = (size_t)B * (size_t)A;
    float *d_anchors, *d_logits, *d_codes, *d_gt;
    int32_t *d_labels, *d_num;
    double* d_out;
    SSDK_TRY(stage_h2d(ctx, 0, anchors, (size_t)A * 4, &d_anchors));
    SSDK_TRY(stage_h2d(ctx, 1, logits, NA * C, &d_logits));
    SSDK_TRY(stage_h2d(ctx, 2, codes, NA * 4, &d_codes));
    SSDK_TRY(stage_h2d(ctx, 3, gt_boxes, (size_t)B * Gmax * 4, &d_gt));
    SSDK_TRY(stage_h2d(ctx, 4, gt_labels, (size_t)B * Gmax, &d_labels));
    SSDK_TRY(stage_h2d(ctx, 5, num_boxes, (size_t)(num_boxes ? B : 0), &d_num));
    SSDK_TRY(stage_h2d(ctx, 6, (const double*)nullptr, 4, &d_out));   // double[3] sums + float[2] losses
    SSDK_TRY(ssdk_ssd_loss_step(ctx, d_anchors, d_logits, d_codes, d_gt, d_labels, d_num, B, A, C, Gmax, pos_thr, neg_thr, gamma, alpha, 0,
                                d_out, (float*)(d_out + 3), nullptr, nullptr, nullptr));
    double h[4];
    SSDK_CHECK_CUDA(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    SSDK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_sums) { out_sums[0] = h[0]; out_sums[1] = h[1]; out_sums[2] = h[2]; }
    if (out_losses) { const float* f = (const float*)&h[3]; out_losses[0] = f[0]; out_losses[1] = f[1]; }
    return SSDK_OK;
}

int ssdk_postprocess_host(ssdk_ctx* ctx, const float* codes, const float* anchors, const float* scores, int flags, int B,
                          int64_t A, int C, double score_threshold, double iou_threshold, int K, float* out_boxes,
                          float* out_scores, int32_t* out_classes, int32_t* out_num) {
    SSDK_ENTER(ctx);
    SsdkWsGuard ws_guard(ctx, SSDK_WS_STAGE);
    SSDK_REQUIRE(B >= 0 && A >= 0 && C > 0 && K > 0, SSDK_ERR_ARG, "ssdk_postprocess_host: bad sizes");
    SSDK_REQUIRE(out_boxes && out_scores && out_classes && out_num, SSDK_ERR_ARG, "ssdk_postprocess_host: null output");
    const size_t NA = (size_t)B * (size_t)A, M = (size_t)B * C * K;
    float *d_codes, *d_anchors, *d_scores, *d_ob, *d_os;
    int32_t *d_oc, *d_on;
    SSDK_TRY(stage_h2d(ctx, 0, (flags & SSDK_BOXES_DECODED) ? (const float*)nullptr : anchors,
                       (flags & SSDK_BOXES_DECODED) ? 0 : (size_t)A * 4, &d_anchors));
    SSDK_TRY(stage_h2d(ctx, 1, scores, NA * C, &d_scores));
    SSDK_TRY(stage_h2d(ctx, 2, codes, NA * 4, &d_codes));
    SSDK_TRY(stage_h2d(ctx, 3, (const float*)nullptr, M * 4, &d_ob));
    SSDK_TRY(stage_h2d(ctx, 4, (const float*)nullptr, M, &d_os));
    SSDK_TRY(stage_h2d(ctx, 5, (const int32_t*)nullptr, M, &d_oc));
    SSDK_TRY(stage_h2d(ctx, 6, (const int32_t*)nullptr, (size_t)B, &d_on));
    if (B == 0) return SSDK_OK;
    SSDK_TRY(ssdk_postprocess(ctx, d_codes, d_anchors, d_scores, flags, B, A, C, score_threshold, iou_threshold, K, d_ob, d_os,
                              d_oc, d_on, nullptr));
    SSDK_CHECK_CUDA(cudaMemcpyAsync(out_boxes, d_ob, M * 16, cudaMemcpyDeviceToHost, ctx->stream));
    SSDK_CHECK_CUDA(cudaMemcpyAsync(out_scores, d_os, M * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SSDK_CHECK_CUDA(cudaMemcpyAsync(out_classes, d_oc, M * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SSDK_CHECK_CUDA(cudaMemcpyAsync(out_num, d_on, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SSDK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SSDK_OK;
}

}  // extern "C"
