// Backward of the fused SSD loss: gradients of  u_loc * localization_loss + u_cls * classification_loss  with respect to
// the two head tensors, class_predictions [B,A,C] and encoded_boxes [B,A,4].
//
// This is what the reference obtains from TensorFlow autodiff (model.py:115-118, optimizer.compute_gradients on
// total_loss = loc_weight * localization_loss + cls_weight * classification_loss, model.py:86-91) through
// detector/ssd.py:89-133 and detector/losses.py:4-50; targets, matches and weights are constants there
// (tf.map_fn(..., back_prop=False), ssd.py:197; tf.stop_gradient in losses.py:30,33,46).  With N = max(num_matches, 1)
// (ssd.py:123, the GLOBAL count after the all-reduce) the closed forms are, per anchor a and class c
// (p = sigmoid(x), sp = softplus(x), z = one-hot target, w = not_ignore weight, v = matched weight):
//     z = 0:  d/dx = w (1-alpha) p^gamma     [ gamma (1-p) sp + p ]            * u_cls / N
//     z = 1:  d/dx = w  alpha   (1-p)^gamma  [ gamma p log p - (1-p) ]         * u_cls / N
//     codes:  d/dp = v (|d| < 1 ? d : sign(d)),  d = p - t                     * u_loc / N      (losses.py:16-19)
//
// Streaming structure = ssd_loss_kernel's: persistent CTAs, one producer warp issuing TMA bulk loads of 64-anchor logit
// tiles into a shared-memory ring, 8 consumer warps that each own 8 anchor rows.  A consumer turns its rows into
// gradients IN PLACE (flat pass with the negative-class form, then the few positive-class / ignored-row fix-ups), and
// its elected lane hands the finished run to the TMA engine (cp.async.bulk shared -> global); the stage is released to
// the producer one tile later, once that store has finished reading shared memory.  Algorithmic traffic: 4AC read +
// 4AC written (+ 48A for codes / targets / grad_codes): the kernel is HBM-bound, the math (2 MUFU + ~20 FMA-class ops
// per element) fits under the 8-bytes-per-element budget.
#include "focal_math.cuh"

// WITH_LOSS: the same pass also accumulates the un-normalised loss sums (forward + backward in one read of the logits);
// `norm_count` then is the matched count obtained BEFORE the pass (ssdk_count_matches, all-reduced on several GPUs).
template <int GAMMA_MODE, bool WITH_LOSS>
__global__ void __launch_bounds__(LOSS_THREADS) ssd_loss_backward_kernel(
    const float* __restrict__ logits, const float4* __restrict__ codes, const float4* __restrict__ reg_t,
    const int* __restrict__ cls_t, const int* __restrict__ matches, long long NA, int C, int rpw, float gamma, float alpha,
    const double* __restrict__ norm_count, const float* __restrict__ upstream, LossSmemLayout L, float* __restrict__ grad_logits,
    float4* __restrict__ grad_codes, double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ out_sums) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* full = (unsigned long long*)smem;                 // [stages]  producer -> consumers
    unsigned long long* empty = full + LOSS_MAX_STAGES;                   // [stages]  consumers -> producer
    unsigned char* stage0 = smem + 128;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rows = rpw * LOSS_CONSUMER_WARPS;
    const long long ntiles = (NA + rows - 1) / rows;
    const long long first = blockIdx.x, step = gridDim.x;

    if (tid == 0) {
        for (unsigned s = 0; s < L.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], LOSS_CONSUMER_WARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == LOSS_CONSUMER_WARPS) {
        // =========================================================================== producer warp
        unsigned s = 0, wrapped = 0, parity = 1;
        unsigned char* st = stage0;
        for (long long tile = first; tile < ntiles; tile += step) {
            if (wrapped) mbar_wait(&empty[s], parity);
            const long long n0 = tile * rows;
            if (n0 + rows <= NA) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[s], L.tile_bytes + 2 * L.meta_bytes);
                    bulk_g2s(st, logits + n0 * C, L.tile_bytes, &full[s]);
                    bulk_g2s(st + L.tile_bytes, matches + n0, L.meta_bytes, &full[s]);
                    bulk_g2s(st + L.tile_bytes + L.meta_bytes, cls_t + n0, L.meta_bytes, &full[s]);
                }
            } else {
                const int nrows = (int)(NA - n0);                         // ragged last tile: plain loads
                float* sx = (float*)st;
                int* sm = (int*)(st + L.tile_bytes);
                int* sc = (int*)(st + L.tile_bytes + L.meta_bytes);
                for (int i = lane; i < rows * C; i += 32) sx[i] = (i < nrows * C) ? logits[n0 * C + i] : -INFINITY;
                for (int i = lane; i < rows; i += 32) {
                    sm[i] = (i < nrows) ? matches[n0 + i] : -2;
                    sc[i] = (i < nrows) ? cls_t[n0 + i] : 0;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
            st += L.stage_bytes;
            if (++s == L.stages) { s = 0; st = stage0; wrapped = 1; parity ^= 1u; }
        }
        return;
    }

    // =============================================================================== consumer warps
    const double norm = fmax(*norm_count, 1.0);                          // ssd.py:123 (global count)
    double acc_cls = 0.0, acc_loc = 0.0, acc_cnt = 0.0;                  // WITH_LOSS: un-normalised sums of this thread
    const float u_loc = upstream ? upstream[0] : 1.0f, u_cls = upstream ? upstream[1] : 1.0f;
    const float k_loc = (float)((double)u_loc / norm);
    const float k_neg = (float)((double)u_cls * (1.0 - (double)alpha) / norm);
    const float k_pos = (float)((double)u_cls * (double)alpha / norm);

    const int r0 = warp * rpw;
    const unsigned x_off = (unsigned)r0 * (unsigned)C * 4u;
    const unsigned m_off = L.tile_bytes + (unsigned)r0 * 4u;
    const unsigned c_off = m_off + L.meta_bytes;
    const int n4 = (rpw * C) >> 2;
    const unsigned run_bytes = (unsigned)rpw * (unsigned)C * 4u;
    const bool row_lane = lane < rpw;
    unsigned s = 0, parity = 0;
    unsigned char* st = stage0;
    long long n0 = first * rows + r0;
    const long long n_step = step * rows;
    int prev_s = -1;                                                      // stage whose TMA store is still in flight
    for (long long tile = first; tile < ntiles; tile += step) {
        float* s_x = (float*)(st + x_off);
        const int* s_m = (const int*)(st + m_off);
        const int* s_c = (const int*)(st + c_off);
        mbar_wait(&full[s], parity);

        const int m = row_lane ? s_m[lane] : -1;
        // ---- encoded_boxes gradient, one lane per anchor row (every row is written: zeros unless matched)
        if (row_lane && n0 + lane < NA) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m >= 0) {
                const float4 p = codes[n0 + lane], t = reg_t[n0 + lane];
                g = make_float4(k_loc * smooth_l1_grad(p.x, t.x), k_loc * smooth_l1_grad(p.y, t.y),
                                k_loc * smooth_l1_grad(p.z, t.z), k_loc * smooth_l1_grad(p.w, t.w));
                if (WITH_LOSS) {
                    acc_loc += (double)(smooth_l1_value(p.x, t.x) + smooth_l1_value(p.y, t.y) + smooth_l1_value(p.z, t.z) +
                                        smooth_l1_value(p.w, t.w));
                    acc_cnt += 1.0;
                }
            }
            grad_codes[n0 + lane] = g;
        }
        // ---- patches, as in the forward kernel: the positive-class logit of a matched row is read and replaced by -inf,
        //      an ignored row (weight 0, ssd.py:103) is replaced by -inf entirely; the negative form maps -inf to value 0
        //      and gradient 0, so the flat pass needs no per-element target
        const bool special = __any_sync(0xffffffffu, m != -1);
        int tc = -1;
        float xpos = 0.0f;
        if (special) {
            if (row_lane) {
                float* x = s_x + lane * C;
                if (m < -1) {
                    for (int c = 0; c < C; ++c) x[c] = -INFINITY;
                } else {
                    tc = s_c[lane] - 1;
                    if (tc >= 0 && tc < C) { xpos = x[tc]; x[tc] = -INFINITY; }
                    else tc = -1;
                }
            }
            __syncwarp();
        }
        // ---- flat pass, in place: every element as a negative
        float4* x4 = (float4*)s_x;
        float tile_neg = 0.0f, tile_fix = 0.0f;       // WITH_LOSS: sum of negative-form values; corrections for special rows
        if (WITH_LOSS) {
#pragma unroll 2
            for (int i = lane; i < n4; i += 32) {
                float4 v = x4[i];
                float f0, f1, f2, f3;
                v.x = k_neg * focal_negative_both<GAMMA_MODE>(v.x, gamma, f0);
                v.y = k_neg * focal_negative_both<GAMMA_MODE>(v.y, gamma, f1);
                v.z = k_neg * focal_negative_both<GAMMA_MODE>(v.z, gamma, f2);
                v.w = k_neg * focal_negative_both<GAMMA_MODE>(v.w, gamma, f3);
                tile_neg += (f0 + f1) + (f2 + f3);
                x4[i] = v;
            }
        } else {
#pragma unroll 2
            for (int i = lane; i < n4; i += 32) {
                float4 v = x4[i];
                v.x = k_neg * focal_negative_grad<GAMMA_MODE>(v.x, gamma);
                v.y = k_neg * focal_negative_grad<GAMMA_MODE>(v.y, gamma);
                v.z = k_neg * focal_negative_grad<GAMMA_MODE>(v.z, gamma);
                v.w = k_neg * focal_negative_grad<GAMMA_MODE>(v.w, gamma);
                x4[i] = v;
            }
        }
        // ---- the positive class of matched rows
        if (special) {
            __syncwarp();
            if (row_lane && tc >= 0) {
                s_x[lane * C + tc] = k_pos * focal_positive_grad<GAMMA_MODE>(xpos, gamma);
                if (WITH_LOSS) tile_fix = alpha * focal_positive_value<GAMMA_MODE>(xpos, gamma);
            }
        }
        if (WITH_LOSS) acc_cls += (double)fmaf(1.0f - alpha, tile_neg, tile_fix);
        // ---- hand the finished run to the TMA engine; release the PREVIOUS stage once its store has read shared memory
        fence_proxy_async();
        __syncwarp();
        if (n0 + rpw <= NA) {
            if (lane == 0) {
                bulk_s2g(grad_logits + n0 * C, s_x, run_bytes);
                bulk_commit();
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                if (prev_s >= 0) mbar_arrive(&empty[prev_s]);
            }
            prev_s = (int)s;
        } else {
            // ragged end of the problem: plain stores of the valid rows, then release everything
            long long valid = NA - n0;
            if (valid < 0) valid = 0;
            const int nvalid = (int)valid * C;
            for (int i = lane; i < nvalid; i += 32) grad_logits[n0 * C + i] = s_x[i];
            __syncwarp();
            if (lane == 0) {
                bulk_wait_read0();
                if (prev_s >= 0) mbar_arrive(&empty[prev_s]);
                mbar_arrive(&empty[s]);
            }
            prev_s = -1;
        }
        n0 += n_step;
        st += L.stage_bytes;
        if (++s == L.stages) { s = 0; st = stage0; parity ^= 1u; }
    }
    if (lane == 0) bulk_wait0();                                          // stores must be complete before the CTA exits
    if (!WITH_LOSS) return;

    // ---- loss sums: warp -> CTA partial -> the last CTA adds the partials in a fixed order (as ssd_loss_kernel)
    __shared__ double s_red[LOSS_CONSUMER_WARPS][3];
    __shared__ double s_fin[64][3];
    __shared__ int s_last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_cls += __shfl_xor_sync(0xffffffffu, acc_cls, o);
        acc_loc += __shfl_xor_sync(0xffffffffu, acc_loc, o);
        acc_cnt += __shfl_xor_sync(0xffffffffu, acc_cnt, o);
    }
    if (lane == 0) { s_red[warp][0] = acc_loc; s_red[warp][1] = acc_cls; s_red[warp][2] = acc_cnt; }
    asm volatile("bar.sync 1, %0;" ::"n"(LOSS_CONSUMER_WARPS * 32) : "memory");
    if (tid < 3) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < LOSS_CONSUMER_WARPS; ++w) t += s_red[w][tid];
        partials[(size_t)blockIdx.x * 3 + tid] = t;
        __threadfence();
    }
    asm volatile("bar.sync 1, %0;" ::"n"(LOSS_CONSUMER_WARPS * 32) : "memory");
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    asm volatile("bar.sync 1, %0;" ::"n"(LOSS_CONSUMER_WARPS * 32) : "memory");
    if (!s_last) return;
    __threadfence();
    if (tid < 64) {
        double t[3] = {0.0, 0.0, 0.0};
        for (int i = tid; i < (int)gridDim.x; i += 64)
            for (int j = 0; j < 3; ++j) t[j] += __ldcg(&partials[(size_t)i * 3 + j]);
        for (int j = 0; j < 3; ++j) s_fin[tid][j] = t[j];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(LOSS_CONSUMER_WARPS * 32) : "memory");
    for (int o = 32; o > 0; o >>= 1) {
        if (tid < o)
            for (int j = 0; j < 3; ++j) s_fin[tid][j] += s_fin[tid + o][j];
        asm volatile("bar.sync 1, %0;" ::"n"(LOSS_CONSUMER_WARPS * 32) : "memory");
    }
    if (tid < 3) out_sums[tid] = s_fin[0][tid];
    if (tid == 0) *ticket = 0u;
}

// number of matched anchors (matches >= 0) of this shard, as a double: the loss normaliser's input when it is needed
// BEFORE the loss pass (fused forward + backward); ssd.py:89,121-122
__global__ void __launch_bounds__(256) count_matches_kernel(const int* __restrict__ matches, long long n, double* __restrict__ out) {
    __shared__ int s_cnt[8];
    int c = 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) c += matches[i] >= 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_cnt[w];
        if (t) atomicAdd(out, (double)t);          // integers: exact and order independent
    }
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

static int launch_backward(ssdk_ctx* ctx, bool with_loss, const float* logits, const float* codes, const float* reg_targets,
                           const int32_t* cls_targets, const int32_t* matches, int64_t B, int64_t A, int C, double gamma,
                           double alpha, const double* norm_count, const float* upstream, float* grad_logits, float* grad_codes,
                           double* out_sums) {
    const long long NA = (long long)B * A;
    // same tile geometry as the forward kernel; three stages because a stage is released one tile late
    int rpw = (int)(24576 / (32 * (long long)C)) / 4 * 4;
    if (rpw < 4) rpw = 4;
    if (rpw > 32) rpw = 32;
    const int rows = rpw * LOSS_CONSUMER_WARPS;
    LossSmemLayout L;
    L.tile_bytes = (unsigned)rows * C * 4;
    L.meta_bytes = (unsigned)rows * 4;
    L.stage_bytes = L.tile_bytes + 2 * L.meta_bytes;
    L.stages = 3;
    SSDK_REQUIRE(128 + 3 * (size_t)L.stage_bytes <= 216 * 1024, SSDK_ERR_SHAPE,
                 "ssdk_ssd_loss_backward: num_classes %d too large for the fused kernel (limit about 570)", C);
    if (L.stage_bytes < 12 * 1024) L.stages = 4;
    const size_t smem = 128 + (size_t)L.stages * L.stage_bytes;
    const long long ntiles = (NA + rows - 1) / rows;
    int per_sm = (int)((227 * 1024) / (smem + 1024 + 2048));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    long long grid = (long long)ctx->num_sms * per_sm;
    if (grid > ntiles) grid = ntiles;
    double* partials = nullptr;
    unsigned* ticket = nullptr;
    if (with_loss) {
        const size_t part_bytes = 16 + (size_t)ctx->num_sms * 8 * 3 * sizeof(double);
        if (ctx->ws_partials.cap < part_bytes) {
            SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_partials, part_bytes));
            SSDK_CHECK_CUDA(cudaMemsetAsync(ctx->ws_partials.p, 0, 16, ctx->stream));
        }
        ticket = (unsigned*)ctx->ws_partials.p;
        partials = (double*)((char*)ctx->ws_partials.p + 16);
    }
#define SSDK_LAUNCH_BW(GM, WL)                                                                                              \
    do {                                                                                                                    \
        SSDK_TRY(ssdk_set_max_smem(ctx, (const void*)ssd_loss_backward_kernel<GM, WL>, (int)smem));                         \
        SSDK_KERNEL(ctx, SSDK_K_LOSS_BACKWARD,                                                                              \
                    ssd_loss_backward_kernel<GM, WL><<<(int)grid, LOSS_THREADS, smem, ctx->stream>>>(                       \
                        logits, (const float4*)codes, (const float4*)reg_targets, cls_targets, matches, NA, C, rpw, (float)gamma, \
                        (float)alpha, norm_count, upstream, L, grad_logits, (float4*)grad_codes, partials, ticket, out_sums)); \
    } while (0)
    const bool g2 = (gamma == 2.0);
    if (g2 && !with_loss) SSDK_LAUNCH_BW(0, false);
    else if (g2) SSDK_LAUNCH_BW(0, true);
    else if (!with_loss) SSDK_LAUNCH_BW(1, false);
    else SSDK_LAUNCH_BW(1, true);
#undef SSDK_LAUNCH_BW
    return SSDK_OK;
}

static int check_backward_args(const char* who, const float* logits, const float* codes, const float* reg_targets,
                               const int32_t* cls_targets, const int32_t* matches, const void* norm, const float* grad_logits,
                               const float* grad_codes) {
    SSDK_REQUIRE(logits && codes && reg_targets && cls_targets && matches && norm && grad_logits && grad_codes, SSDK_ERR_ARG,
                 "%s: null pointer", who);
    SSDK_REQUIRE(aligned16(logits) && aligned16(codes) && aligned16(reg_targets) && aligned16(cls_targets) && aligned16(matches) &&
                     aligned16(grad_logits) && aligned16(grad_codes),
                 SSDK_ERR_SHAPE, "%s: tensors must be 16-byte aligned", who);
    return SSDK_OK;
}

// adds the number of entries >= 0 to *out_count (which the caller has zeroed)
int ssdk_count_impl(ssdk_ctx* ctx, const int32_t* matches, int64_t n, double* out_count) {
    if (n == 0) return SSDK_OK;
    long long blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > ctx->num_sms * 8) blocks = ctx->num_sms * 8;
    SSDK_KERNEL(ctx, SSDK_K_OTHER, count_matches_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(matches, n, out_count));
    return SSDK_OK;
}

extern "C" {

int ssdk_ssd_loss_backward(ssdk_ctx* ctx, const float* logits, const float* codes, const float* reg_targets,
                           const int32_t* cls_targets, const int32_t* matches, int64_t B, int64_t A, int C, double gamma,
                           double alpha, const double* sums, const float* upstream, float* grad_logits, float* grad_codes) {
    SSDK_ENTER(ctx);
    SsdkWsGuard ws_guard(ctx, SSDK_WS_LOSS);
    SSDK_REQUIRE(B >= 0 && A >= 0 && C > 0, SSDK_ERR_ARG, "ssdk_ssd_loss_backward: bad sizes");
    if ((long long)B * A == 0) return SSDK_OK;
    SSDK_TRY(check_backward_args("ssdk_ssd_loss_backward", logits, codes, reg_targets, cls_targets, matches, sums, grad_logits, grad_codes));
    return launch_backward(ctx, false, logits, codes, reg_targets, cls_targets, matches, B, A, C, gamma, alpha, sums + 2, upstream,
                           grad_logits, grad_codes, nullptr);
}

int ssdk_ssd_loss_forward_backward(ssdk_ctx* ctx, const float* logits, const float* codes, const float* reg_targets,
                                   const int32_t* cls_targets, const int32_t* matches, int64_t B, int64_t A, int C,
                                   double gamma, double alpha, const double* num_matches, const float* upstream,
                                   double* out_sums, float* grad_logits, float* grad_codes) {
    SSDK_ENTER(ctx);
    SsdkWsGuard ws_guard(ctx, SSDK_WS_LOSS);
    SSDK_REQUIRE(B >= 0 && A >= 0 && C > 0 && out_sums, SSDK_ERR_ARG, "ssdk_ssd_loss_forward_backward: bad arguments");
    if ((long long)B * A == 0) {
        SSDK_CHECK_CUDA(cudaMemsetAsync(out_sums, 0, 3 * sizeof(double), ctx->stream));
        return SSDK_OK;
    }
    SSDK_TRY(check_backward_args("ssdk_ssd_loss_forward_backward", logits, codes, reg_targets, cls_targets, matches, num_matches,
                                 grad_logits, grad_codes));
    return launch_backward(ctx, true, logits, codes, reg_targets, cls_targets, matches, B, A, C, gamma, alpha, num_matches, upstream,
                           grad_logits, grad_codes, out_sums);
}

int ssdk_count_matches(ssdk_ctx* ctx, const int32_t* matches, int64_t n, double* out_count) {
    SSDK_ENTER(ctx);
    SsdkWsGuard ws_guard(ctx, SSDK_WS_LOSS);
    SSDK_REQUIRE(n >= 0 && out_count && (n == 0 || matches), SSDK_ERR_ARG, "ssdk_count_matches: bad arguments");
    SSDK_CHECK_CUDA(cudaMemsetAsync(out_count, 0, sizeof(double), ctx->stream));
    return ssdk_count_impl(ctx, matches, n, out_count);
}

}  // extern "C"
