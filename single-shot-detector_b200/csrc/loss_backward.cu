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
#include "stream.cuh"

// d/dx of the negative-class term, without the (1-alpha) * u/N factor
template <int GAMMA_MODE>
__device__ __forceinline__ float focal_negative_grad(float x, float gamma) {
    const float e = ex2_approx(-fabsf(x) * 1.4426950408889634f);
    const float r = rcp_approx(1.0f + e);
    const float er = e * r;
    const float p = (x >= 0.0f) ? r : er;                    // sigmoid(x)
    const float q = (x >= 0.0f) ? er : r;                    // 1 - sigmoid(x)
    const float sp = fmaxf(x, 0.0f) + log1p_unit(e);         // softplus(x)
    if (GAMMA_MODE == 0) return p * p * fmaf(2.0f * q, sp, p);
    return powf(p, gamma) * fmaf(gamma * q, sp, p);
}

// d/dx of the positive-class term, without the alpha * u/N factor (at most one per anchor row: libm)
template <int GAMMA_MODE>
__device__ __forceinline__ float focal_positive_grad(float x, float gamma) {
    const float p = 1.0f / (1.0f + expf(-x));
    const float q = 1.0f / (1.0f + expf(x));                 // 1 - p without cancellation
    const float logp = -(fmaxf(-x, 0.0f) + log1pf(expf(-fabsf(x))));
    const float mod = (GAMMA_MODE == 0) ? q * q : powf(q, gamma);
    return mod * (gamma * p * logp - q);
}

__device__ __forceinline__ float smooth_l1_grad(float p, float t) {
    const float d = p - t;
    return (fabsf(d) < 1.0f) ? d : ((d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f));
}

template <int GAMMA_MODE>
__global__ void __launch_bounds__(LOSS_THREADS) ssd_loss_backward_kernel(
    const float* __restrict__ logits, const float4* __restrict__ codes, const float4* __restrict__ reg_t,
    const int* __restrict__ cls_t, const int* __restrict__ matches, long long NA, int C, int rpw, float gamma, float alpha,
    const double* __restrict__ sums, const float* __restrict__ upstream, LossSmemLayout L, float* __restrict__ grad_logits,
    float4* __restrict__ grad_codes) {
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
    const double norm = fmax(sums[2], 1.0);                              // ssd.py:123 (global count)
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
            }
            grad_codes[n0 + lane] = g;
        }
        // ---- positive-class logits are read before the flat pass overwrites them
        const bool special = __any_sync(0xffffffffu, m != -1);
        int tc = -1;
        float xpos = 0.0f;
        if (special && row_lane && m >= -1) {
            tc = s_c[lane] - 1;
            if (tc >= 0 && tc < C) xpos = s_x[lane * C + tc];
            else tc = -1;
        }
        __syncwarp();
        // ---- flat pass, in place: every element as a negative
        float4* x4 = (float4*)s_x;
#pragma unroll 2
        for (int i = lane; i < n4; i += 32) {
            float4 v = x4[i];
            v.x = k_neg * focal_negative_grad<GAMMA_MODE>(v.x, gamma);
            v.y = k_neg * focal_negative_grad<GAMMA_MODE>(v.y, gamma);
            v.z = k_neg * focal_negative_grad<GAMMA_MODE>(v.z, gamma);
            v.w = k_neg * focal_negative_grad<GAMMA_MODE>(v.w, gamma);
            x4[i] = v;
        }
        // ---- fix-ups: the positive class of matched rows, ignored rows (weight 0, ssd.py:103)
        if (special) {
            __syncwarp();
            if (row_lane) {
                float* x = s_x + lane * C;
                if (m < -1) {
                    for (int c = 0; c < C; ++c) x[c] = 0.0f;
                } else if (tc >= 0) {
                    x[tc] = k_pos * focal_positive_grad<GAMMA_MODE>(xpos, gamma);
                }
            }
        }
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
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

extern "C" int ssdk_ssd_loss_backward(ssdk_ctx* ctx, const float* logits, const float* codes, const float* reg_targets,
                                      const int32_t* cls_targets, const int32_t* matches, int64_t B, int64_t A, int C,
                                      double gamma, double alpha, const double* sums, const float* upstream,
                                      float* grad_logits, float* grad_codes) {
    SSDK_TRY(ssdk_ctx_enter(ctx));
    SSDK_REQUIRE(B >= 0 && A >= 0 && C > 0, SSDK_ERR_ARG, "ssdk_ssd_loss_backward: bad sizes");
    const long long NA = (long long)B * A;
    if (NA == 0) return SSDK_OK;
    SSDK_REQUIRE(logits && codes && reg_targets && cls_targets && matches && sums && grad_logits && grad_codes, SSDK_ERR_ARG,
                 "ssdk_ssd_loss_backward: null pointer");
    SSDK_REQUIRE(aligned16(logits) && aligned16(codes) && aligned16(reg_targets) && aligned16(cls_targets) && aligned16(matches) &&
                     aligned16(grad_logits) && aligned16(grad_codes),
                 SSDK_ERR_SHAPE, "ssdk_ssd_loss_backward: tensors must be 16-byte aligned");
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
    SSDK_REQUIRE(128 + 3 * (size_t)L.stage_bytes <= 220 * 1024, SSDK_ERR_SHAPE,
                 "ssdk_ssd_loss_backward: num_classes %d too large for the fused kernel (limit about 570)", C);
    if (L.stage_bytes < 12 * 1024) L.stages = 4;
    const size_t smem = 128 + (size_t)L.stages * L.stage_bytes;
    const long long ntiles = (NA + rows - 1) / rows;
    int per_sm = (int)((227 * 1024) / (smem + 1024 + 256));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    long long grid = (long long)ctx->num_sms * per_sm;
    if (grid > ntiles) grid = ntiles;
    if (gamma == 2.0) {
        SSDK_TRY(ssdk_set_max_smem(ctx, (const void*)ssd_loss_backward_kernel<0>, (int)smem));
        SSDK_KERNEL(ctx, SSDK_K_LOSS_BACKWARD,
                    ssd_loss_backward_kernel<0><<<(int)grid, LOSS_THREADS, smem, ctx->stream>>>(
                        logits, (const float4*)codes, (const float4*)reg_targets, cls_targets, matches, NA, C, rpw, (float)gamma,
                        (float)alpha, sums, upstream, L, grad_logits, (float4*)grad_codes));
    } else {
        SSDK_TRY(ssdk_set_max_smem(ctx, (const void*)ssd_loss_backward_kernel<1>, (int)smem));
        SSDK_KERNEL(ctx, SSDK_K_LOSS_BACKWARD,
                    ssd_loss_backward_kernel<1><<<(int)grid, LOSS_THREADS, smem, ctx->stream>>>(
                        logits, (const float4*)codes, (const float4*)reg_targets, cls_targets, matches, NA, C, rpw, (float)gamma,
                        (float)alpha, sums, upstream, L, grad_logits, (float4*)grad_codes));
    }
    return SSDK_OK;
}
