// Fused SSD loss: sigmoid focal classification loss + smooth-L1 localisation loss + matched-anchor count.
// Replaces detector/ssd.py:89-133 (SSD.loss after target creation) with detector/losses.py:4-50 inlined.
//
// The reference materialises one_hot(cls_targets)[B,A,C+1], its slice, nlpt, p, p_t, the modulating factor, the
// weighted loss and their product -- eight [B,A,C] float tensors -- and reads the logits several times.  Here
// the logits are streamed from HBM exactly once by a persistent, warp-specialised kernel:
//   * CTA = 8 consumer warps + 1 producer warp, a ring of shared-memory stages with full/empty mbarriers.
//     The producer's elected lane issues TMA bulk copies (cp.async.bulk global->shared, complete_tx on the
//     stage's `full` barrier) of the next tile: 8*rpw consecutive anchors = 8*rpw*C contiguous floats, plus the
//     tile's matches / cls_targets.  No __syncthreads in the steady state.
//   * each consumer warp owns rpw rows of the tile (a contiguous, 16-byte aligned run of rpw*C floats):
//       - lanes < rpw do the per-row work: smooth-L1 + matched count for matched rows, and the row "patches":
//         the positive class of a matched row is evaluated apart (full-precision libm) and overwritten with
//         -inf, an ignored row (matches == -2, weight 0, ssd.py:103) is overwritten with -inf entirely;
//       - then the run is summed FLAT as if every element were a negative: 128-bit LDS, no index math,
//         focal_negative(-inf) == 0 exactly.  Per element: e = exp(-|x|) (MUFU.EX2), r = 1/(1+e) (MUFU.RCP),
//         1-p_t = x>=0 ? r : 1-r, softplus = max(x,0) + e*P7(e) (degree-7 minimax polynomial for log1p(e)/e,
//         relative error 2e-7); the polynomial and the products run on packed FFMA2/FMUL2/FADD2
//         (fma.rn.f32x2, two elements per issue slot) because the kernel is issue-bound, not FMA-pipe-bound;
//       - the warp releases the stage (fence.proxy.async + arrive on `empty`).
//     When per-anchor outputs are requested (PER_ANCHOR) a row mapping is used instead: lanes over classes,
//     shuffle reduction per row, the reference's 1-(1-p) rounding reproduced per element.
//   * sums are carried per thread in double across tiles, reduced warp -> CTA, written as per-CTA partials and
//     combined in a fixed order by a second kernel: deterministic, no float atomics.
#include <stdlib.h>

#include "focal_math.cuh"

// ---------------------------------------------------------------------------------------------- kernel

template <int GAMMA_MODE, bool PER_ANCHOR>
__global__ void __launch_bounds__(LOSS_THREADS, 4) ssd_loss_kernel(
    const float* __restrict__ logits, const float4* __restrict__ codes, const float4* __restrict__ reg_t,
    const int* __restrict__ cls_t, const int* __restrict__ matches, long long NA, int C, int rpw /*rows per warp*/,
    float gamma, float alpha, float one_minus_alpha, LossSmemLayout L, float* __restrict__ cls_losses,
    float* __restrict__ loc_losses, double* __restrict__ partials /*[grid][3]*/, unsigned* __restrict__ ticket /*zero between launches*/,
    double* __restrict__ out_sums /*[3]*/) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* full = (unsigned long long*)smem;                 // [stages]  producer -> consumers
    unsigned long long* empty = full + LOSS_MAX_STAGES;                   // [stages]  consumers -> producer
    unsigned char* stage0 = smem + 128;
    __shared__ double s_red[LOSS_CONSUMER_WARPS][3];

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
        unsigned s = 0, wrapped = 0, parity = 1;              // `empty` phase to wait for once the ring has wrapped (toggles per wrap: 0, 1, ...)
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
                // ragged last tile of the whole problem: plain loads by the producer warp, rows beyond NA are
                // marked ignored (-2) so that the consumers skip them
                const int nrows = (int)(NA - n0);
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
    // Everything that does not change from tile to tile is hoisted: the kernel is issue-bound, and the 8-element
    // steps of the flat sum should be (almost) all a warp executes per tile.
    double acc_cls = 0.0, acc_loc = 0.0, acc_cnt = 0.0;
    const int r0 = warp * rpw;                                           // first row of this warp inside a tile
    const unsigned x_off = (unsigned)r0 * (unsigned)C * 4u;              // byte offset of the warp's rpw*C floats in a stage
    const unsigned m_off = L.tile_bytes + (unsigned)r0 * 4u;
    const unsigned c_off = m_off + L.meta_bytes;
    const int n4 = (rpw * C) >> 2;                                       // rpw % 4 == 0 -> exact
    const int full_steps = n4 >> 6;                                      // steps of 64 float4 (8 elements per lane)
    const int rem = n4 & 63;                                             // float4 left for the (predicated) last step
    const bool row_lane = lane < rpw;
    unsigned s = 0, parity = 0;
    unsigned char* st = stage0;
    long long n0 = first * rows + r0;                                    // global anchor index of the warp's first row
    const long long n_step = step * rows;
    for (long long tile = first; tile < ntiles; tile += step) {
        float* s_x = (float*)(st + x_off);
        const int* s_m = (const int*)(st + m_off);
        const int* s_c = (const int*)(st + c_off);
        mbar_wait(&full[s], parity);

        // ---- per-row work, one lane per anchor row.  Background rows (matches == -1, the overwhelming majority) need
        //      none: the whole block is skipped warp-uniformly unless some row of this warp is matched or ignored.
        float tile_acc = 0.0f;
        const int m = row_lane ? s_m[lane] : -1;
        const bool special = __any_sync(0xffffffffu, m != -1) || PER_ANCHOR || (loc_losses != nullptr);
        if (special) {
            // localisation loss + matched count (ssd.py:89,117,121); rows beyond NA carry -2 (ragged tile)
            if (row_lane && n0 + lane < NA) {
                float l = 0.0f;
                if (m >= 0) {
                    l = smooth_l1_4(codes[n0 + lane], reg_t[n0 + lane]);
                    acc_loc += (double)l;
                    acc_cnt += 1.0;
                }
                if (loc_losses) loc_losses[n0 + lane] = l;
            }
        }

        if (PER_ANCHOR) {
            // ---- focal loss, row mapping: lanes over classes (ssd.py:96-109, losses.py:34-50)
            float mine = 0.0f;                       // lane r keeps the loss of row r
            for (int r = 0; r < rpw; ++r) {
                const int mr = __shfl_sync(0xffffffffu, m, r);
                float row = 0.0f;
                if (mr >= -1) {                                      // not_ignore (ssd.py:103)
                    const int tc = s_c[r] - 1;                       // one_hot(cls, C+1)[1:] -> class index, -1 = background
                    const float* x = s_x + r * C;
                    float neg = 0.0f;
                    if (tc < 0 || tc >= C) {
                        for (int c = lane; c < C; c += 32) neg += focal_negative<GAMMA_MODE>(x[c], gamma);
                        row = one_minus_alpha * neg;
                    } else {
                        for (int c = lane; c < C; c += 32)
                            if (c != tc) neg += focal_negative<GAMMA_MODE>(x[c], gamma);
                        row = one_minus_alpha * neg;
                        if (lane == (tc & 31)) row += alpha * focal_positive<GAMMA_MODE>(x[tc], gamma);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) row += __shfl_xor_sync(0xffffffffu, row, o);
                if (lane == r) mine = row;
            }
            if (row_lane && n0 + lane < NA) {
                cls_losses[n0 + lane] = mine;
                tile_acc = mine;
            }
        } else {
            // ---- patches (see the header), then the flat sum
            if (special) {
                if (row_lane) {
                    float* x = s_x + lane * C;
                    if (m < -1) {
                        for (int c = 0; c < C; ++c) x[c] = -INFINITY;
                    } else {
                        const int tc = s_c[lane] - 1;
                        if (tc >= 0 && tc < C) {
                            tile_acc = alpha * focal_positive<GAMMA_MODE>(x[tc], gamma);
                            x[tc] = -INFINITY;
                        }
                    }
                }
                __syncwarp();
            }
            const float4* x4 = (const float4*)s_x;
            float s0, s1;
            bool general = (GAMMA_MODE != 0);
            if (GAMMA_MODE == 0) {
                // fast path: every element assumed < +0 (checked through `allneg`); 8 elements per lane and step
                f32x2 acc[4] = {0ull, 0ull, 0ull, 0ull};
                unsigned allneg = 0x80000000u;
                const float4* xp = x4 + lane;
#pragma unroll 1
                for (int it = 0; it < full_steps; ++it, xp += 64) focal_fast8(xp[0], xp[32], acc, allneg);
                if (rem) {
                    const float4 ninf4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                    const float4 v = (lane < rem) ? xp[0] : ninf4;
                    const float4 w = (lane + 32 < rem) ? xp[32] : ninf4;
                    focal_fast8(v, w, acc, allneg);
                }
                unpack2(add2(add2(acc[0], acc[1]), add2(acc[2], acc[3])), s0, s1);
                general = __any_sync(0xffffffffu, (allneg >> 31) == 0u);
            }
            if (general) {
                // some logit of this warp's rows is >= 0 (or gamma != 2): re-sum with the general form
                f32x2 a01 = 0ull, a23 = 0ull;
#pragma unroll 2
                for (int i = lane; i < n4; i += 32) {
                    const float4 v = x4[i];
                    a01 = focal_negative2<GAMMA_MODE>(v.x, v.y, gamma, a01);
                    a23 = focal_negative2<GAMMA_MODE>(v.z, v.w, gamma, a23);
                }
                unpack2(add2(a01, a23), s0, s1);
            }
            tile_acc += one_minus_alpha * (s0 + s1);
        }
        acc_cls += (double)tile_acc;

        // ---- release the stage.  Generic-proxy WRITES (the patches) must be ordered before the next TMA write to the
        //      stage (fence.proxy.async); plain reads need no proxy fence.
        __syncwarp();
        if (lane == 0) {
            if (special && !PER_ANCHOR) fence_proxy_async();
            mbar_arrive(&empty[s]);
        }
        n0 += n_step;
        st += L.stage_bytes;
        if (++s == L.stages) { s = 0; st = stage0; parity ^= 1u; }
    }

    // ---- CTA reduction (fixed order) -> partials[blockIdx.x]
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_cls += __shfl_xor_sync(0xffffffffu, acc_cls, o);
        acc_loc += __shfl_xor_sync(0xffffffffu, acc_loc, o);
        acc_cnt += __shfl_xor_sync(0xffffffffu, acc_cnt, o);
    }
    if (lane == 0) { s_red[warp][0] = acc_loc; s_red[warp][1] = acc_cls; s_red[warp][2] = acc_cnt; }
    asm volatile("bar.sync 1, %0;" ::"n"(LOSS_CONSUMER_WARPS * 32) : "memory");   // consumers only (the producer has left)
    if (tid < 3) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < LOSS_CONSUMER_WARPS; ++w) t += s_red[w][tid];
        partials[(size_t)blockIdx.x * 3 + tid] = t;
        __threadfence();
    }
    // ---- the CTA that finishes last sums the per-CTA partials in a fixed order (deterministic whichever CTA it is)
    //      -> out_sums[3] = { sum loc, sum cls, num_matches }, and re-arms the ticket counter for the next launch
    __shared__ int s_last;
    __shared__ double s_fin[64][3];
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

// normalizer = max(num_matches, 1) and the two scalar losses: ssd.py:123,131-133
__global__ void loss_finalize_kernel(const double* __restrict__ sums, float* __restrict__ out) {
    const double norm = fmax(sums[2], 1.0);
    out[0] = (float)(sums[0] / norm);
    out[1] = (float)(sums[1] / norm);
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <int GM, bool PA>
static int launch_loss(ssdk_ctx* ctx, int grid, size_t smem, const float* logits, const float* codes, const float* reg_t,
                       const int* cls_t, const int* matches, long long NA, int C, int rpw, double gamma, double alpha,
                       LossSmemLayout L, float* cls_losses, float* loc_losses, double* partials, unsigned* ticket, double* out_sums) {
    auto kern = ssd_loss_kernel<GM, PA>;
    SSDK_TRY(ssdk_set_max_smem(ctx, (const void*)kern, (int)smem));
    SSDK_KERNEL(ctx, SSDK_K_LOSS,
                kern<<<grid, LOSS_THREADS, smem, ctx->stream>>>(logits, (const float4*)codes, (const float4*)reg_t, cls_t, matches,
                                                                NA, C, rpw, (float)gamma, (float)alpha, (float)(1.0 - alpha), L,
                                                                cls_losses, loc_losses, partials, ticket, out_sums));
    return SSDK_OK;
}

extern "C" {

int ssdk_ssd_loss(ssdk_ctx* ctx, const float* logits, const float* codes, const float* reg_targets,
                  const int32_t* cls_targets, const int32_t* matches, int64_t B, int64_t A, int C, double gamma,
                  double alpha, double* out_sums, float* out_cls_losses, float* out_loc_losses) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(B >= 0 && A >= 0 && C > 0, SSDK_ERR_ARG, "ssdk_ssd_loss: bad sizes");
    SSDK_REQUIRE(out_sums != nullptr, SSDK_ERR_ARG, "ssdk_ssd_loss: out_sums is NULL");
    SsdkWsGuard ws_guard(ctx, SSDK_WS_LOSS);
    const long long NA = (long long)B * A;
    if (NA == 0) {
        SSDK_CHECK_CUDA(cudaMemsetAsync(out_sums, 0, 3 * sizeof(double), ctx->stream));
        return SSDK_OK;
    }
    SSDK_REQUIRE(logits && codes && reg_targets && cls_targets && matches, SSDK_ERR_ARG, "ssdk_ssd_loss: null pointer");
    SSDK_REQUIRE(aligned16(logits) && aligned16(codes) && aligned16(reg_targets) && aligned16(cls_targets) && aligned16(matches),
                 SSDK_ERR_SHAPE, "ssdk_ssd_loss: inputs must be 16-byte aligned");

    // tile = 8 consumer warps x rpw rows; rpw is a multiple of 4 (so rpw*C*4 and rpw*4 bytes are multiples of 16)
    // and at most 32 (one lane per row); about 24 KB per stage
    int rpw = (int)(24576 / (32 * (long long)C)) / 4 * 4;
    if (ctx->tune_loss_rpw) rpw = ctx->tune_loss_rpw;                    // SSDK_LOSS_RPW (a multiple of 4 in [4,32], validated at context creation)
    if (rpw < 4) rpw = 4;
    if (rpw > 32) rpw = 32;
    const int rows = rpw * LOSS_CONSUMER_WARPS;
    LossSmemLayout L;
    L.tile_bytes = (unsigned)rows * C * 4;
    L.meta_bytes = (unsigned)rows * 4;
    L.stage_bytes = L.tile_bytes + 2 * L.meta_bytes;
    L.stages = 2;
    SSDK_REQUIRE(128 + 2 * (size_t)L.stage_bytes <= 200 * 1024, SSDK_ERR_SHAPE,
                 "ssdk_ssd_loss: num_classes %d too large for the fused kernel (limit about 780); use ssdk_focal_loss", C);
    if (L.stage_bytes < 12 * 1024) L.stages = 4;
    else if (L.stage_bytes < 20 * 1024) L.stages = 3;
    if (ctx->tune_loss_stages) L.stages = (unsigned)ctx->tune_loss_stages;         // SSDK_LOSS_STAGES (2..4)
    if (L.stages < 2) L.stages = 2;
    if (L.stages > LOSS_MAX_STAGES) L.stages = LOSS_MAX_STAGES;
    while (L.stages > 2 && 128 + (size_t)L.stages * L.stage_bytes > 223 * 1024) --L.stages;   // whatever the knob says, the ring must fit
    const size_t smem = 128 + (size_t)L.stages * L.stage_bytes;

    const long long ntiles = (NA + rows - 1) / rows;
    int per_sm = (int)((227 * 1024) / (smem + 1024 + 2048));   // + reserved + static shared memory
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 5) per_sm = 5;
    if (ctx->tune_loss_ctas) per_sm = ctx->tune_loss_ctas;                       // SSDK_LOSS_CTAS (1..8: the partials buffer holds num_sms * 8 CTAs)
    while (per_sm > 1 && per_sm * (smem + 3072) > 227 * 1024) --per_sm;          // resident CTAs the shared memory allows
    long long grid = (long long)ctx->num_sms * per_sm;
    if (grid > ntiles) grid = ntiles;

    // per-CTA partials + the ticket counter of the fused final reduction (zeroed when allocated, re-armed by the kernel)
    const size_t part_bytes = 16 + (size_t)ctx->num_sms * 8 * 3 * sizeof(double);
    if (ctx->ws_partials.cap < part_bytes) {
        SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_partials, part_bytes));
        SSDK_CHECK_CUDA(cudaMemsetAsync(ctx->ws_partials.p, 0, 16, ctx->stream));
    }
    unsigned* ticket = (unsigned*)ctx->ws_partials.p;
    double* partials = (double*)((char*)ctx->ws_partials.p + 16);
    const bool pa = out_cls_losses != nullptr;
    const bool g2 = (gamma == 2.0);
    int st;
    if (g2 && !pa) st = launch_loss<0, false>(ctx, (int)grid, smem, logits, codes, reg_targets, cls_targets, matches, NA, C, rpw, gamma, alpha, L, out_cls_losses, out_loc_losses, partials, ticket, out_sums);
    else if (g2 && pa) st = launch_loss<0, true>(ctx, (int)grid, smem, logits, codes, reg_targets, cls_targets, matches, NA, C, rpw, gamma, alpha, L, out_cls_losses, out_loc_losses, partials, ticket, out_sums);
    else if (!pa) st = launch_loss<1, false>(ctx, (int)grid, smem, logits, codes, reg_targets, cls_targets, matches, NA, C, rpw, gamma, alpha, L, out_cls_losses, out_loc_losses, partials, ticket, out_sums);
    else st = launch_loss<1, true>(ctx, (int)grid, smem, logits, codes, reg_targets, cls_targets, matches, NA, C, rpw, gamma, alpha, L, out_cls_losses, out_loc_losses, partials, ticket, out_sums);
    SSDK_TRY(st);
    return SSDK_OK;
}

int ssdk_loss_finalize(ssdk_ctx* ctx, const double* sums, float* out_losses) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(sums && out_losses, SSDK_ERR_ARG, "ssdk_loss_finalize: null pointer");
    SSDK_KERNEL(ctx, SSDK_K_OTHER, loss_finalize_kernel<<<1, 1, 0, ctx->stream>>>(sums, out_losses));
    return SSDK_OK;
}

}  // extern "C"
