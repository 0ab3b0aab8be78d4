// Fused SSD loss: sigmoid focal classification loss + smooth-L1 localisation loss + matched-anchor count.
// Replaces detector/ssd.py:89-133 (SSD.loss after target creation) with detector/losses.py:4-50 inlined.
//
// The reference materialises one_hot(cls_targets)[B,A,C+1], its slice, nlpt, p, p_t, the modulating factor, the
// weighted loss and their product -- eight [B,A,C] float tensors -- and reads the logits several times.  Here
// the logits are streamed from HBM exactly once:
//   * a persistent CTA owns a ring of shared-memory stages; one thread issues TMA bulk copies
//     (cp.async.bulk global->shared, mbarrier complete_tx) for the next tiles of `rows` consecutive anchors
//     (rows*C contiguous floats) together with the tile's matches / cls_targets, while all warps compute on the
//     current stage;
//   * the tile is summed FLAT (128-bit LDS, no per-element index math) as if every element were a negative;
//     beforehand one thread per anchor row patches the exceptions in shared memory: the positive class of a
//     matched row is evaluated apart and overwritten with -inf, ignored rows (matches == -2) become -inf
//     (the negative term of -inf is exactly 0).  When per-anchor outputs are requested a row mapping is used
//     instead (warps over rows, lanes over classes, shuffle reduction per row);
//   * per element: e = exp(-|x|) (MUFU.EX2), r = 1/(1+e) (MUFU.RCP), p = x>=0 ? r : e*r,
//     q = 1-(1-p) (the reference's own cancellation, losses.py:38,41), softplus = max(x,0) + e*P7(e) with a
//     degree-7 minimax polynomial for log1p(e)/e on the FMA pipe (rel. error 2e-7), term = q^gamma * softplus;
//   * sums are carried per thread in double across tiles, reduced warp -> CTA, written as per-CTA partials and
//     combined in a fixed order by a second kernel: deterministic, no float atomics.
#include "common.cuh"

#define LOSS_THREADS 256
#define LOSS_WARPS (LOSS_THREADS / 32)
#define LOSS_MAX_ROWS 256

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------------------------- per-element math
// log1p(e) for e in [0,1]:  e * P7(e), minimax fit of log1p(e)/e (relative error 1.9e-7)
__device__ __forceinline__ float log1p_unit(float e) {
    float p = -8.539209655e-03f;
    p = fmaf(p, e, 4.408963566e-02f);
    p = fmaf(p, e, -1.076817184e-01f);
    p = fmaf(p, e, 1.774524181e-01f);
    p = fmaf(p, e, -2.449546295e-01f);
    p = fmaf(p, e, 3.327547979e-01f);
    p = fmaf(p, e, -4.999740540e-01f);
    p = fmaf(p, e, 9.999998057e-01f);
    return p * e;
}

// Negative-class term without the (1-alpha) factor: (1 - p_t)^gamma * nlpt with targets == 0
// (losses.py:36-41: nlpt = max(x,0) + log1p(exp(-|x|)), p_t = 1 - p, modulating factor (1 - p_t)^gamma).
template <int GAMMA_MODE>
__device__ __forceinline__ float focal_negative(float x, float gamma) {
    const float e = ex2_approx(-fabsf(x) * 1.4426950408889634f);
    const float r = rcp_approx(1.0f + e);
    const float p = (x >= 0.0f) ? r : e * r;                 // sigmoid(x)
    const float q = f_sub(1.0f, f_sub(1.0f, p));             // 1 - p_t, rounded as in the reference
    const float nlpt = fmaxf(x, 0.0f) + log1p_unit(e);
    const float mod = (GAMMA_MODE == 0) ? q * q : powf(q, gamma);
    return mod * nlpt;
}

// Positive-class term without the alpha factor (targets == 1): (1 - p)^gamma * (max(x,0) - x + log1p(exp(-|x|))).
// At most one per anchor row: full-precision libm calls.
template <int GAMMA_MODE>
__device__ __forceinline__ float focal_positive(float x, float gamma) {
    const float nlpt = f_add(f_sub(fmaxf(x, 0.0f), x), log1pf(expf(-fabsf(x))));
    const float p = f_div(1.0f, f_add(1.0f, expf(-x)));
    const float q = f_sub(1.0f, p);
    const float mod = (GAMMA_MODE == 0) ? f_mul(q, q) : powf(q, gamma);
    return f_mul(mod, nlpt);
}

// smooth-L1 over the 4 coordinates: losses.py:16-19
__device__ __forceinline__ float smooth_l1_4(const float4 a, const float4 b) {
    const float d[4] = {fabsf(f_sub(a.x, b.x)), fabsf(f_sub(a.y, b.y)), fabsf(f_sub(a.z, b.z)), fabsf(f_sub(a.w, b.w))};
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) s = f_add(s, d[k] < 1.0f ? f_mul(0.5f, f_mul(d[k], d[k])) : f_sub(d[k], 0.5f));
    return s;
}

// ---------------------------------------------------------------------------------------------- kernel
struct LossSmemLayout {
    unsigned tile_bytes;    // rows*C*4 (multiple of 16)
    unsigned meta_bytes;    // rows*4   (multiple of 16)
    unsigned stage_bytes;   // tile + 2*meta
    unsigned stages;
};

template <int GAMMA_MODE, bool PER_ANCHOR>
__global__ void __launch_bounds__(LOSS_THREADS) ssd_loss_kernel(
    const float* __restrict__ logits, const float4* __restrict__ codes, const float4* __restrict__ reg_t,
    const int* __restrict__ cls_t, const int* __restrict__ matches, long long NA, int C, int rows, float gamma,
    float alpha, float one_minus_alpha, LossSmemLayout L, float* __restrict__ cls_losses, float* __restrict__ loc_losses,
    double* __restrict__ partials /*[grid][3]*/) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* full = (unsigned long long*)smem;                 // [stages]
    unsigned char* stage0 = smem + 128;
    float* s_out = (float*)(stage0 + (size_t)L.stages * L.stage_bytes);    // [2][rows] per-anchor cls losses

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long ntiles = (NA + rows - 1) / rows;
    const long long first = blockIdx.x, step = gridDim.x;

    if (tid == 0) {
        for (unsigned s = 0; s < L.stages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](long long tile, unsigned s) {
        // full tiles only (the caller handles the ragged last tile with plain loads)
        const long long n0 = tile * rows;
        unsigned char* st = stage0 + (size_t)s * L.stage_bytes;
        mbar_arrive_expect_tx(&full[s], L.tile_bytes + 2 * L.meta_bytes);
        bulk_g2s(st, logits + n0 * C, L.tile_bytes, &full[s]);
        bulk_g2s(st + L.tile_bytes, matches + n0, L.meta_bytes, &full[s]);
        bulk_g2s(st + L.tile_bytes + L.meta_bytes, cls_t + n0, L.meta_bytes, &full[s]);
    };
    auto is_full = [&](long long tile) { return (tile + 1) * rows <= NA; };

    if (tid == 0) {
        for (unsigned s = 0; s < L.stages; ++s) {
            const long long t = first + (long long)s * step;
            if (t < ntiles && is_full(t)) issue(t, s);
        }
    }

    double acc_cls = 0.0, acc_loc = 0.0, acc_cnt = 0.0;

    long long k = 0;
    for (long long tile = first; tile < ntiles; tile += step, ++k) {
        const unsigned s = (unsigned)(k % L.stages);
        const unsigned parity = (unsigned)((k / L.stages) & 1);
        unsigned char* st = stage0 + (size_t)s * L.stage_bytes;
        float* s_x = (float*)st;
        int* s_m = (int*)(st + L.tile_bytes);
        int* s_c = (int*)(st + L.tile_bytes + L.meta_bytes);
        const long long n0 = tile * rows;
        const int nrows = (int)min((long long)rows, NA - n0);

        if (is_full(tile)) {
            mbar_wait(&full[s], parity);
        } else {
            // ragged last tile: plain cooperative loads
            for (int i = tid; i < nrows * C; i += LOSS_THREADS) s_x[i] = logits[n0 * C + i];
            for (int i = tid; i < nrows; i += LOSS_THREADS) { s_m[i] = matches[n0 + i]; s_c[i] = cls_t[n0 + i]; }
            __syncthreads();
        }

        // ---- per-row work, one thread per anchor row of the tile: localisation loss + matched count
        //      (ssd.py:89,117,121) and, on the flat path, the row "patches" described below
        float tile_acc = 0.0f;
        if (tid < nrows) {
            const int m = s_m[tid];
            float l = 0.0f;
            if (m >= 0) {
                l = smooth_l1_4(codes[n0 + tid], reg_t[n0 + tid]);
                acc_loc += (double)l;
                acc_cnt += 1.0;
            }
            if (loc_losses) loc_losses[n0 + tid] = l;
            if (!PER_ANCHOR) {
                // Flat path: afterwards every element of the tile is summed as a NEGATIVE (target 0, weight 1).
                // Rows that are not like that are patched in shared memory first: the positive class of a
                // matched row is evaluated here and replaced by -inf, an ignored row (matches == -2, weight 0,
                // ssd.py:103) is replaced by -inf entirely; focal_negative(-inf) == 0 exactly.
                float* x = s_x + tid * C;
                if (m < -1) {
                    for (int c = 0; c < C; ++c) x[c] = -INFINITY;
                } else {
                    const int tc = s_c[tid] - 1;                 // one_hot(cls, C+1)[1:] -> class index, -1 = background
                    if (tc >= 0 && tc < C) {
                        tile_acc = alpha * focal_positive<GAMMA_MODE>(x[tc], gamma);
                        x[tc] = -INFINITY;
                    }
                }
            }
        }

        if (PER_ANCHOR) {
            // ---- focal loss, row mapping: warps over rows, lanes over classes (ssd.py:96-109, losses.py:34-50)
            float* out = s_out + (k & 1) * rows;
            for (int r = warp; r < nrows; r += LOSS_WARPS) {
                const int m = s_m[r];
                float row = 0.0f;
                if (m >= -1) {                                       // not_ignore (ssd.py:103)
                    const int tc = s_c[r] - 1;
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
                if (lane == 0) { out[r] = row; tile_acc += row; }
            }
        } else {
            // ---- focal loss, flat mapping: the tile is one contiguous run of nrows*C floats (a multiple of 4 for
            //      full tiles); 128-bit LDS, four independent chains per thread and iteration
            __syncthreads();                                         // patches visible
            const int n4 = (nrows * C) >> 2;
            const float4* x4 = (const float4*)s_x;
            float neg = 0.0f;
#pragma unroll 2
            for (int i = tid; i < n4; i += LOSS_THREADS) {
                const float4 v = x4[i];
                neg += (focal_negative<GAMMA_MODE>(v.x, gamma) + focal_negative<GAMMA_MODE>(v.y, gamma)) +
                       (focal_negative<GAMMA_MODE>(v.z, gamma) + focal_negative<GAMMA_MODE>(v.w, gamma));
            }
            for (int i = (n4 << 2) + tid; i < nrows * C; i += LOSS_THREADS)   // ragged last tile only
                neg += focal_negative<GAMMA_MODE>(s_x[i], gamma);
            tile_acc += one_minus_alpha * neg;
        }
        acc_cls += (double)tile_acc;

        __syncthreads();   // every warp is done with stage s (and s_out[k&1] is complete)
        if (tid == 0) {
            const long long nt = tile + (long long)L.stages * step;
            if (nt < ntiles && is_full(nt)) {
                fence_proxy_async();
                issue(nt, s);
            }
        }
        if (PER_ANCHOR) {
            const float* out = s_out + (k & 1) * rows;
            for (int i = tid; i < nrows; i += LOSS_THREADS) cls_losses[n0 + i] = out[i];
        }
    }

    // ---- CTA reduction (fixed order) -> partials[blockIdx.x]
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_cls += __shfl_xor_sync(0xffffffffu, acc_cls, o);
        acc_loc += __shfl_xor_sync(0xffffffffu, acc_loc, o);
        acc_cnt += __shfl_xor_sync(0xffffffffu, acc_cnt, o);
    }
    __shared__ double s_red[LOSS_WARPS][3];
    if (lane == 0) { s_red[warp][0] = acc_loc; s_red[warp][1] = acc_cls; s_red[warp][2] = acc_cnt; }
    __syncthreads();
    if (tid < 3) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < LOSS_WARPS; ++w) t += s_red[w][tid];
        partials[(size_t)blockIdx.x * 3 + tid] = t;
    }
}

// Sum the per-CTA partials in a fixed order -> out_sums[3] = { sum loc, sum cls, num_matches }.
__global__ void __launch_bounds__(256) loss_reduce_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
    __shared__ double s[256][3];
    double t[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += 256)
        for (int j = 0; j < 3; ++j) t[j] += partials[(size_t)i * 3 + j];
    for (int j = 0; j < 3; ++j) s[threadIdx.x][j] = t[j];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int j = 0; j < 3; ++j) s[threadIdx.x][j] += s[threadIdx.x + o][j];
        __syncthreads();
    }
    if (threadIdx.x < 3) out[threadIdx.x] = s[0][threadIdx.x];
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
                       const int* cls_t, const int* matches, long long NA, int C, int rows, double gamma, double alpha,
                       LossSmemLayout L, float* cls_losses, float* loc_losses, double* partials) {
    auto kern = ssd_loss_kernel<GM, PA>;
    SSDK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SSDK_KERNEL(ctx, SSDK_K_LOSS,
                kern<<<grid, LOSS_THREADS, smem, ctx->stream>>>(logits, (const float4*)codes, (const float4*)reg_t, cls_t, matches,
                                                                NA, C, rows, (float)gamma, (float)alpha, (float)(1.0 - alpha), L,
                                                                cls_losses, loc_losses, partials));
    return SSDK_OK;
}

extern "C" {

int ssdk_ssd_loss(ssdk_ctx* ctx, const float* logits, const float* codes, const float* reg_targets,
                  const int32_t* cls_targets, const int32_t* matches, int64_t B, int64_t A, int C, double gamma,
                  double alpha, double* out_sums, float* out_cls_losses, float* out_loc_losses) {
    SSDK_TRY(ssdk_ctx_enter(ctx));
    SSDK_REQUIRE(B >= 0 && A >= 0 && C > 0, SSDK_ERR_ARG, "ssdk_ssd_loss: bad sizes");
    SSDK_REQUIRE(out_sums != nullptr, SSDK_ERR_ARG, "ssdk_ssd_loss: out_sums is NULL");
    const long long NA = (long long)B * A;
    if (NA == 0) {
        SSDK_CHECK_CUDA(cudaMemsetAsync(out_sums, 0, 3 * sizeof(double), ctx->stream));
        return SSDK_OK;
    }
    SSDK_REQUIRE(logits && codes && reg_targets && cls_targets && matches, SSDK_ERR_ARG, "ssdk_ssd_loss: null pointer");
    SSDK_REQUIRE(aligned16(logits) && aligned16(codes) && aligned16(reg_targets) && aligned16(cls_targets) && aligned16(matches),
                 SSDK_ERR_SHAPE, "ssdk_ssd_loss: inputs must be 16-byte aligned");
    SSDK_REQUIRE(C <= 8192, SSDK_ERR_SHAPE, "ssdk_ssd_loss: num_classes %d > 8192 not supported", C);

    // tile = `rows` anchors (multiple of 8 so that rows*C*4 and rows*4 are multiples of 16), ~20 KB per stage
    int rows = (int)((20480 / (4 * (long long)C)) / 8 * 8);
    if (rows < 8) rows = 8;
    if (rows > LOSS_MAX_ROWS) rows = LOSS_MAX_ROWS;
    LossSmemLayout L;
    L.tile_bytes = (unsigned)rows * C * 4;
    L.meta_bytes = (unsigned)rows * 4;
    L.stage_bytes = L.tile_bytes + 2 * L.meta_bytes;
    L.stages = 3;
    while (L.stages > 1 && 128 + (size_t)L.stages * L.stage_bytes + 2 * rows * 4 > 200 * 1024) L.stages--;
    const size_t smem = 128 + (size_t)L.stages * L.stage_bytes + 2 * (size_t)rows * 4;
    SSDK_REQUIRE(smem <= 227 * 1024, SSDK_ERR_SHAPE, "ssdk_ssd_loss: tile does not fit shared memory (C=%d)", C);

    const long long ntiles = (NA + rows - 1) / rows;
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    long long grid = (long long)ctx->num_sms * per_sm;
    if (grid > ntiles) grid = ntiles;

    SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_partials, (size_t)grid * 3 * sizeof(double)));
    double* partials = (double*)ctx->ws_partials.p;
    const bool pa = out_cls_losses != nullptr;
    const bool g2 = (gamma == 2.0);
    int st;
    if (g2 && !pa) st = launch_loss<0, false>(ctx, (int)grid, smem, logits, codes, reg_targets, cls_targets, matches, NA, C, rows, gamma, alpha, L, out_cls_losses, out_loc_losses, partials);
    else if (g2 && pa) st = launch_loss<0, true>(ctx, (int)grid, smem, logits, codes, reg_targets, cls_targets, matches, NA, C, rows, gamma, alpha, L, out_cls_losses, out_loc_losses, partials);
    else if (!pa) st = launch_loss<1, false>(ctx, (int)grid, smem, logits, codes, reg_targets, cls_targets, matches, NA, C, rows, gamma, alpha, L, out_cls_losses, out_loc_losses, partials);
    else st = launch_loss<1, true>(ctx, (int)grid, smem, logits, codes, reg_targets, cls_targets, matches, NA, C, rows, gamma, alpha, L, out_cls_losses, out_loc_losses, partials);
    SSDK_TRY(st);
    SSDK_KERNEL(ctx, SSDK_K_LOSS_REDUCE, loss_reduce_kernel<<<1, 256, 0, ctx->stream>>>(partials, (int)grid, out_sums));
    return SSDK_OK;
}

int ssdk_loss_finalize(ssdk_ctx* ctx, const double* sums, float* out_losses) {
    SSDK_TRY(ssdk_ctx_enter(ctx));
    SSDK_REQUIRE(sums && out_losses, SSDK_ERR_ARG, "ssdk_loss_finalize: null pointer");
    loss_finalize_kernel<<<1, 1, 0, ctx->stream>>>(sums, out_losses);
    SSDK_CHECK_LAUNCH(ctx);
    return SSDK_OK;
}

}  // extern "C"
