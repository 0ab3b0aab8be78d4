// Leaf box math exposed one-to-one with detector/utils/box_utils.py (area :53, intersection :30,
// iou :14, encode :80, decode :114, batch_decode :145) and detector/losses.py (localization_loss :4,
// focal_loss :22 with dense one-hot targets).  These are the reference's public helpers; the fused
// training / inference kernels live in matcher.cu, loss.cu and postprocess.cu.
#include "common.cuh"

__global__ void __launch_bounds__(256) area_kernel(const float4* __restrict__ b, long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = box_area(b[i]);
}

template <bool IOU>
__global__ void __launch_bounds__(256) pairwise_kernel(const float4* __restrict__ b1, long long n,
                                                       const float4* __restrict__ b2, long long m,
                                                       float* __restrict__ out) {
    // grid.y strides over rows of boxes1; threads over boxes2 -> coalesced [n,m] writes
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const float4 bj = b2[j];
    const float aj = box_area(bj);
    for (long long i = blockIdx.y; i < n; i += gridDim.y) {
        const float4 bi = b1[i];
        out[i * m + j] = IOU ? box_iou_areas(bi, box_area(bi), bj, aj) : box_intersection(bi, bj);
    }
}

template <int MODE>  // 0 encode, 1 decode, 2 batch_decode (anchor index = i % A, clip)
__global__ void __launch_bounds__(256) coder_kernel(const float4* __restrict__ x, const float4* __restrict__ anchors,
                                                    long long n, long long A, float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (MODE == 0) out[i] = box_encode(x[i], anchors[i]);
    if (MODE == 1) out[i] = box_decode(x[i], anchors[i]);
    if (MODE == 2) out[i] = box_clip01(box_decode(x[i], anchors[i % A]));
}

// localization_loss: losses.py:16-19
__global__ void __launch_bounds__(256) loc_loss_kernel(const float4* __restrict__ p, const float4* __restrict__ t,
                                                       const float* __restrict__ w, long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = p[i], b = t[i];
    const float d[4] = {fabsf(f_sub(a.x, b.x)), fabsf(f_sub(a.y, b.y)), fabsf(f_sub(a.z, b.z)), fabsf(f_sub(a.w, b.w))};
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) s = f_add(s, d[k] < 1.0f ? f_mul(0.5f, f_mul(d[k], d[k])) : f_sub(d[k], 0.5f));
    out[i] = f_mul(w[i], s);
}

// focal_loss with dense float targets: losses.py:34-50, op for op (accurate libm functions; this is
// the reference-API helper, the bandwidth-optimised path is ssdk_ssd_loss).
__global__ void __launch_bounds__(256) focal_dense_kernel(const float* __restrict__ x, const float* __restrict__ z,
                                                          const float* __restrict__ w, long long n, int C, float gamma,
                                                          float alpha, float one_minus_alpha, float* __restrict__ out) {
    // one warp per anchor row
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    float s = 0.0f;
    for (int c = lane; c < C; c += 32) {
        const float xv = x[row * C + c], zv = z[row * C + c];
        const bool pos = zv == 1.0f;
        const float nlpt = f_add(f_sub(fmaxf(xv, 0.0f), f_mul(xv, zv)), log1pf(expf(-fabsf(xv))));
        const float p = f_div(1.0f, f_add(1.0f, expf(-xv)));
        const float pt = pos ? p : f_sub(1.0f, p);
        const float mod = powf(f_sub(1.0f, pt), gamma);
        const float wl = pos ? f_mul(alpha, nlpt) : f_mul(one_minus_alpha, nlpt);
        s += f_mul(mod, wl);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = f_mul(w[row], s);
}

extern "C" {

int ssdk_area(ssdk_ctx* ctx, const float* boxes, int64_t n, float* out) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(n >= 0 && (n == 0 || (boxes && out)), SSDK_ERR_ARG, "ssdk_area: bad arguments");
    if (n == 0) return SSDK_OK;
    area_kernel<<<ceil_div_i(n, 256), 256, 0, ctx->stream>>>((const float4*)boxes, n, out);
    SSDK_CHECK_LAUNCH(ctx);
    return SSDK_OK;
}

static int pairwise(ssdk_ctx* ctx, bool is_iou, const float* b1, int64_t n, const float* b2, int64_t m, float* out) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(n >= 0 && m >= 0, SSDK_ERR_ARG, "pairwise: negative size");
    if (n == 0 || m == 0) return SSDK_OK;
    SSDK_REQUIRE(b1 && b2 && out, SSDK_ERR_ARG, "pairwise: null pointer");
    dim3 grid(ceil_div_i(m, 256), (unsigned)(n < 1024 ? n : 1024));
    if (is_iou) pairwise_kernel<true><<<grid, 256, 0, ctx->stream>>>((const float4*)b1, n, (const float4*)b2, m, out);
    else pairwise_kernel<false><<<grid, 256, 0, ctx->stream>>>((const float4*)b1, n, (const float4*)b2, m, out);
    SSDK_CHECK_LAUNCH(ctx);
    return SSDK_OK;
}

int ssdk_intersection(ssdk_ctx* ctx, const float* b1, int64_t n, const float* b2, int64_t m, float* out) {
    return pairwise(ctx, false, b1, n, b2, m, out);
}

int ssdk_iou(ssdk_ctx* ctx, const float* b1, int64_t n, const float* b2, int64_t m, float* out) {
    return pairwise(ctx, true, b1, n, b2, m, out);
}

static int coder(ssdk_ctx* ctx, int mode, const float* x, const float* anchors, int64_t n, int64_t A, float* out) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(n >= 0, SSDK_ERR_ARG, "coder: negative size");
    if (n == 0) return SSDK_OK;
    SSDK_REQUIRE(x && anchors && out, SSDK_ERR_ARG, "coder: null pointer");
    const int grid = ceil_div_i(n, 256);
    if (mode == 0) coder_kernel<0><<<grid, 256, 0, ctx->stream>>>((const float4*)x, (const float4*)anchors, n, A, (float4*)out);
    if (mode == 1) coder_kernel<1><<<grid, 256, 0, ctx->stream>>>((const float4*)x, (const float4*)anchors, n, A, (float4*)out);
    if (mode == 2) coder_kernel<2><<<grid, 256, 0, ctx->stream>>>((const float4*)x, (const float4*)anchors, n, A, (float4*)out);
    SSDK_CHECK_LAUNCH(ctx);
    return SSDK_OK;
}

int ssdk_encode(ssdk_ctx* ctx, const float* boxes, const float* anchors, int64_t n, float* out) {
    return coder(ctx, 0, boxes, anchors, n, n, out);
}
int ssdk_decode(ssdk_ctx* ctx, const float* codes, const float* anchors, int64_t n, float* out) {
    return coder(ctx, 1, codes, anchors, n, n, out);
}
int ssdk_batch_decode(ssdk_ctx* ctx, const float* codes, const float* anchors, int64_t B, int64_t A, float* out) {
    SSDK_REQUIRE(B >= 0 && A >= 0, SSDK_ERR_ARG, "ssdk_batch_decode: negative size");
    return coder(ctx, 2, codes, anchors, B * A, A, out);
}

int ssdk_localization_loss(ssdk_ctx* ctx, const float* p, const float* t, const float* w, int64_t B, int64_t A, float* out) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(B >= 0 && A >= 0, SSDK_ERR_ARG, "ssdk_localization_loss: negative size");
    const int64_t n = B * A;
    if (n == 0) return SSDK_OK;
    SSDK_REQUIRE(p && t && w && out, SSDK_ERR_ARG, "ssdk_localization_loss: null pointer");
    loc_loss_kernel<<<ceil_div_i(n, 256), 256, 0, ctx->stream>>>((const float4*)p, (const float4*)t, w, n, out);
    SSDK_CHECK_LAUNCH(ctx);
    return SSDK_OK;
}

int ssdk_focal_loss(ssdk_ctx* ctx, const float* logits, const float* targets, const float* weights, int64_t B,
                    int64_t A, int C, double gamma, double alpha, float* out) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(B >= 0 && A >= 0 && C > 0, SSDK_ERR_ARG, "ssdk_focal_loss: bad sizes");
    const int64_t n = B * A;
    if (n == 0) return SSDK_OK;
    SSDK_REQUIRE(logits && targets && weights && out, SSDK_ERR_ARG, "ssdk_focal_loss: null pointer");
    focal_dense_kernel<<<ceil_div_i(n * 32, 256), 256, 0, ctx->stream>>>(logits, targets, weights, n, C, (float)gamma,
                                                                        (float)alpha, (float)(1.0 - alpha), out);
    SSDK_CHECK_LAUNCH(ctx);
    return SSDK_OK;
}

}  // extern "C"
