// Anchor generation over the FPN levels.
// Replaces detector/anchor_generator.py:40-120 (AnchorGenerator.__call__) and :123-170 (tile_anchors).
// One thread per anchor; every float op rounds separately (bit-exact w.r.t. the float32 reference graph).
#include <math.h>

#include "common.cuh"

#define SSDK_MAX_LEVELS 8
#define SSDK_MAX_PER_LOC 32

struct AnchorParams {
    int num_levels, per_loc;
    float H, W;
    int start[SSDK_MAX_LEVELS + 1];   // first anchor of each level
    int gh[SSDK_MAX_LEVELS], gw[SSDK_MAX_LEVELS];
    float stride[SSDK_MAX_LEVELS];
    float scales[SSDK_MAX_LEVELS * SSDK_MAX_PER_LOC];
    float ratios[SSDK_MAX_PER_LOC];
};

__global__ void __launch_bounds__(256) anchors_kernel(const AnchorParams p, float4* __restrict__ out,
                                                      float4* __restrict__ raw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.start[p.num_levels]) return;
    int lvl = 0;
#pragma unroll
    for (int l = 1; l < SSDK_MAX_LEVELS; ++l)
        if (l < p.num_levels && i >= p.start[l]) lvl = l;
    const int j = i - p.start[lvl];
    const int k = j % p.per_loc;
    const int cell = j / p.per_loc;
    const int ix = cell % p.gw[lvl], iy = cell / p.gw[lvl];
    const float s = p.stride[lvl];
    // offsets: anchor_generator.py:92-93
    const float off_y = f_mul(0.5f, f_sub(p.H, f_mul(f_sub((float)p.gh[lvl], 1.0f), s)));
    const float off_x = f_mul(0.5f, f_sub(p.W, f_mul(f_sub((float)p.gw[lvl], 1.0f), s)));
    // sizes: anchor_generator.py:144-146
    const float rs = __fsqrt_rn(p.ratios[k]);
    const float scale = p.scales[lvl * p.per_loc + k];
    const float hgt = f_div(scale, rs), wid = f_mul(scale, rs);
    // centres: anchor_generator.py:151-152
    const float cy = f_add(f_mul((float)iy, s), off_y), cx = f_add(f_mul((float)ix, s), off_x);
    const float hh = f_mul(0.5f, hgt), hw = f_mul(0.5f, wid);
    const float4 b = make_float4(f_sub(cy, hh), f_sub(cx, hw), f_add(cy, hh), f_add(cx, hw));   // :166
    if (raw) raw[i] = b;
    out[i] = make_float4(f_div(b.x, p.H), f_div(b.y, p.W), f_div(b.z, p.H), f_div(b.w, p.W));  // :110-114
}

static void level_grid(int H, int W, int stride, int* gh, int* gw) {
    // float32 division + ceil, as tf.to_int32(tf.ceil(tf.to_float(H) / stride))  (:59-60)
    *gh = (int)ceilf((float)H / (float)stride);
    *gw = (int)ceilf((float)W / (float)stride);
}

extern "C" int ssdk_num_anchors(int H, int W, const int* strides, int L, int per_loc, int64_t* out_total,
                                int32_t* out_per_level) {
    SSDK_REQUIRE(H > 0 && W > 0 && strides && L > 0 && per_loc > 0, SSDK_ERR_ARG, "ssdk_num_anchors: bad arguments");
    int64_t total = 0;
    for (int l = 0; l < L; ++l) {
        SSDK_REQUIRE(strides[l] > 0, SSDK_ERR_ARG, "stride %d must be positive", l);
        int gh, gw;
        level_grid(H, W, strides[l], &gh, &gw);
        const int64_t n = (int64_t)gh * gw * per_loc;
        if (out_per_level) out_per_level[l] = (int32_t)n;
        total += n;
    }
    if (out_total) *out_total = total;
    return SSDK_OK;
}

extern "C" int ssdk_anchors(ssdk_ctx* ctx, int H, int W, const int* strides, const float* scales,
                            const float* ratios, int L, int per_loc, float* out, float* raw) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(H > 0 && W > 0 && strides && scales && ratios && out, SSDK_ERR_ARG, "ssdk_anchors: bad arguments");
    SSDK_REQUIRE(L > 0 && L <= SSDK_MAX_LEVELS, SSDK_ERR_SHAPE, "num_levels %d not in [1,%d]", L, SSDK_MAX_LEVELS);
    SSDK_REQUIRE(per_loc > 0 && per_loc <= SSDK_MAX_PER_LOC, SSDK_ERR_SHAPE, "anchors_per_location %d not in [1,%d]",
                 per_loc, SSDK_MAX_PER_LOC);
    AnchorParams p;
    p.num_levels = L; p.per_loc = per_loc; p.H = (float)H; p.W = (float)W;
    int64_t total = 0;
    for (int l = 0; l < L; ++l) {
        SSDK_REQUIRE(strides[l] > 0, SSDK_ERR_ARG, "stride %d must be positive", l);
        level_grid(H, W, strides[l], &p.gh[l], &p.gw[l]);
        p.stride[l] = (float)strides[l];
        p.start[l] = (int)total;
        total += (int64_t)p.gh[l] * p.gw[l] * per_loc;
        for (int k = 0; k < per_loc; ++k) p.scales[l * per_loc + k] = scales[l * per_loc + k];
    }
    SSDK_REQUIRE(total < (1ll << 31), SSDK_ERR_SHAPE, "too many anchors");
    for (int l = L; l <= SSDK_MAX_LEVELS; ++l) p.start[l] = (int)total;
    for (int k = 0; k < per_loc; ++k) p.ratios[k] = ratios[k];
    SSDK_KERNEL(ctx, SSDK_K_ANCHORS,
                anchors_kernel<<<ceil_div_i(total, 256), 256, 0, ctx->stream>>>(p, (float4*)out, (float4*)raw));
    return SSDK_OK;
}
