// Per-element arithmetic of the SSD losses (detector/losses.py:4-50) and of their derivatives, shared by the streaming
// kernels of loss.cu, loss_backward.cu and head.cu.
#pragma once
#include "stream.cuh"

// ---------------------------------------------------------------------------------------------- per-element math
// Negative-class term without the (1-alpha) factor, reference rounding: (1 - p_t)^gamma * nlpt with targets == 0
// (losses.py:36-41: nlpt = max(x,0) + log1p(exp(-|x|)), p_t = 1 - p, modulating factor (1 - (1 - p))^gamma).
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

// Two negatives at once, packed: returns acc + mod * nlpt for both lanes of the pair.
// 1 - p_t is formed as  x >= 0 ? r : 1 - r  with r = 1/(1+e) = sigmoid(|x|): for x < 0 this is the same quantity
// as the reference's 1 - fl(1 - p) (r is fl(1 - p) to within an ulp), for x >= 0 the reference's 1 - (1 - p)
// is exact (Sterbenz) and equals p = r.
template <int GAMMA_MODE>
__device__ __forceinline__ f32x2 focal_negative2(float x0, float x1, float gamma, f32x2 acc) {
    const float e0 = ex2_approx(-fabsf(x0) * 1.4426950408889634f);
    const float e1 = ex2_approx(-fabsf(x1) * 1.4426950408889634f);
    const f32x2 e = pack2(e0, e1);
    f32x2 p = fma2(splat2(L1P_C7), e, splat2(L1P_C6));
    p = fma2(p, e, splat2(L1P_C5)); p = fma2(p, e, splat2(L1P_C4)); p = fma2(p, e, splat2(L1P_C3));
    p = fma2(p, e, splat2(L1P_C2)); p = fma2(p, e, splat2(L1P_C1)); p = fma2(p, e, splat2(L1P_C0));
    const f32x2 nlpt = fma2(p, e, pack2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));    // max(x,0) + log1p(exp(-|x|))
    float d0, d1;
    unpack2(add2(e, splat2(1.0f)), d0, d1);
    const float r0 = rcp_approx(d0), r1 = rcp_approx(d1);
    float c0, c1;
    unpack2(fma2(pack2(r0, r1), splat2(-1.0f), splat2(1.0f)), c0, c1);         // 1 - r
    const float q0 = (x0 >= 0.0f) ? r0 : c0, q1 = (x1 >= 0.0f) ? r1 : c1;
    if (GAMMA_MODE == 0) {
        const f32x2 q = pack2(q0, q1);
        return fma2(mul2(q, q), nlpt, acc);
    }
    return fma2(pack2(powf(q0, gamma), powf(q1, gamma)), nlpt, acc);
}


// Fast path for negatives with x <= 0 (the overwhelming majority: background logits), gamma == 2:
//     (1 - p_t)^2 * nlpt = sigmoid(x)^2 * softplus(x) = e^3 * g(e),  e = exp(x) in [0,1],  g(e) = log1p(e) / (e (1+e)^2)
// with g a degree-9 minimax polynomial in t = 2e - 1 (relative error 1.1e-6; the centred variable keeps the float32
// Horner evaluation well conditioned).  One MUFU (EX2) per element instead of two and ~30 % fewer issue slots than
// the general form; profiles/microbench/focal_math_bench.cu measures 1.63 vs 1.18 elements/clk/SMSP, against the
// 1.41 that HBM can deliver.  Eight elements (four packed pairs) are processed together with their Horner steps
// interleaved, because the kernel is issue/latency-bound, not FMA-pipe-bound.  -inf (patched elements) gives e = 0 and
// contributes exactly 0.  `allneg` collects the AND of the sign bits: if any element is >= +0 the caller discards the
// result and re-sums the tile with the general form.
#define FG0 3.604137897e-01f
#define FG1 -3.043911755e-01f
#define FG2 1.775974035e-01f
#define FG3 -8.836640418e-02f
#define FG4 4.034566879e-02f
#define FG5 -1.723237708e-02f
#define FG6 6.659520790e-03f
#define FG7 -2.797316527e-03f
#define FG8 1.626353362e-03f
#define FG9 -5.688594538e-04f

__device__ __forceinline__ void focal_fast8(const float4 v, const float4 w, f32x2 (&acc)[4], unsigned& allneg) {
    const float x[8] = {v.x, v.y, v.z, v.w, w.x, w.y, w.z, w.w};
    const float G[10] = {FG0, FG1, FG2, FG3, FG4, FG5, FG6, FG7, FG8, FG9};
    f32x2 e[4], t[4], p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        allneg &= __float_as_uint(x[2 * j]) & __float_as_uint(x[2 * j + 1]);
        e[j] = pack2(ex2_approx(x[2 * j] * 1.4426950408889634f), ex2_approx(x[2 * j + 1] * 1.4426950408889634f));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        t[j] = fma2(e[j], splat2(2.0f), splat2(-1.0f));
        p[j] = fma2(splat2(G[9]), t[j], splat2(G[8]));
    }
#pragma unroll
    for (int k = 7; k >= 0; --k)
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = fma2(p[j], t[j], splat2(G[k]));
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = fma2(mul2(mul2(e[j], e[j]), e[j]), p[j], acc[j]);
}

// Fast path of the FLAT passes (head.cu, train_step.cu), gamma == 2, for negatives with x <= -ln 2, i.e. e = exp(x) <= 1/2:
//     sigmoid(x)^2 * softplus(x) = e^3 * h(e),   h(e) = log1p(e) / (e (1+e)^2)  as a degree-FOCAL_HALF_DEG polynomial in e itself
// (h(0) = 1 exactly; minimax relative error on [0, 1/2]: 4.4e-6 for degree 6, 8.8e-7 for degree 7, float32 Horner included;
// loss-weighted error on prior-bias logits N(-4.6, 1): 2e-7).  Against focal_fast8 (degree 9 in t = 2e - 1 on [0, 1]) this
// removes four of the fourteen FMA-pipe operations per element: the flat pass moves the same stream as the score scan of
// postprocess.cu, and with 59 % of the FMA pipe busy (ncu, round 1) it reached 0.85 of the HBM roofline where the scan reaches
// 0.97.  `mx` collects the maximum logit seen (FMNMX, ALU pipe): a caller whose elements are not all <= -ln 2 discards the
// result and re-sums those elements with the general form.  -inf gives e = 0 and contributes exactly 0.
#ifndef FOCAL_HALF_DEG
#define FOCAL_HALF_DEG 6
#endif
#define FOCAL_HALF_MAX_X (-0.6931472f)
#if FOCAL_HALF_DEG == 6
#define FH_COEFFS {1.000000000e+00f, -2.499542475e+00f, 4.315966606e+00f, -6.187100410e+00f, 7.222608566e+00f, -5.857295513e+00f, 2.317409754e+00f}
#elif FOCAL_HALF_DEG == 7
#define FH_COEFFS {1.000000000e+00f, -2.499929428e+00f, 4.329836845e+00f, -6.355987549e+00f, 8.180599213e+00f, -8.630109787e+00f, 6.292103767e+00f, -2.239150286e+00f}
#else
#error "FOCAL_HALF_DEG must be 6 or 7"
#endif

__device__ __forceinline__ void focal_half8(const float4 v, const float4 w, f32x2 (&acc)[4], float& mx) {
    const float H[FOCAL_HALF_DEG + 1] = FH_COEFFS;
    mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    mx = fmaxf(mx, fmaxf(fmaxf(w.x, w.y), fmaxf(w.z, w.w)));
    const f32x2 xs[4] = {pack2(v.x, v.y), pack2(v.z, v.w), pack2(w.x, w.y), pack2(w.z, w.w)};
    f32x2 e[4], p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float y0, y1;
        unpack2(mul2(xs[j], splat2(1.4426950408889634f)), y0, y1);
        e[j] = pack2(ex2_approx(y0), ex2_approx(y1));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = fma2(splat2(H[FOCAL_HALF_DEG]), e[j], splat2(H[FOCAL_HALF_DEG - 1]));
#pragma unroll
    for (int k = FOCAL_HALF_DEG - 2; k >= 0; --k)
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = fma2(p[j], e[j], splat2(H[k]));
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = fma2(mul2(e[j], e[j]), mul2(e[j], p[j]), acc[j]);
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

// ---------------------------------------------------------------------------------------------- derivatives
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

// value and d/dx of the negative-class term (both without the (1-alpha) factor): shares e, r, softplus
template <int GAMMA_MODE>
__device__ __forceinline__ float focal_negative_both(float x, float gamma, float& value) {
    const float e = ex2_approx(-fabsf(x) * 1.4426950408889634f);
    const float r = rcp_approx(1.0f + e);
    const float er = e * r;
    const float p = (x >= 0.0f) ? r : er;
    const float q = (x >= 0.0f) ? er : r;
    const float sp = fmaxf(x, 0.0f) + log1p_unit(e);
    const float mod = (GAMMA_MODE == 0) ? p * p : powf(p, gamma);
    value = mod * sp;
    return mod * fmaf(gamma * q, sp, p);
}

// value of the positive-class term without the alpha factor: losses.py:36-41 with targets == 1
template <int GAMMA_MODE>
__device__ __forceinline__ float focal_positive_value(float x, float gamma) {
    const float nlpt = (fmaxf(x, 0.0f) - x) + log1pf(expf(-fabsf(x)));
    const float q = 1.0f / (1.0f + expf(x));
    const float mod = (GAMMA_MODE == 0) ? q * q : powf(q, gamma);
    return mod * nlpt;
}

__device__ __forceinline__ float smooth_l1_value(float p, float t) {
    const float d = fabsf(p - t);
    return d < 1.0f ? 0.5f * d * d : d - 0.5f;
}

__device__ __forceinline__ float smooth_l1_grad(float p, float t) {
    const float d = p - t;
    return (fabsf(d) < 1.0f) ? d : ((d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f));
}
