// The flat pass over the class logits, shared by head_flat_kernel (head.cu) and the fused training-step kernel
// (train_step.cu): every level's class tensor is streamed as one flat array of 16 KB chunks (FLAT_THREADS x FLAT_U 128-bit
// no-allocate loads), each element is summed as if it were a negative (detector/losses.py:22-50 with targets == 0), and the
// few matched / ignored anchors are corrected elsewhere.  See head.cu for why this is layout independent.
#pragma once
#include "focal_math.cuh"

#define FLAT_THREADS 256
#ifndef FLAT_U
#define FLAT_U 4                                        // float4 per thread and chunk (2: 0.166 ms, 8: 0.160 ms, 4: 0.148 ms forward)
#endif
#define FLAT_CHUNK4 (FLAT_THREADS * FLAT_U)             // float4 per chunk (16 KB)

struct FlatSegs {
    int n;                                              // class-tensor segments [0, n); WITH_GRAD: zero-fill segments [n, 2n)
    int nseg;                                           // n or 2n
    const float* src[SSDK_MAX_LEVELS];
    float* dst[2 * SSDK_MAX_LEVELS];                    // WITH_GRAD: [0,n) class gradients (same flat indexing), [n,2n) box gradients
    long long count[2 * SSDK_MAX_LEVELS];               // floats per segment (B * n*C * h*w, resp. B * n*4 * h*w)
    long long chunk0[2 * SSDK_MAX_LEVELS + 1];          // prefix sums of the per-segment chunk counts
};

struct FlatChunk {
    float4 v[FLAT_U];
    long long i4;        // index of v[0] (in float4) for this thread
    int lvl;
};

// Loads chunk g (global chunk index over all segments) for this thread; `cursor` is the segment of the previously loaded
// chunk (chunks are visited in ascending order by a CTA).
__device__ __forceinline__ void flat_load(const FlatSegs& S, FlatChunk& ck, long long g, int& cursor, int tid) {
    while (g >= S.chunk0[cursor + 1]) ++cursor;
    ck.lvl = cursor;
    ck.i4 = (g - S.chunk0[cursor]) * FLAT_CHUNK4 + tid;
    const long long n4 = S.count[cursor] >> 2;
    const float4* src4 = (const float4*)S.src[cursor];
    const float4 ninf4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int u = 0; u < FLAT_U; ++u) {
        const long long i = ck.i4 + u * FLAT_THREADS;
        ck.v[u] = (i < n4) ? ld_stream_f4(src4 + i) : ninf4;
    }
}

// Sum over FLAT_U float4 of the negative-class focal term (without the (1-alpha) factor), forward only.
template <int GAMMA_MODE>
__device__ __forceinline__ float flat_math(const float4 (&v)[FLAT_U], float gamma) {
    float s;
    bool general = (GAMMA_MODE != 0);
    if (GAMMA_MODE == 0) {
        f32x2 a4[4] = {0ull, 0ull, 0ull, 0ull};
        float mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < FLAT_U; u += 2) focal_half8(v[u], v[u + 1], a4, mx);
        float s0, s1;
        unpack2(add2(add2(a4[0], a4[1]), add2(a4[2], a4[3])), s0, s1);
        s = s0 + s1;
        general = mx > FOCAL_HALF_MAX_X;                    // some logit > -ln 2: redo these 16 with the general form
    }
    if (general) {
        f32x2 a01 = 0ull, a23 = 0ull;
#pragma unroll
        for (int u = 0; u < FLAT_U; ++u) {
            a01 = focal_negative2<GAMMA_MODE>(v[u].x, v[u].y, gamma, a01);
            a23 = focal_negative2<GAMMA_MODE>(v[u].z, v[u].w, gamma, a23);
        }
        float s0, s1;
        unpack2(add2(a01, a23), s0, s1);
        s = s0 + s1;
    }
    return s;
}

// The same for a loaded chunk, plus the (< 4) floats of a level beyond its last float4 (handled with the level's last chunk).
template <int GAMMA_MODE>
__device__ __forceinline__ float flat_value(const FlatSegs& S, const FlatChunk& ck, long long g, float gamma, int tid) {
    float s = flat_math<GAMMA_MODE>(ck.v, gamma);
    if (g + 1 == S.chunk0[ck.lvl + 1]) {
        const long long n = S.count[ck.lvl];
        const int tail = (int)(n & 3);
        if (tid < tail) s += focal_negative<GAMMA_MODE>(S.src[ck.lvl][(n & ~3ll) + tid], gamma);
    }
    return s;
}

// Accumulators of the flat pass.  FlatAccDouble: one double per thread (the order of the additions is part of the result, so the
// chunk -> thread assignment must be static for reproducible sums).  FlatAccFixed: 2^-32 fixed point in 64 bits -- integer addition
// is associative, so the sum does not depend on which CTA happened to take which chunk: this is what lets the fused training
// step hand out chunks DYNAMICALLY and still return bit-identical sums run after run.  Every per-thread chunk sum (a float,
// 16 elements) is rounded to a multiple of 2^-32 (1.2e-10 absolute; the sums are ~1e-3 and larger, and 1e4 in total); chunk sums
// of 2^30 and more (never with finite logits of sane size) and NaN go to a double on the side.
struct FlatAccDouble {
    double v = 0.0;
    __device__ __forceinline__ void add(float s) { v += (double)s; }
};
struct FlatAccFixed {
    long long fx = 0;
    double big = 0.0;
    __device__ __forceinline__ void add(float s) {
        if (fabsf(s) < 1073741824.0f) fx += __float2ll_rn(s * 4294967296.0f);
        else big += (double)s;
    }
};

// This thread's share of the chunks g, g + stride, g + 2 stride, ... < g_end of the class-tensor segments (forward only), in
// ascending order.  Segment by segment: the FULL chunks of a segment are read through a pointer that advances by a constant,
// without bounds tests or segment lookups (in the generic form above those cost more instructions per chunk than the focal
// arithmetic itself: ncu source view of the round-2 training-step kernel, profiles/r2b_ncu_summary.txt); only a segment's last,
// partial chunk takes the bounded path.  Same per-chunk float sums and the same order of additions as chunk-by-chunk
// flat_load + flat_value.
template <int GAMMA_MODE, class Acc>
__device__ __forceinline__ void flat_sum_range(const FlatSegs& S, Acc& acc, long long g, long long g_end, long long stride, float gamma, int tid) {
    const long long pstep = stride * FLAT_CHUNK4;
#pragma unroll 1
    for (int sg = 0; sg < S.n && g < g_end; ++sg) {
        const long long c1 = S.chunk0[sg + 1];
        if (g >= c1) continue;
        const long long c0 = S.chunk0[sg];
        long long full_end = c0 + (S.count[sg] >> 2) / FLAT_CHUNK4;          // chunks [c0, full_end) hold FLAT_CHUNK4 float4 each
        if (full_end > g_end) full_end = g_end;
        const float4* p = (const float4*)S.src[sg] + (g - c0) * FLAT_CHUNK4 + tid;
        // a 32-bit trip count instead of 64-bit chunk indices in the loop (the kernel lives on 40 registers)
        const int n_iter = g < full_end ? (int)((full_end - g + stride - 1) / stride) : 0;
#pragma unroll 1
        for (int it = 0; it < n_iter; ++it, p += pstep) {
            float4 v[FLAT_U];
#pragma unroll
            for (int u = 0; u < FLAT_U; ++u) v[u] = ld_stream_f4(p + u * FLAT_THREADS);
            acc.add(flat_math<GAMMA_MODE>(v, gamma));
        }
        g += (long long)n_iter * stride;
        if (g < c1 && g < g_end) {                                            // the segment's partial last chunk
            FlatChunk ck;
            int cursor = sg;
            flat_load(S, ck, g, cursor, tid);
            acc.add(flat_value<GAMMA_MODE>(S, ck, g, gamma, tid));
            g += stride;
        }
    }
}
template <int GAMMA_MODE>
__device__ __forceinline__ double flat_sum_range(const FlatSegs& S, long long g, long long g_end, long long stride, float gamma, int tid) {
    FlatAccDouble acc;
    flat_sum_range<GAMMA_MODE>(S, acc, g, g_end, stride, gamma, tid);
    return acc.v;
}
