// Micro-benchmark: per-element throughput of candidate focal-loss (negative class) formulations on sm_100a, data
// resident in shared memory (one 8-row x 90-class tile per warp, as in ssd_loss_kernel), no global traffic.
// Answers: is the loss kernel's math below the HBM budget (1.3 elements/clk/SMSP at 6.5 TB/s), and which mix of
// scalar FFMA / packed FFMA2 / MUFU gets the most elements per clock.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o focal_math_bench focal_math_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 splat2(float c) { return pack2(c, c); }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

#define L1P_C7 -8.539209655e-03f
#define L1P_C6 4.408963566e-02f
#define L1P_C5 -1.076817184e-01f
#define L1P_C4 1.774524181e-01f
#define L1P_C3 -2.449546295e-01f
#define L1P_C2 3.327547979e-01f
#define L1P_C1 -4.999740540e-01f
#define L1P_C0 9.999998057e-01f

// g(e) = log1p(e) / (e (1+e)^2) on [0,1] as a degree-10 polynomial in t = 2e - 1 (relative error 3.6e-7 in float32)
#define G0 3.604134657e-01f
#define G1 -3.043935497e-01f
#define G2 1.776154122e-01f
#define G3 -8.833344504e-02f
#define G4 4.019098684e-02f
#define G5 -1.735448921e-02f
#define G6 7.114912402e-03f
#define G7 -2.631109393e-03f
#define G8 1.086219742e-03f
#define G9 -6.438818984e-04f
#define G10 2.223111762e-04f

#define LOG2E 1.4426950408889634f

// ---- variant 0: current kernel math, packed
__device__ __forceinline__ f32x2 cur2(float x0, float x1, f32x2 acc) {
    const float e0 = ex2_approx(-fabsf(x0) * LOG2E), e1 = ex2_approx(-fabsf(x1) * LOG2E);
    const f32x2 e = pack2(e0, e1);
    f32x2 p = fma2(splat2(L1P_C7), e, splat2(L1P_C6));
    p = fma2(p, e, splat2(L1P_C5)); p = fma2(p, e, splat2(L1P_C4)); p = fma2(p, e, splat2(L1P_C3));
    p = fma2(p, e, splat2(L1P_C2)); p = fma2(p, e, splat2(L1P_C1)); p = fma2(p, e, splat2(L1P_C0));
    const f32x2 nlpt = fma2(p, e, pack2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
    float d0, d1;
    unpack2(add2(e, splat2(1.0f)), d0, d1);
    const float r0 = rcp_approx(d0), r1 = rcp_approx(d1);
    float c0, c1;
    unpack2(fma2(pack2(r0, r1), splat2(-1.0f), splat2(1.0f)), c0, c1);
    const float q0 = (x0 >= 0.0f) ? r0 : c0, q1 = (x1 >= 0.0f) ? r1 : c1;
    const f32x2 q = pack2(q0, q1);
    return fma2(mul2(q, q), nlpt, acc);
}
// ---- variant 1: current math, scalar
__device__ __forceinline__ float cur1(float x, float acc) {
    const float e = ex2_approx(-fabsf(x) * LOG2E);
    float p = L1P_C7;
    p = fmaf(p, e, L1P_C6); p = fmaf(p, e, L1P_C5); p = fmaf(p, e, L1P_C4); p = fmaf(p, e, L1P_C3);
    p = fmaf(p, e, L1P_C2); p = fmaf(p, e, L1P_C1); p = fmaf(p, e, L1P_C0);
    const float nlpt = fmaf(p, e, fmaxf(x, 0.0f));
    const float r = rcp_approx(1.0f + e);
    const float q = (x >= 0.0f) ? r : 1.0f - r;
    return fmaf(q * q, nlpt, acc);
}
// ---- variant 2: x < 0 fast path, f = e^3 g(e), packed
__device__ __forceinline__ f32x2 poly2(float x0, float x1, f32x2 acc) {
    const float e0 = ex2_approx(x0 * LOG2E), e1 = ex2_approx(x1 * LOG2E);
    const f32x2 e = pack2(e0, e1);
    const f32x2 t = fma2(e, splat2(2.0f), splat2(-1.0f));
    f32x2 p = fma2(splat2(G10), t, splat2(G9));
    p = fma2(p, t, splat2(G8)); p = fma2(p, t, splat2(G7)); p = fma2(p, t, splat2(G6)); p = fma2(p, t, splat2(G5));
    p = fma2(p, t, splat2(G4)); p = fma2(p, t, splat2(G3)); p = fma2(p, t, splat2(G2)); p = fma2(p, t, splat2(G1));
    p = fma2(p, t, splat2(G0));
    const f32x2 e2 = mul2(e, e);
    return fma2(mul2(e2, e), p, acc);
}
// ---- variant 3: same, scalar
__device__ __forceinline__ float poly1(float x, float acc) {
    const float e = ex2_approx(x * LOG2E);
    const float t = fmaf(e, 2.0f, -1.0f);
    float p = G10;
    p = fmaf(p, t, G9); p = fmaf(p, t, G8); p = fmaf(p, t, G7); p = fmaf(p, t, G6); p = fmaf(p, t, G5);
    p = fmaf(p, t, G4); p = fmaf(p, t, G3); p = fmaf(p, t, G2); p = fmaf(p, t, G1); p = fmaf(p, t, G0);
    return fmaf(e * e * e, p, acc);
}


// ---- variants 8-10: NP packed pair-chains interleaved by hand (Horner steps issued round-robin over the chains)
template <int NP, int DEG, bool CLAMP>
__device__ __forceinline__ void poly_pairs(const float* x /*[2*NP]*/, f32x2* acc /*[NP]*/, unsigned& allneg) {
    const float G[11] = {G0, G1, G2, G3, G4, G5, G6, G7, G8, G9, G10};
    f32x2 e[NP], t[NP], p[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        float x0 = x[2 * j], x1 = x[2 * j + 1];
        if (CLAMP) {
            allneg &= __float_as_uint(x0) & __float_as_uint(x1);
            x0 = fminf(x0, 0.0f); x1 = fminf(x1, 0.0f);
        }
        e[j] = pack2(ex2_approx(x0 * LOG2E), ex2_approx(x1 * LOG2E));
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) { t[j] = fma2(e[j], splat2(2.0f), splat2(-1.0f)); p[j] = fma2(splat2(G[DEG]), t[j], splat2(G[DEG - 1])); }
#pragma unroll
    for (int k = DEG - 2; k >= 0; --k)
#pragma unroll
        for (int j = 0; j < NP; ++j) p[j] = fma2(p[j], t[j], splat2(G[k]));
#pragma unroll
    for (int j = 0; j < NP; ++j) { const f32x2 e2 = mul2(e[j], e[j]); acc[j] = fma2(mul2(e2, e[j]), p[j], acc[j]); }
}
template <int NS, int DEG>
__device__ __forceinline__ void poly_scalars(const float* x, float* acc) {
    const float G[11] = {G0, G1, G2, G3, G4, G5, G6, G7, G8, G9, G10};
    float e[NS], t[NS], p[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) e[j] = ex2_approx(x[j] * LOG2E);
#pragma unroll
    for (int j = 0; j < NS; ++j) { t[j] = fmaf(e[j], 2.0f, -1.0f); p[j] = fmaf(G[DEG], t[j], G[DEG - 1]); }
#pragma unroll
    for (int k = DEG - 2; k >= 0; --k)
#pragma unroll
        for (int j = 0; j < NS; ++j) p[j] = fmaf(p[j], t[j], G[k]);
#pragma unroll
    for (int j = 0; j < NS; ++j) acc[j] = fmaf(e[j] * e[j] * e[j], p[j], acc[j]);
}

template <int V>
__global__ void __launch_bounds__(256) k(const float* __restrict__ in, float* out, int iters, int n4 /*float4 per warp tile*/) {
    extern __shared__ float4 smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* tile = smem + warp * n4;
    for (int i = lane; i < n4; i += 32) tile[i] = ((const float4*)in)[(blockIdx.x * 8 + warp) % 64 * n4 + i];
    __syncwarp();
    f32x2 a01 = 0ull, a23 = 0ull;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int it = 0; it < iters; ++it) {

        if (V >= 8) {
            f32x2 acc[4] = {a01, a23, 0ull, 0ull};
            float sacc[8] = {s0, s1, s2, s3, 0.f, 0.f, 0.f, 0.f};
            unsigned allneg = 0x80000000u;
            for (int i = lane; i < n4; i += 64) {
                const float4 v = tile[i];
                const float4 w = tile[min(i + 32, n4 - 1)];
                const float x[8] = {v.x, v.y, v.z, v.w, w.x, w.y, w.z, w.w};
                if (V == 8) poly_pairs<4, 10, false>(x, acc, allneg);
                if (V == 9) poly_pairs<4, 9, false>(x, acc, allneg);
                if (V == 10) poly_pairs<4, 9, true>(x, acc, allneg);
                if (V == 11) poly_scalars<8, 9>(x, sacc);
                if (V == 12) { poly_pairs<2, 9, false>(x, acc, allneg); poly_scalars<4, 9>(x + 4, sacc); }
            }
            a01 = add2(acc[0], acc[2]); a23 = add2(acc[1], acc[3]);
            s0 = sacc[0] + sacc[4]; s1 = sacc[1] + sacc[5]; s2 = sacc[2] + sacc[6]; s3 = sacc[3] + sacc[7];
            if (V == 10 && !(allneg >> 31)) s0 += 1.0f;
            continue;
        }
#pragma unroll 2
        for (int i = lane; i < n4; i += 32) {
            const float4 v = tile[i];
            if (V == 0) { a01 = cur2(v.x, v.y, a01); a23 = cur2(v.z, v.w, a23); }
            if (V == 1) { s0 = cur1(v.x, s0); s1 = cur1(v.y, s1); s2 = cur1(v.z, s2); s3 = cur1(v.w, s3); }
            if (V == 2) { a01 = poly2(v.x, v.y, a01); a23 = poly2(v.z, v.w, a23); }
            if (V == 3) { s0 = poly1(v.x, s0); s1 = poly1(v.y, s1); s2 = poly1(v.z, s2); s3 = poly1(v.w, s3); }
            if (V == 4) { a01 = poly2(v.x, v.y, a01); s2 = poly1(v.z, s2); s3 = poly1(v.w, s3); }
            if (V == 5) {   // fast path with the sign test, general path as fallback
                const unsigned sg = __float_as_uint(v.x) & __float_as_uint(v.y) & __float_as_uint(v.z) & __float_as_uint(v.w);
                if ((int)sg < 0) { a01 = poly2(v.x, v.y, a01); a23 = poly2(v.z, v.w, a23); }
                else { a01 = cur2(v.x, v.y, a01); a23 = cur2(v.z, v.w, a23); }
            }
            if (V == 6) {   // scalar fast path with the sign test
                const unsigned sg = __float_as_uint(v.x) & __float_as_uint(v.y) & __float_as_uint(v.z) & __float_as_uint(v.w);
                if ((int)sg < 0) { s0 = poly1(v.x, s0); s1 = poly1(v.y, s1); s2 = poly1(v.z, s2); s3 = poly1(v.w, s3); }
                else { s0 = cur1(v.x, s0); s1 = cur1(v.y, s1); s2 = cur1(v.z, s2); s3 = cur1(v.w, s3); }
            }
            if (V == 7) {   // mixed with the sign test
                const unsigned sg = __float_as_uint(v.x) & __float_as_uint(v.y) & __float_as_uint(v.z) & __float_as_uint(v.w);
                if ((int)sg < 0) { a01 = poly2(v.x, v.y, a01); s2 = poly1(v.z, s2); s3 = poly1(v.w, s3); }
                else { a01 = cur2(v.x, v.y, a01); s2 = cur1(v.z, s2); s3 = cur1(v.w, s3); }
            }
        }
    }
    float u0, u1, u2, u3;
    unpack2(a01, u0, u1); unpack2(a23, u2, u3);
    out[blockIdx.x * blockDim.x + threadIdx.x] = u0 + u1 + u2 + u3 + s0 + s1 + s2 + s3;
}

template <int V>
void run(const char* name, const float* in, float* out, int ctas_per_sm) {
    const int n4 = 180, iters = 400;
    const size_t smem = 8 * n4 * 16;
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<V><<<148 * ctas_per_sm, 256, smem>>>(in, out, iters, n4);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double elems = 148.0 * ctas_per_sm * 8 * n4 * 4.0 * iters;
    printf("%-34s ctas/SM %d: %8.3f ms  %7.2f Gelem/s  %5.2f elem/clk/SMSP @1.965GHz  == %6.0f GB/s of logits\n", name, ctas_per_sm,
           best, elems / best / 1e6, elems / (best * 1e-3) / (148 * 4) / 1.965e9, elems * 4 / best / 1e6);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
}

int main() {
    const int n = 64 * 180 * 4;
    float* h = (float*)malloc(n * 4);
    unsigned s = 12345;
    for (int i = 0; i < n; ++i) {   // roughly N(-4.6, 1): sum of 12 uniforms
        float u = 0;
        for (int j = 0; j < 12; ++j) { s = s * 1664525u + 1013904223u; u += (s >> 8) * (1.0f / 16777216.0f); }
        h[i] = -4.595f + (u - 6.0f);
    }
    float *in, *out;
    cudaMalloc(&in, n * 4); cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaMemcpy(in, h, n * 4, cudaMemcpyHostToDevice);
    for (int c = 2; c <= 6; c += 2) {
        run<0>("cur packed (2 MUFU)", in, out, c);
        run<1>("cur scalar (2 MUFU)", in, out, c);
        run<2>("poly10 packed (1 MUFU)", in, out, c);
        run<3>("poly10 scalar (1 MUFU)", in, out, c);
        run<4>("poly10 mixed xy packed zw scalar", in, out, c);
        run<5>("poly10 packed + sign test", in, out, c);
        run<6>("poly10 scalar + sign test", in, out, c);
        run<7>("poly10 mixed + sign test", in, out, c);
        run<8>("poly10 packed x4 interleaved", in, out, c);
        run<9>("poly9 packed x4 interleaved", in, out, c);
        run<10>("poly9 packed x4 + clamp + flag", in, out, c);
        run<11>("poly9 scalar x8 interleaved", in, out, c);
        run<12>("poly9 2 packed + 4 scalar", in, out, c);
    }
    printf("HBM budget at 6541 GB/s: 1.41 elem/clk/SMSP\n");
    return 0;
}
