// Micro-benchmark: issue throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
template <int MODE>
__global__ void k(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float b = 0.999f, c = 1e-4f;
    if (MODE == 0) {
        for (int i = 0; i < iters; ++i) {
            a0 = fma1(a0, b, c); a1 = fma1(a1, b, c); a2 = fma1(a2, b, c); a3 = fma1(a3, b, c);
            a4 = fma1(a4, b, c); a5 = fma1(a5, b, c); a6 = fma1(a6, b, c); a7 = fma1(a7, b, c);
        }
    } else {
        unsigned long long p0, p1, p2, p3, bb, cc;
        float2 t;
        t = make_float2(a0, a1); p0 = *(unsigned long long*)&t; t = make_float2(a2, a3); p1 = *(unsigned long long*)&t;
        t = make_float2(a4, a5); p2 = *(unsigned long long*)&t; t = make_float2(a6, a7); p3 = *(unsigned long long*)&t;
        t = make_float2(b, b); bb = *(unsigned long long*)&t; t = make_float2(c, c); cc = *(unsigned long long*)&t;
        for (int i = 0; i < iters; ++i) {
            p0 = fma2(p0, bb, cc); p1 = fma2(p1, bb, cc); p2 = fma2(p2, bb, cc); p3 = fma2(p3, bb, cc);
        }
        t = *(float2*)&p0; a0 = t.x; a1 = t.y; t = *(float2*)&p1; a2 = t.x; a3 = t.y;
        t = *(float2*)&p2; a4 = t.x; a5 = t.y; t = *(float2*)&p3; a6 = t.x; a7 = t.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters); else k<1><<<148 * 8, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fmas = 148.0 * 8 * 256 * 8.0 * iters;
            printf("%s: %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM @1.965GHz)\n", mode ? "FFMA2" : "FFMA ", ms, fmas / ms / 1e9, fmas / (ms * 1e-3) / 148 / 1.965e9);
        }
    }
    return 0;
}
