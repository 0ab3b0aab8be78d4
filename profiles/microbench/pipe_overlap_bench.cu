// Micro-benchmark: do MUFU (XU pipe) and FFMA (FMA pipe) overlap on sm_100a, or do their cycles add?
// Each thread keeps 8 independent chains; per iteration it issues NM MUFU.EX2 and NF FFMA per chain-set.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_overlap_bench pipe_overlap_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmx(float a, float b) { float d; asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
template <int NM, int NF, int NA>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    float a[8], m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = threadIdx.x * 1e-3f + j; m[j] = -1.0f - j * 0.01f - threadIdx.x * 1e-4f; }
    const float b = 0.999f, c = 1e-4f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < NF; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma1(a[j], b, c);
#pragma unroll
        for (int r = 0; r < NM; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = ex2a(m[j]) - 1.5f * 0.0f - 1.0f;   // keeps the argument in range; the sub is folded below
#pragma unroll
        for (int r = 0; r < NA; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fmx(a[j], m[j]);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j] + m[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NM, int NF, int NA>
void run() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<NM, NF, NA><<<148 * 8, 256>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // warp-instructions per SMSP: 148*8 CTAs * 8 warps / (148*4 SMSP) = 16 warps per SMSP
    const double cyc = best * 1e-3 * 1.965e9;
    const double per_iter = cyc / iters / 16.0 / 8.0;    // cycles per (chain-set of NM mufu [+NM fadd] + NF ffma + NA fmnmx) per warp per SMSP
    printf("MUFU %d (+%d FADD)  FFMA %d  FMNMX %d : %7.3f ms  -> %6.2f SMSP-cycles per group  (if additive: 8*%d + %d + %d + %d/2?)\n", NM, NM, NF, NA, best, per_iter, NM, NM, NF, NA);
    cudaFree(out);
}
int main() {
    run<1, 0, 0>(); run<0, 8, 0>(); run<0, 16, 0>(); run<1, 8, 0>(); run<1, 16, 0>(); run<2, 16, 0>(); run<1, 24, 0>();
    run<0, 8, 8>(); run<0, 16, 8>(); run<1, 8, 8>(); run<0, 0, 8>();
    return 0;
}
