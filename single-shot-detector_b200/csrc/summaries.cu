// Observability by-products of the training path: the per-level statistics behind the reference's TensorBoard summaries
// (detector/ssd.py:125-129 and :135-163), computed on the GPU from the per-anchor vectors the loss kernels can emit.
//   level_matches_kernel   matched anchors (matches >= 0) per (image, level)                     ssd.py:152-163, :129
//   level_topk_kernel      for a per-anchor loss vector: the k_l = ceil(n_l * 0.2) biggest values of every image on
//                          level l (tf.nn.top_k, ssd.py:146-147), reduced to what does not depend on top_k's unspecified
//                          output order: their mean and the k_l-th biggest value.  Exact selection by a 4-pass, 8-bit radix
//                          select on the order-preserving integer image of the floats (one CTA per (image, level), the
//                          level's slice stays in L2), then one pass that sums everything above the threshold in double.
#include "common.cuh"

#define SUMM_THREADS 1024

__global__ void __launch_bounds__(256) level_matches_kernel(const int* __restrict__ matches, long long A, int L,
                                                            const int* __restrict__ level_off /*[L+1]*/, float* __restrict__ out) {
    const int l = blockIdx.x, b = blockIdx.y;
    const int* m = matches + (size_t)b * A + level_off[l];
    const int n = level_off[l + 1] - level_off[l];
    int c = 0;
    for (int i = threadIdx.x; i < n; i += 256) c += m[i] >= 0;
    __shared__ int s_c[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_c[w];
        out[(size_t)b * L + l] = (float)t;
    }
}

__device__ __forceinline__ unsigned float_order(float v) {     // ascending unsigned order == ascending float order
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float order_float(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

__global__ void __launch_bounds__(SUMM_THREADS) level_topk_kernel(const float* __restrict__ values, long long A, int L,
                                                                  const int* __restrict__ level_off, const int* __restrict__ level_k,
                                                                  float* __restrict__ out_mean, float* __restrict__ out_kth) {
    const int l = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const float* v = values + (size_t)b * A + level_off[l];
    const int n = level_off[l + 1] - level_off[l];
    int k = level_k[l];
    if (k > n) k = n;
    if (n == 0 || k <= 0) {
        if (tid == 0) { out_mean[(size_t)b * L + l] = 0.0f; out_kth[(size_t)b * L + l] = 0.0f; }
        return;
    }
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_prefix, s_need;
    // radix select of the k-th biggest key: after each pass `prefix` holds the decided high bits and `need` the rank still
    // to be found among the keys that share them
    unsigned prefix = 0u, need = (unsigned)k;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const unsigned hi_mask = pass ? (0xFFFFFFFFu << (shift + 8)) : 0u;
        if (tid < 256) s_hist[tid] = 0u;
        __syncthreads();
        for (int i = tid; i < n; i += SUMM_THREADS) {
            const unsigned o = float_order(v[i]);
            if ((o & hi_mask) == prefix) atomicAdd(&s_hist[(o >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned left = need;
            int bin = 255;
            for (; bin > 0; --bin) {
                if (s_hist[bin] >= left) break;
                left -= s_hist[bin];
            }
            s_prefix = prefix | ((unsigned)bin << shift);
            s_need = left;
        }
        __syncthreads();
        prefix = s_prefix;
        need = s_need;
        __syncthreads();
    }
    // prefix = key of the k-th biggest value T; `need` of the keys equal to T belong to the selection
    const float T = order_float(prefix);
    double acc = 0.0;
    for (int i = tid; i < n; i += SUMM_THREADS) {
        const float x = v[i];
        if (float_order(x) > prefix) acc += (double)x;
    }
    __shared__ double s_sum[SUMM_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) s_sum[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < SUMM_THREADS / 32; ++w) t += s_sum[w];
        t += (double)need * (double)T;
        out_mean[(size_t)b * L + l] = (float)(t / (double)k);
        out_kth[(size_t)b * L + l] = T;
    }
}

extern "C" int ssdk_level_summaries(ssdk_ctx* ctx, const float* values, const int32_t* matches, int B, int64_t A,
                                    const int32_t* per_level, int num_levels, double top_fraction, float* out_topk_mean,
                                    float* out_topk_kth, float* out_matches) {
    SSDK_ENTER(ctx);
    SsdkWsGuard ws_guard(ctx, SSDK_WS_SUMM);
    SSDK_REQUIRE(B >= 0 && A >= 0 && num_levels >= 1 && num_levels <= 64 && per_level, SSDK_ERR_ARG, "ssdk_level_summaries: bad sizes");
    SSDK_REQUIRE(B <= 65535, SSDK_ERR_SHAPE, "ssdk_level_summaries: batch %d > 65535", B);
    SSDK_REQUIRE(top_fraction > 0.0 && top_fraction <= 1.0, SSDK_ERR_ARG, "ssdk_level_summaries: top_fraction %g not in (0,1]", top_fraction);
    SSDK_REQUIRE(!values || (out_topk_mean && out_topk_kth), SSDK_ERR_ARG, "ssdk_level_summaries: values given without outputs");
    SSDK_REQUIRE(!matches || out_matches, SSDK_ERR_ARG, "ssdk_level_summaries: matches given without output");
    int host[2 * 64 + 1];
    long long off = 0;
    const float frac = (float)top_fraction;
    for (int l = 0; l < num_levels; ++l) {
        SSDK_REQUIRE(per_level[l] >= 0, SSDK_ERR_ARG, "ssdk_level_summaries: negative level size");
        host[l] = (int)off;
        const volatile float nk = (float)per_level[l] * frac;            // float32 product, as tf.to_float(n) * 0.20 (ssd.py:146)
        host[num_levels + 1 + l] = (int)ceilf(nk);
        off += per_level[l];
    }
    host[num_levels] = (int)off;
    SSDK_REQUIRE(off == A, SSDK_ERR_SHAPE, "ssdk_level_summaries: levels hold %lld anchors, expected A = %lld", off, (long long)A);
    if (B == 0) return SSDK_OK;
    const size_t tab_bytes = (size_t)(2 * num_levels + 1) * sizeof(int);
    SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_summ, 1024));
    SSDK_CHECK_CUDA(cudaMemcpyAsync(ctx->ws_summ.p, host, tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
    // the table is consumed by the kernels below on the same stream; `host` is pageable stack memory, so the copy has
    // been staged by the driver when cudaMemcpyAsync returns
    const int* d_off = (const int*)ctx->ws_summ.p;
    const int* d_k = d_off + num_levels + 1;
    const dim3 grid(num_levels, B);
    if (matches)
        SSDK_KERNEL(ctx, SSDK_K_OTHER, level_matches_kernel<<<grid, 256, 0, ctx->stream>>>(matches, A, num_levels, d_off, out_matches));
    if (values)
        SSDK_KERNEL(ctx, SSDK_K_OTHER,
                    level_topk_kernel<<<grid, SUMM_THREADS, 0, ctx->stream>>>(values, A, num_levels, d_off, d_k, out_topk_mean, out_topk_kth));
    return SSDK_OK;
}
