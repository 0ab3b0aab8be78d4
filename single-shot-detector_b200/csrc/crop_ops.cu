// Ground-truth side box operations of the random-crop augmentation, detector/input_pipeline/random_image_crop.py:
// ioa :190-209, prune_completely_outside_window :102-131, prune_non_overlapping_boxes :134-159,
// change_coordinate_frame :162-187, and their composition inside randomly_crop_image (:86-99) for a batch of images.
// They reuse area / intersection (box_utils.py:30-61).  The crop WINDOW itself comes from TensorFlow's
// sample_distorted_bounding_box (JPEG decoding and the random sampler are input-pipeline work, out of scope).
// Sizes are tiny (at most a few hundred boxes per image): one CTA per image / call, ordered compaction with ballots, so
// that the kept indices come out ascending exactly as tf.where + tf.gather produce them.
#include "common.cuh"

#define CROP_THREADS 256
#define CROP_MAX_BOXES 4096

// ioa(boxes1, boxes2)[i, j] = clip(intersection(b1_i, b2_j) / (area(b2_j) + eps), 0, 1)
__device__ __forceinline__ float box_ioa(const float4 b1, const float4 b2) {
    const float q = f_div(box_intersection(b1, b2), f_add(box_area(b2), SSDK_EPS));
    return fminf(fmaxf(q, 0.0f), 1.0f);
}

__global__ void __launch_bounds__(256) ioa_kernel(const float4* __restrict__ b1, long long n, const float4* __restrict__ b2,
                                                  long long m, float* __restrict__ out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const float4 bj = b2[j];
    for (long long i = blockIdx.y; i < n; i += gridDim.y) out[i * m + j] = box_ioa(b1[i], bj);
}

// change_coordinate_frame: subtract the window origin, divide by the window size, clip to [0, 1]
__device__ __forceinline__ float4 change_frame(const float4 b, const float4 w) {
    const float wh = f_sub(w.z, w.x), ww = f_sub(w.w, w.y);
    float4 r;
    r.x = f_div(f_sub(b.x, w.x), wh); r.y = f_div(f_sub(b.y, w.y), ww);
    r.z = f_div(f_sub(b.z, w.x), wh); r.w = f_div(f_sub(b.w, w.y), ww);
    return box_clip01(r);
}

__global__ void __launch_bounds__(256) change_frame_kernel(const float4* __restrict__ b, long long n, const float4* __restrict__ window,
                                                           float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = change_frame(b[i], window[0]);
}

__device__ __forceinline__ bool outside_window(const float4 b, const float4 w) {   // :121-124
    return b.x >= w.z || b.y >= w.w || b.z <= w.x || b.w <= w.y;
}

// Ordered compaction of the boxes of one image that pass `keep`: blockDim = CROP_THREADS, any n.
// MODE 0: prune_completely_outside_window   (keep = !outside(window))
// MODE 1: prune_non_overlapping_boxes       (keep = max_j ioa(boxes2_j, box) >= min_overlap)
// MODE 2: randomly_crop_image :86-99        (keep = MODE 0 && MODE 1 with boxes2 = {window}; boxes moved to the window's frame)
template <int MODE>
__global__ void __launch_bounds__(CROP_THREADS) prune_kernel(const float4* __restrict__ boxes, const int* __restrict__ num_boxes,
                                                             int Gmax, const float4* __restrict__ boxes2, int m, float min_overlap,
                                                             float4* __restrict__ out_boxes, int* __restrict__ out_idx,
                                                             int* __restrict__ out_num) {
    __shared__ int s_warp[CROP_THREADS / 32];
    __shared__ int s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = num_boxes ? min(max(num_boxes[b], 0), Gmax) : Gmax;
    const float4* in = boxes + (size_t)b * Gmax;
    float4* ob = out_boxes + (size_t)b * Gmax;
    int* oi = out_idx + (size_t)b * Gmax;
    const float4* b2 = (MODE == 2) ? boxes2 + b : boxes2;            // MODE 2: one window per image
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += CROP_THREADS) {
        const int i = i0 + tid;
        bool keep = false;
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) {
            bx = in[i];
            if (MODE == 0) keep = !outside_window(bx, b2[0]);
            if (MODE == 1) {
                float best = 0.0f;                                     // ioa is clipped to [0,1]; m >= 1
                for (int j = 0; j < m; ++j) best = fmaxf(best, box_ioa(b2[j], bx));
                keep = best >= min_overlap;
            }
            if (MODE == 2) keep = !outside_window(bx, b2[0]) && box_ioa(b2[0], bx) >= min_overlap;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (keep) {
            const int pos = before + __popc(bal & ((1u << lane) - 1u));
            ob[pos] = (MODE == 2) ? change_frame(bx, b2[0]) : bx;
            oi[pos] = i;
        }
        __syncthreads();
        if (tid == 0) {
            int t = s_base;
            for (int w = 0; w < CROP_THREADS / 32; ++w) t += s_warp[w];
            s_base = t;
        }
        __syncthreads();
    }
    const int kept = s_base;
    for (int i = kept + tid; i < Gmax; i += CROP_THREADS) {            // zero padding (the pipeline's padded format)
        ob[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        oi[i] = -1;
    }
    if (tid == 0) out_num[b] = kept;
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

extern "C" {

int ssdk_ioa(ssdk_ctx* ctx, const float* boxes1, int64_t n, const float* boxes2, int64_t m, float* out) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(n >= 0 && m >= 0, SSDK_ERR_ARG, "ssdk_ioa: negative size");
    if (n == 0 || m == 0) return SSDK_OK;
    SSDK_REQUIRE(boxes1 && boxes2 && out, SSDK_ERR_ARG, "ssdk_ioa: null pointer");
    SSDK_REQUIRE(aligned16(boxes1) && aligned16(boxes2), SSDK_ERR_SHAPE, "ssdk_ioa: box arrays must be 16-byte aligned");
    const dim3 grid(ceil_div_i(m, 256), (unsigned)(n < 1024 ? n : 1024));
    SSDK_KERNEL(ctx, SSDK_K_OTHER, ioa_kernel<<<grid, 256, 0, ctx->stream>>>((const float4*)boxes1, n, (const float4*)boxes2, m, out));
    return SSDK_OK;
}

int ssdk_change_coordinate_frame(ssdk_ctx* ctx, const float* boxes, int64_t n, const float* window, float* out) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(n >= 0, SSDK_ERR_ARG, "ssdk_change_coordinate_frame: negative size");
    if (n == 0) return SSDK_OK;
    SSDK_REQUIRE(boxes && window && out, SSDK_ERR_ARG, "ssdk_change_coordinate_frame: null pointer");
    SSDK_REQUIRE(aligned16(boxes) && aligned16(window) && aligned16(out), SSDK_ERR_SHAPE,
                 "ssdk_change_coordinate_frame: arrays must be 16-byte aligned");
    SSDK_KERNEL(ctx, SSDK_K_OTHER,
                change_frame_kernel<<<ceil_div_i(n, 256), 256, 0, ctx->stream>>>((const float4*)boxes, n, (const float4*)window, (float4*)out));
    return SSDK_OK;
}

static int prune_check(const char* who, const float* boxes, int64_t n, const float* other, float* out_boxes, int32_t* out_idx,
                       int32_t* out_num) {
    SSDK_REQUIRE(n >= 0 && n <= (1 << 24), SSDK_ERR_ARG, "%s: bad size", who);
    SSDK_REQUIRE(out_num != nullptr, SSDK_ERR_ARG, "%s: out_num is NULL", who);
    SSDK_REQUIRE(n == 0 || (boxes && other && out_boxes && out_idx), SSDK_ERR_ARG, "%s: null pointer", who);
    SSDK_REQUIRE(aligned16(boxes) && aligned16(other) && aligned16(out_boxes), SSDK_ERR_SHAPE, "%s: box arrays must be 16-byte aligned", who);
    return SSDK_OK;
}

int ssdk_prune_completely_outside_window(ssdk_ctx* ctx, const float* boxes, int64_t n, const float* window, float* out_boxes,
                                         int32_t* out_indices, int32_t* out_num) {
    SSDK_ENTER(ctx);
    SSDK_TRY(prune_check("ssdk_prune_completely_outside_window", boxes, n, window, out_boxes, out_indices, out_num));
    SSDK_KERNEL(ctx, SSDK_K_OTHER,
                prune_kernel<0><<<1, CROP_THREADS, 0, ctx->stream>>>((const float4*)boxes, nullptr, (int)n, (const float4*)window, 1, 0.0f,
                                                                    (float4*)out_boxes, out_indices, out_num));
    return SSDK_OK;
}

int ssdk_prune_non_overlapping_boxes(ssdk_ctx* ctx, const float* boxes1, int64_t n, const float* boxes2, int64_t m, double min_overlap,
                                     float* out_boxes, int32_t* out_indices, int32_t* out_num) {
    SSDK_ENTER(ctx);
    SSDK_TRY(prune_check("ssdk_prune_non_overlapping_boxes", boxes1, n, boxes2, out_boxes, out_indices, out_num));
    SSDK_REQUIRE(m >= 1 && m <= (1 << 20), SSDK_ERR_ARG, "ssdk_prune_non_overlapping_boxes: boxes2 must hold at least one box");
    SSDK_KERNEL(ctx, SSDK_K_OTHER,
                prune_kernel<1><<<1, CROP_THREADS, 0, ctx->stream>>>((const float4*)boxes1, nullptr, (int)n, (const float4*)boxes2, (int)m,
                                                                    (float)min_overlap, (float4*)out_boxes, out_indices, out_num));
    return SSDK_OK;
}

int ssdk_crop_boxes(ssdk_ctx* ctx, const float* boxes, const int32_t* num_boxes, const float* windows, int B, int Gmax,
                    double overlap_thresh, float* out_boxes, int32_t* out_keep_indices, int32_t* out_num) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(B >= 0 && Gmax >= 0, SSDK_ERR_ARG, "ssdk_crop_boxes: negative size");
    if (B == 0) return SSDK_OK;
    SSDK_REQUIRE(windows && out_num && (Gmax == 0 || (boxes && out_boxes && out_keep_indices)), SSDK_ERR_ARG, "ssdk_crop_boxes: null pointer");
    SSDK_REQUIRE(aligned16(boxes) && aligned16(windows) && aligned16(out_boxes), SSDK_ERR_SHAPE, "ssdk_crop_boxes: box arrays must be 16-byte aligned");
    SSDK_KERNEL(ctx, SSDK_K_OTHER,
                prune_kernel<2><<<B, CROP_THREADS, 0, ctx->stream>>>((const float4*)boxes, num_boxes, Gmax, (const float4*)windows, 1,
                                                                    (float)overlap_thresh, (float4*)out_boxes, out_keep_indices, out_num));
    return SSDK_OK;
}

}  // extern "C"
