// Device pieces of the matcher shared by match_kernel (matcher.cu) and the fused flat-pass + matching kernel (head.cu).
#pragma once
#include "common.cuh"

#define MATCH_THREADS 256
#define GT_CHUNK 512

// matches value from the thresholds: training_target_creation.py:92-100
__device__ __forceinline__ int threshold_match(int best_g, float best_v, float pos_thr, float neg_thr, bool same_thr) {
    if (best_v >= pos_thr) return best_g;
    if (same_thr) return -1;
    return (neg_thr > best_v) ? -1 : -2;
}


// Forced matches: training_target_creation.py:105-126.  For GT g: fid[g] = first anchor with the row maximum,
// ok[g] = (row maximum >= 0.1).  Anchor a is overridden iff some ok GT picked it; the value written is the
// LOWEST GT index among all GTs that picked a, ok or not (argmax over the unmasked one-hot, :117).
// Runs in one CTA per image: either force_match_kernel or the last match_kernel CTA of the image.
template <bool WRITE_TARGETS>
__device__ __forceinline__ void force_match_image(
    int b, int N, int* s_fid, unsigned char* s_ok, const float4* __restrict__ anchors, int A,
    const float4* __restrict__ gt_boxes, const int* __restrict__ gt_labels, int Gmax,
    const unsigned long long* gt_best, int* matches, float4* reg, int* cls, int* s_new_matched = nullptr) {
    for (int g = threadIdx.x; g < N; g += blockDim.x) {
        const unsigned long long key = __ldcg(&gt_best[(size_t)b * Gmax + g]);
        // key == 0: the whole IoU row is 0 -> argmax is anchor 0, value 0
        s_fid[g] = key ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : 0;
        s_ok[g] = __uint_as_float((unsigned)(key >> 32)) >= 0.1f;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < N; g += blockDim.x) {
        const int a = s_fid[g];
        bool first = true, any_ok = false;
        for (int h = 0; h < N; ++h) {
            if (s_fid[h] == a) {
                if (h < g) first = false;
                any_ok |= (s_ok[h] != 0);
            }
        }
        if (first && any_ok) {
            const size_t o = (size_t)b * A + a;
            if (s_new_matched && __ldcg(&matches[o]) < 0) atomicAdd(s_new_matched, 1);   // a forced match of a so far unmatched anchor
            matches[o] = g;
            if (WRITE_TARGETS) {
                reg[o] = box_encode(gt_boxes[(size_t)b * Gmax + g], anchors[a]);
                cls[o] = gt_labels[(size_t)b * Gmax + g] + 1;
            }
        }
    }
}


// One 256-anchor chunk of one image, by a whole CTA of MATCH_THREADS threads: the per-chunk form of match_kernel's body
// (GT boxes staged per chunk, per-GT maxima flushed per chunk), for callers that interleave matching chunks with other work.
// `tickets[b]` counts finished chunks; the CTA that finishes the image's last chunk applies the forced matches.
// Requires N <= GT_CHUNK.  s_box / s_area / s_best: CTA-shared scratch of GT_CHUNK entries each.
struct MatchJob {
    const float4* anchors;
    const float4* gt_boxes;
    const int* gt_labels;
    const int* num_boxes;
    unsigned long long* gt_best;     // [B,Gmax] zeroed
    int* tickets;                    // [B] zeroed
    int* matches;
    float4* reg;
    int* cls;
    int A, Gmax, chunks_per_image;
    float pos_thr, neg_thr;
    int same_thr;
    long long nchunks;               // B * chunks_per_image
};

static __device__ __noinline__ void match_chunk(const MatchJob J, long long job, float4* s_box, float* s_area, unsigned long long* s_best,
                                         int* s_flag) {
    const int b = (int)(job / J.chunks_per_image), chunk = (int)(job - (long long)b * J.chunks_per_image);
    const int lane = threadIdx.x & 31;
    const int A = J.A, Gmax = J.Gmax;
    const int N = J.num_boxes ? min(max(J.num_boxes[b], 0), Gmax) : Gmax;
    const float4* gtb = J.gt_boxes + (size_t)b * Gmax;
    __syncthreads();                                                   // scratch may still be in use by the previous job
    for (int t = threadIdx.x; t < N; t += MATCH_THREADS) {
        const float4 gb = gtb[t];
        s_box[t] = gb;
        s_area[t] = box_area(gb);
        s_best[t] = 0ull;
    }
    __syncthreads();
    const int a = chunk * MATCH_THREADS + threadIdx.x;
    const bool valid = a < A;
    float4 anc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) anc = J.anchors[a];
    const float area_a = box_area(anc);
    float wy0 = valid ? anc.x : INFINITY, wx0 = valid ? anc.y : INFINITY;
    float wy1 = valid ? anc.z : -INFINITY, wx1 = valid ? anc.w : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wy0 = fminf(wy0, __shfl_xor_sync(0xffffffffu, wy0, o));
        wx0 = fminf(wx0, __shfl_xor_sync(0xffffffffu, wx0, o));
        wy1 = fmaxf(wy1, __shfl_xor_sync(0xffffffffu, wy1, o));
        wx1 = fmaxf(wx1, __shfl_xor_sync(0xffffffffu, wx1, o));
    }
    float best_v = 0.0f;
    int best_g = 0;
    for (int t0 = 0; t0 < N; t0 += 32) {
        bool near = false;
        if (t0 + lane < N) {
            const float4 gb = s_box[t0 + lane];
            near = !(gb.z <= wy0 || gb.x >= wy1 || gb.w <= wx0 || gb.y >= wx1);
        }
        unsigned todo = __ballot_sync(0xffffffffu, near);
        while (todo) {
            const int t = t0 + __ffs(todo) - 1;
            todo &= todo - 1;
            const float4 gb = s_box[t];
            const float inter = box_intersection(gb, anc);
            float v = 0.0f;
            if (valid && inter > 0.0f) {
                const float uni = f_sub(f_add(s_area[t], area_a), inter);
                v = fminf(fmaxf(f_div(inter, f_add(uni, SSDK_EPS)), 0.0f), 1.0f);
            }
            if (v > best_v) { best_v = v; best_g = t; }
            const unsigned bits = __float_as_uint(v);
            const unsigned wmax = __reduce_max_sync(0xffffffffu, bits);
            if (wmax != 0u) {
                const unsigned ball = __ballot_sync(0xffffffffu, bits == wmax);
                if (lane == __ffs(ball) - 1)
                    atomicMax(&s_best[t], ((unsigned long long)wmax << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)a));
            }
        }
    }
    if (valid) {
        const int m = (N > 0) ? threshold_match(best_g, best_v, J.pos_thr, J.neg_thr, J.same_thr != 0) : -1;
        const size_t o = (size_t)b * A + a;
        J.matches[o] = m;
        if (m >= 0) {
            J.reg[o] = box_encode(gtb[m], anc);
            J.cls[o] = J.gt_labels[(size_t)b * Gmax + m] + 1;
        } else {
            J.reg[o] = make_float4(0.f, 0.f, 0.f, 0.f);
            J.cls[o] = 0;
        }
    }
    if (N == 0) return;
    __syncthreads();
    for (int t = threadIdx.x; t < N; t += MATCH_THREADS)
        if (s_best[t] != 0ull) atomicMax(&J.gt_best[(size_t)b * Gmax + t], s_best[t]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *s_flag = (atomicAdd(&J.tickets[b], 1) == J.chunks_per_image - 1);
    __syncthreads();
    if (*s_flag) {
        __threadfence();
        force_match_image<true>(b, N, (int*)s_area, (unsigned char*)s_box, J.anchors, A, J.gt_boxes, J.gt_labels, Gmax, J.gt_best,
                                J.matches, J.reg, J.cls);
    }
}
