// Inference post-processing: score threshold -> per-(image, class) candidate lists -> sort -> greedy NMS -> pack.
// Replaces detector/utils/nms.py:48-102 (batch_multiclass_non_max_suppression), :6-45
// (multiclass_non_max_suppression), the sigmoid of detector/ssd.py:60 and TensorFlow 1.12's NonMaxSuppressionV3
// (external C++ kernel called at nms.py:33; semantics restated in oracle/nms.py).
//
// Pipeline (all on the context's stream, no host synchronisation):
//   1. filter_kernel      streams the [B,A,C] scores (or logits) once with 128-bit no-allocate loads and appends a
//                         packed 64-bit key per (anchor, class) with score > threshold to the candidate list of its
//                         (image, class) SEGMENT (a fixed region of A keys -- a class cannot have more candidates than
//                         anchors -- and one atomic counter per segment).  This is the HBM-bound kernel.
//                         key = class | ~order(score) | anchor  ->  ascending u64 order == score descending, anchor
//                         index ascending inside a segment.
//                         The reference's `is_confident` anchor pre-filter (nms.py:71-74, '>=') only removes
//                         anchors that have no candidate at all (candidates need '>'), so it cannot change any
//                         output and is not materialised.
//   2. nms_small_kernel   one WARP per segment with at most 32 candidates (the vast majority): the keys are sorted with a
//                         shuffle bitonic network, decoded (box_utils.py:114-142) and clipped (nms.py:77), a 32x32
//                         suppression bit matrix is built and walked greedily.  Larger segments are queued.
//   3. nms_kernel         one CTA per queued segment: bitonic sort of the segment's keys (shared memory up to 4096 keys,
//                         in place in global memory beyond), then candidates are taken 64 at a time in sorted order (eight
//                         threads per candidate), tested against the boxes kept so far (shared memory) and against the
//                         chunk's earlier candidates (64x64 bit matrix); the greedy order is resolved with ballots; stops at K.
//   4. pack_kernel        class-major concatenation, zero padding to C*K and num_boxes (nms.py:83-93).
// There is no separate sort pass and nothing of the size of a whole image is ever sorted: the filter buckets by class,
// each segment is sorted by the warp / CTA that runs its NMS.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

#define FILTER_THREADS 256
#define FILTER_UNROLL 4
#define NMS_SORT_SMEM_KEYS 4096       // heavy segments up to this many keys are sorted in shared memory (32 KB)
#define NMS_THREADS 512
#define NMS_CH 64                    // candidates per chunk
#define NMS_TPC (NMS_THREADS / NMS_CH)  // threads per candidate (a power of two <= 32)

struct KeyFormat {
    int abits;       // bits for the anchor index
    int cshift;      // 32 + abits
};

__device__ __forceinline__ unsigned order_desc(float s) {
    unsigned u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending total order of floats
    return ~u;                                        // descending
}
__device__ __forceinline__ float key_score(unsigned long long key, KeyFormat f) {
    unsigned u = ~(unsigned)((key >> f.abits) & 0xFFFFFFFFull);
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    return __uint_as_float(u);
}
__device__ __forceinline__ int key_anchor(unsigned long long key, KeyFormat f) {
    return (int)(key & ((1ull << f.abits) - 1ull));
}
__device__ __forceinline__ int key_class(unsigned long long key, KeyFormat f) { return (int)(key >> f.cshift); }
__device__ __forceinline__ unsigned long long make_key(int c, float s, int a, KeyFormat f) {
    return ((unsigned long long)c << f.cshift) | ((unsigned long long)order_desc(s) << f.abits) | (unsigned long long)a;
}

// ---------------------------------------------------------------------------------------------- 1. filter
template <bool IS_LOGITS>
__device__ __forceinline__ bool is_candidate(float v, float thr, float x_lo, float* score) {
    if (IS_LOGITS) {
        if (!(v > x_lo)) return false;                               // cheap reject in logit space (with margin)
        const float s = f_div(1.0f, f_add(1.0f, expf(-v)));         // tf.sigmoid (ssd.py:60)
        *score = s;
        return s > thr;
    }
    *score = v;
    return v > thr;                                                  // NonMaxSuppressionV3: strict '>'
}

// A candidate takes the next slot of its (image, class) segment.  The lanes that arrive here together and target the same
// segment (the normal case when a channels_first plane is scanned: 128 consecutive floats share image and class) share one
// atomic; otherwise one atomic per candidate -- candidates are rare unless the scores are dense.
__device__ __forceinline__ void append_key(unsigned long long* __restrict__ cand, long long capc, int* __restrict__ seg_count,
                                           long long seg, unsigned long long key) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    const long long seg_l = __shfl_sync(m, seg, leader);
    int pos;
    if (__all_sync(m, seg == seg_l)) {
        int first = 0;
        if (lane == leader) first = atomicAdd(seg_count + seg, __popc(m));
        pos = __shfl_sync(m, first, leader) + __popc(m & ((1u << lane) - 1u));
    } else {
        pos = atomicAdd(seg_count + seg, 1);
    }
    cand[(size_t)seg * capc + pos] = key;
}

// The streaming scan shared by both layouts: `count` floats at `base` are read once with 128-bit no-allocate loads;
// emit(e, score) is called for every element e with score > threshold.
template <bool IS_LOGITS, typename Emit>
__device__ __forceinline__ void scan_candidates(const float* __restrict__ base, long long count, float thr, float x_lo, Emit emit) {
    const int lane = threadIdx.x & 31;
    // peel to 16-byte alignment: head scalars | body float4 | tail scalars
    const unsigned mis = (unsigned)(((uintptr_t)base >> 2) & 3);
    long long head = mis ? (4 - mis) : 0;
    if (head > count) head = count;
    const long long nbody4 = (count - head) >> 2;
    const long long tail0 = head + (nbody4 << 2);
    const float4* body = (const float4*)(base + head);

    if (blockIdx.x == 0 && threadIdx.x < 32) {
        // head + tail elements (< 8 in total), one lane each
        long long e = -1;
        if (lane < head) e = lane;
        else if (lane - head < count - tail0) e = tail0 + (lane - head);
        if (e >= 0) {
            float s;
            if (is_candidate<IS_LOGITS>(base[e], thr, x_lo, &s)) emit(e, s);
        }
    }

    const long long stride = (long long)gridDim.x * FILTER_THREADS * FILTER_UNROLL;
    for (long long i0 = (long long)blockIdx.x * FILTER_THREADS * FILTER_UNROLL; i0 < nbody4; i0 += stride) {
        float4 v[FILTER_UNROLL];
        bool inb[FILTER_UNROLL];
#pragma unroll
        for (int u = 0; u < FILTER_UNROLL; ++u) {
            const long long i = i0 + u * FILTER_THREADS + threadIdx.x;
            inb[u] = i < nbody4;
            v[u] = inb[u] ? ld_stream_f4(body + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
#pragma unroll
        for (int u = 0; u < FILTER_UNROLL; ++u) {
            const float lim = IS_LOGITS ? x_lo : thr;
            const float mx = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
            if (!(inb[u] && (mx > lim))) continue;
            const long long e0 = head + ((i0 + u * FILTER_THREADS + threadIdx.x) << 2);
            // fully unrolled on purpose: the compiler predicates this rare path; written as a rolled loop it becomes a
            // branch region per group and the scan loses 30 % of its bandwidth (measured: 4.3 vs 6.2 TB/s)
            const float vals[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float sc;
                if (is_candidate<IS_LOGITS>(vals[j], thr, x_lo, &sc)) emit(e0 + j, sc);
            }
        }
    }
}

// anchor-major layout: grid (gx, B), image blockIdx.y is the [A,C] array scanned
template <bool IS_LOGITS>
__global__ void __launch_bounds__(FILTER_THREADS) filter_kernel(
    const float* __restrict__ scores, long long per_image /*A*C*/, int C, float thr, float x_lo, KeyFormat fmt,
    unsigned long long* __restrict__ cand, long long capc /*keys per segment = A*/, int* __restrict__ seg_count /*[B*C]*/) {
    const int b = blockIdx.y;
    const long long seg0 = (long long)b * C;                            // first segment of this image
    auto emit = [&](long long e, float s) {
        const int a = (int)(e / C), c = (int)(e - (long long)a * C);
        append_key(cand, capc, seg_count, seg0 + c, make_key(c, s, a, fmt));
    };
    scan_candidates<IS_LOGITS>(scores + (size_t)b * per_image, per_image, thr, x_lo, emit);
}

// Dense scores (a large fraction of all (anchor, class) pairs above the threshold; BASELINE.json's stress config): one
// global atomic per candidate would mean tens of millions of atomics on B*C counters.  Here a CTA works on tiles of 4096
// consecutive elements (every class occurs 4096/C times in a tile): candidates take a rank from a shared-memory histogram,
// one global atomic per class and tile reserves the slots, then the keys are written.  Chosen by the host when the previous
// call on this context met segments with more than NMS_SORT_SMEM_KEYS candidates (a hint: both kernels are correct for any
// input).  grid (gx, B), dynamic shared memory 2*C ints.
#define FD_U 4
template <bool IS_LOGITS>
__global__ void __launch_bounds__(FILTER_THREADS) filter_dense_kernel(
    const float* __restrict__ scores, long long per_image /*A*C*/, int C, float thr, float x_lo, KeyFormat fmt,
    unsigned long long* __restrict__ cand, long long capc, int* __restrict__ seg_count) {
    extern __shared__ int s_dense[];
    int* s_hist = s_dense;
    int* s_base = s_dense + C;
    const int b = blockIdx.y, tid = threadIdx.x;
    const float* base = scores + (size_t)b * per_image;
    const long long seg0 = (long long)b * C;
    for (int c = tid; c < C; c += FILTER_THREADS) s_hist[c] = 0;
    __syncthreads();
    const unsigned mis = (unsigned)(((uintptr_t)base >> 2) & 3);
    long long head = mis ? (4 - mis) : 0;
    if (head > per_image) head = per_image;
    const long long nbody4 = (per_image - head) >> 2;
    const long long tail0 = head + (nbody4 << 2);
    const float4* body = (const float4*)(base + head);
    if (blockIdx.x == 0 && tid < 32) {                                  // the (< 8) unaligned head / tail elements
        long long e = -1;
        if (tid < head) e = tid;
        else if (tid - head < per_image - tail0) e = tail0 + (tid - head);
        float s;
        if (e >= 0 && is_candidate<IS_LOGITS>(base[e], thr, x_lo, &s)) {
            const int a = (int)(e / C), c = (int)(e - (long long)a * C);
            cand[(size_t)(seg0 + c) * capc + atomicAdd(seg_count + seg0 + c, 1)] = make_key(c, s, a, fmt);
        }
    }
    const long long ntiles = (nbody4 + FILTER_THREADS * FD_U - 1) / (FILTER_THREADS * FD_U);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        float sc[4 * FD_U];
        unsigned short rk[4 * FD_U];
        unsigned hit = 0u;
        // phase 1: candidates take their rank inside (tile, class)
#pragma unroll
        for (int u = 0; u < FD_U; ++u) {
            const long long i4 = tile * (FILTER_THREADS * FD_U) + u * FILTER_THREADS + tid;
            const float4 v = i4 < nbody4 ? ld_stream_f4(body + i4) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            const float vals[4] = {v.x, v.y, v.z, v.w};
            const unsigned e0 = (unsigned)(head + (i4 << 2));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (is_candidate<IS_LOGITS>(vals[j], thr, x_lo, &sc[u * 4 + j])) {
                    const unsigned e = e0 + j, a = e / (unsigned)C, c = e - a * (unsigned)C;
                    rk[u * 4 + j] = (unsigned short)atomicAdd(&s_hist[c], 1);
                    hit |= 1u << (u * 4 + j);
                }
            }
        }
        __syncthreads();
        // phase 2: one global atomic per class reserves the tile's slots
        for (int c = tid; c < C; c += FILTER_THREADS) {
            const int h = s_hist[c];
            if (h) {
                s_base[c] = atomicAdd(seg_count + seg0 + c, h);
                s_hist[c] = 0;
            }
        }
        __syncthreads();
        // phase 3: write the keys
#pragma unroll
        for (int u = 0; u < FD_U; ++u) {
            const long long i4 = tile * (FILTER_THREADS * FD_U) + u * FILTER_THREADS + tid;
            const unsigned e0 = (unsigned)(head + (i4 << 2));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if ((hit >> (u * 4 + j)) & 1u) {
                    const unsigned e = e0 + j, a = e / (unsigned)C, c = e - a * (unsigned)C;
                    cand[(size_t)(seg0 + c) * capc + s_base[c] + rk[u * 4 + j]] = make_key((int)c, sc[u * 4 + j], (int)a, fmt);
                }
            }
        }
        // no barrier here: the next tile's phase 2 (the only writer of s_base) comes after its phase-1 barrier, which every
        // thread reaches only after this phase 3
    }
}

// Head layout (head-layout fusion, see head.cu): grid (gx, num_levels); level blockIdx.y's class tensor [B, n*C, h, w] (or
// [B, h, w, n*C]) is scanned as ONE flat array; only a candidate pays for the index arithmetic that recovers
// (image, anchor, class) from its flat position.
struct LevelGeom {        // one level of a HeadGeom, by value
    int hw, per_loc, C, channels_first, anchor_off;
};
__device__ __forceinline__ void head_decompose(const LevelGeom g, long long e, int& b, int& a, int& c) {
    const int hw = g.hw, n = g.per_loc, C = g.C;
    const long long per_image = (long long)n * C * hw;
    b = (int)(e / per_image);
    const int r = (int)(e - (long long)b * per_image);
    int loc, q;
    if (g.channels_first) { q = r / hw; loc = r - q * hw; }
    else { loc = r / (n * C); q = r - loc * (n * C); }
    const int k = q / C;
    c = q - k * C;
    a = g.anchor_off + loc * n + k;
}

template <bool IS_LOGITS>
__global__ void __launch_bounds__(FILTER_THREADS) head_filter_kernel(const HeadGeom G, int B, float thr, float x_lo, KeyFormat fmt,
                                                                    unsigned long long* __restrict__ cand, long long capc,
                                                                    int* __restrict__ seg_count) {
    const int l = blockIdx.y;
    LevelGeom g;
    g.hw = G.hw[l]; g.per_loc = G.per_loc; g.C = G.C; g.channels_first = G.channels_first; g.anchor_off = G.anchor_off[l];
    const long long count = (long long)B * g.per_loc * g.C * g.hw;
    auto emit = [&](long long e, float s) {
        int b, a, c;
        head_decompose(g, e, b, a, c);
        append_key(cand, capc, seg_count, (long long)b * g.C + c, make_key(c, s, a, fmt));
    };
    scan_candidates<IS_LOGITS>(G.cls[l], count, thr, x_lo, emit);
}

// ---------------------------------------------------------------------------------------------- 2. sorting helpers
// 32 keys, one per lane, ascending (bitonic network on shuffles; ~0 pads sort last).
__device__ __forceinline__ unsigned long long warp_sort_keys(unsigned long long key, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, j);
            const bool up = (lane & k) == 0;                // k == 32: every lane ascending
            const bool lower = (lane & j) == 0;
            const unsigned long long lo = key < other ? key : other, hi = key < other ? other : key;
            key = (lower == up) ? lo : hi;
        }
    }
    return key;
}

// CTA-wide bitonic sort of keys[0, n) (shared or global memory), ascending, for ANY n: the network is the all-ascending
// formulation (first step of a merge compares i with its mirror image, the others i with i + j); positions >= n behave as
// +infinity, never move, and their compare-exchanges are simply skipped.  Four independent pairs per thread are loaded
// before any is written back, so that a sort in global memory (L2) has several loads in flight per thread.
template <int SORT_ILP>
__device__ __forceinline__ void sort_step(unsigned long long* keys, int n, int halfP, int tid, int nthreads, int lk, int j /*0: flip step*/) {
    const int k = 1 << lk, half = k >> 1;
    for (int t0 = tid; t0 < halfP; t0 += nthreads * SORT_ILP) {
        int ii[SORT_ILP], ll[SORT_ILP];
        unsigned long long x[SORT_ILP], y[SORT_ILP];
#pragma unroll
        for (int u = 0; u < SORT_ILP; ++u) {
            const int t = t0 + u * nthreads;
            if (j == 0) {
                const int blk = t >> (lk - 1), o = t & (half - 1);
                ii[u] = (blk << lk) + o;
                ll[u] = (blk << lk) + (k - 1 - o);
            } else {
                ii[u] = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                ll[u] = ii[u] | j;
            }
            if (t >= halfP || ll[u] >= n) ll[u] = -1;
            if (ll[u] >= 0) { x[u] = keys[ii[u]]; y[u] = keys[ll[u]]; }
        }
#pragma unroll
        for (int u = 0; u < SORT_ILP; ++u)
            if (ll[u] >= 0 && x[u] > y[u]) { keys[ii[u]] = y[u]; keys[ll[u]] = x[u]; }
    }
}
template <int SORT_ILP>
__device__ __forceinline__ void cta_sort_keys(unsigned long long* keys, int n, int tid, int nthreads) {
    int P = 1, logP = 0;
    while (P < n) { P <<= 1; ++logP; }
    for (int lk = 1; lk <= logP; ++lk) {
        sort_step<SORT_ILP>(keys, n, P >> 1, tid, nthreads, lk, 0);
        __syncthreads();
        for (int j = 1 << (lk - 2 >= 0 ? lk - 2 : 0); lk >= 2 && j > 0; j >>= 1) {
            sort_step<SORT_ILP>(keys, n, P >> 1, tid, nthreads, lk, j);
            __syncthreads();
        }
    }
}

// Best-first selection for a large segment: copies the smallest keys of keys[0, n) -- at least `want` of them (or all n),
// at most `cap` -- into out[] (unsorted) and returns their number.  Radix select on the key bits below `top_bit` (the bits
// above are the class, identical inside a segment), 8 bits per pass from the top; a pass is a scan of the whole segment
// (L2 resident) with 8 loads in flight per thread.  Stops as soon as the bucket boundary gives a count in [want, cap].
#define SEL_ILP 4
__device__ __forceinline__ int select_best_keys(const unsigned long long* __restrict__ keys, int n, unsigned long long* out, int want, int cap,
                                int top_bit, unsigned long long class_prefix, int tid, int nthreads, unsigned* s_hist /*[256]*/,
                                unsigned long long* s_u64 /*[2]*/, int* s_int /*[4]*/) {
    int hi = top_bit;                           // bits [0, hi) are still undecided inside the boundary bucket
    unsigned long long prefix = class_prefix;   // decided high bits of the boundary bucket
    int below = 0;                              // keys known to be smaller than every key of the boundary bucket
    unsigned long long T = ~0ull;               // final threshold: select keys < T
    while (true) {
        const int bits = hi < 8 ? hi : 8, shift = hi - bits;
        for (int i = tid; i < 256; i += nthreads) s_hist[i] = 0u;
        __syncthreads();
        for (int base = 0; base < n; base += nthreads * SEL_ILP) {      // uniform trip count: the lanes vote below
            unsigned long long k[SEL_ILP];
#pragma unroll
            for (int u = 0; u < SEL_ILP; ++u) { const int i = base + tid + u * nthreads; k[u] = i < n ? keys[i] : 0ull; }
#pragma unroll
            for (int u = 0; u < SEL_ILP; ++u) {
                const int i = base + tid + u * nthreads;
                const bool in = i < n && (k[u] >> hi) == (prefix >> hi);
                const unsigned digit = in ? ((unsigned)(k[u] >> shift) & ((1u << bits) - 1u)) : 0xFFFFFFFFu;
                // lanes with the same digit share one shared-memory atomic (the leading score bits take few distinct values)
                const unsigned same = __match_any_sync(0xffffffffu, digit);
                if (in && (threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&s_hist[digit], (unsigned)__popc(same));
            }
        }
        __syncthreads();
        if (tid == 0) {
            int lo = below, b = 0;
            const int nb = 1 << bits;
            for (; b < nb - 1; ++b) {
                if (lo + (int)s_hist[b] >= want) break;
                lo += (int)s_hist[b];
            }
            const int upto = lo + (int)s_hist[b];                  // keys < prefix | (b+1) << shift
            if (upto <= cap || shift == 0) {                        // shift == 0: buckets hold single keys (keys are unique)
                s_int[0] = 1;
                s_u64[0] = (b == nb - 1 && shift + bits >= 64) ? ~0ull : (prefix | ((unsigned long long)(b + 1) << shift));
                if (b == nb - 1) {                                          // whole boundary bucket: threshold = next prefix
                    const unsigned long long next = ((prefix >> hi) + 1ull) << hi;
                    s_u64[0] = next > prefix ? next : ~0ull;                     // (wrap-around: nothing lies above)
                }
            } else {
                s_int[0] = 0;
                s_int[1] = lo;
                s_u64[0] = prefix | ((unsigned long long)b << shift);
            }
        }
        __syncthreads();
        const int done = s_int[0];
        if (done) { T = s_u64[0]; break; }
        below = s_int[1];
        prefix = s_u64[0];
        hi = shift;
        __syncthreads();
    }
    // compaction (order does not matter: the caller sorts)
    if (tid == 0) s_int[2] = 0;
    __syncthreads();
    for (int i0 = tid; i0 < n; i0 += nthreads * SEL_ILP) {
        unsigned long long k[SEL_ILP];
#pragma unroll
        for (int u = 0; u < SEL_ILP; ++u) { const int i = i0 + u * nthreads; k[u] = i < n ? keys[i] : ~0ull; }
#pragma unroll
        for (int u = 0; u < SEL_ILP; ++u) {
            const int i = i0 + u * nthreads;
            if (i < n && k[u] < T) {
                const int pos = atomicAdd(&s_int[2], 1);
                if (pos < cap) out[pos] = k[u];
            }
        }
    }
    __syncthreads();
    const int m = s_int[2];
    __syncthreads();
    return m < cap ? m : cap;
}

// ---------------------------------------------------------------------------------------------- 3. NMS
// Where the four box codes of (image b, anchor a) live: the anchor-major tensor [B,A,4], or the per-level head tensors.
struct CodeView {
    const float4* flat;      // [B,A,4] or nullptr (then `head` is used)
    HeadGeom head;
};
__device__ __forceinline__ float4 load_code(const CodeView& cv, int b, long long A, int a) {
    if (cv.flat) return cv.flat[(size_t)b * A + a];
    return head_load_code(cv.head, b, a);
}

struct NmsBox {          // corners min/max-normalised as NonMaxSuppressionV3 does; area <= 0 never suppresses
    float ymin, xmin, ymax, xmax;
};

// IoU(a, b) > thr with the exact float32 semantics of  inter / (area_a + area_b - inter) > thr  (IEEE divide).
// nms_fast decides without the division whenever |inter - thr*union| > 2^-18 * |thr| * union (both roundings
// involved are below 2^-23 relative, so outside that band the comparison is certain) and flags the rest as
// ambiguous; nms_exact is the division.  Splitting the two lets four independent tests be in flight per lane.
__device__ __forceinline__ void nms_fast(const NmsBox a, float area_a, const NmsBox b, float area_b, float thr, float band,
                                         bool& yes, bool& ambiguous) {
    const float ih = fmaxf(f_sub(fminf(a.ymax, b.ymax), fmaxf(a.ymin, b.ymin)), 0.0f);
    const float iw = fmaxf(f_sub(fminf(a.xmax, b.xmax), fmaxf(a.xmin, b.xmin)), 0.0f);
    const float inter = f_mul(ih, iw);
    const float uni = f_sub(f_add(area_a, area_b), inter);
    const bool valid = (area_a > 0.0f) && (area_b > 0.0f);      // NonMaxSuppressionV3: area <= 0 never suppresses
    const float d = fmaf(-thr, uni, inter);
    const float m = band * uni;
    yes = valid && (d > m);
    ambiguous = valid && (fabsf(d) <= m);
}
__device__ __forceinline__ bool nms_exact(const NmsBox a, float area_a, const NmsBox b, float area_b, float thr) {
    const float ih = fmaxf(f_sub(fminf(a.ymax, b.ymax), fmaxf(a.ymin, b.ymin)), 0.0f);
    const float iw = fmaxf(f_sub(fminf(a.xmax, b.xmax), fmaxf(a.xmin, b.xmin)), 0.0f);
    const float inter = f_mul(ih, iw);
    return f_div(inter, f_sub(f_add(area_a, area_b), inter)) > thr;
}

// Segments with at most 32 candidates (the vast majority: background classes) are resolved by ONE WARP each, one
// candidate per lane: decode, a 32x32 suppression bit matrix (one column per lane), a greedy walk with shuffles.
// Larger segments are pushed to a queue for nms_kernel, which spends a whole CTA on each of them.
#define NMS_SMALL_WARPS 8
template <bool DECODED>
__global__ void __launch_bounds__(NMS_SMALL_WARPS * 32) nms_small_kernel(
    const unsigned long long* __restrict__ cand, long long capc, KeyFormat fmt, const int* __restrict__ seg_count,
    const CodeView codes, const float4* __restrict__ anchors, long long A,
    long long nseg, int C, int K, float iou_thr, float4* __restrict__ seg_box, float* __restrict__ seg_score,
    int* __restrict__ seg_anchor, int* __restrict__ seg_kept, int* __restrict__ heavy_queue, int* __restrict__ heavy_count) {
    __shared__ NmsBox s_tile[NMS_SMALL_WARPS][32];
    __shared__ float s_tile_area[NMS_SMALL_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long seg = (long long)blockIdx.x * NMS_SMALL_WARPS + warp;
    if (seg >= nseg) return;
    const int n = seg_count[seg];
    if (n <= 0) {
        if (lane == 0) seg_kept[seg] = 0;
        return;
    }
    if (n > 32) {                                      // heavy_count[0] / [1]: the two queues (heavy | large), nseg entries each
        if (lane == 0) {
            const int large = n > NMS_SORT_SMEM_KEYS ? 1 : 0;
            heavy_queue[(size_t)large * nseg + atomicAdd(heavy_count + large, 1)] = (int)seg;
        }
        return;
    }
    const int b = (int)(seg / C);
    const size_t obase = (size_t)seg * K;
    const float band = fabsf(iou_thr) * 3.814697265625e-06f;
    bool alive = lane < n;
    NmsBox box = {0.f, 0.f, 0.f, 0.f};
    float4 raw = make_float4(0.f, 0.f, 0.f, 0.f);
    float area = 0.f, score = 0.f;
    int a = 0;
    // the segment's keys arrive in arbitrary order: sort them (score descending, anchor ascending)
    const unsigned long long key = warp_sort_keys(alive ? cand[(size_t)seg * capc + lane] : ~0ull, lane);
    if (alive) {
        a = key_anchor(key, fmt);
        score = key_score(key, fmt);
        raw = load_code(codes, b, A, a);
        if (!DECODED) raw = box_clip01(box_decode(raw, anchors[a]));                         // nms.py:76-77
        box.ymin = fminf(raw.x, raw.z); box.xmin = fminf(raw.y, raw.w);
        box.ymax = fmaxf(raw.x, raw.z); box.xmax = fmaxf(raw.y, raw.w);
        area = f_mul(f_sub(box.ymax, box.ymin), f_sub(box.xmax, box.xmin));
    }
    s_tile[warp][lane] = box;
    s_tile_area[warp][lane] = area;
    __syncwarp();
    // column of the suppression matrix: earlier lanes t that suppress this lane
    unsigned col = 0u;
    if (n > 1) {
        for (int t = 0; t < n - 1; ++t) {
            bool y, am;
            nms_fast(box, area, s_tile[warp][t], s_tile_area[warp][t], iou_thr, band, y, am);
            if (am) y = nms_exact(box, area, s_tile[warp][t], s_tile_area[warp][t], iou_thr);
            if (y && t < lane) col |= 1u << t;
        }
    }
    // greedy walk in score order
    unsigned keep = 0u;
    int room = K;
    for (unsigned todo = __ballot_sync(0xffffffffu, alive); todo && room > 0;) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        const unsigned cj = __shfl_sync(0xffffffffu, col, j);
        if (!(cj & keep)) { keep |= 1u << j; --room; }
    }
    if ((keep >> lane) & 1u) {
        const int pos = __popc(keep & ((1u << lane) - 1u));
        seg_box[obase + pos] = raw;
        seg_score[obase + pos] = score;
        seg_anchor[obase + pos] = a;
    }
    if (lane == 0) seg_kept[seg] = __popc(keep);
}

// One CTA per queued (image, class) segment, NMS_CH candidates per chunk in sorted order, NMS_TPC threads per candidate:
//   (a) every candidate is decoded + clipped and tested against the boxes kept from earlier chunks (shared memory);
//       the threads of a candidate split the kept list and OR their verdicts with shuffles;
//   (b) column c of the chunk's suppression bit matrix (bit t set <=> t < c, t survived (a), IoU(t, c) > threshold) is
//       built the same way, the threads splitting t;
//   (c) warp 0 resolves the greedy order WITHOUT walking it: a candidate is removed as soon as one of its suppressors is
//       known to be kept, and kept as soon as all of them are known to be removed; every round decides at least the first
//       undecided candidate, in practice a handful of rounds of ballots decide all 64;  the result is cut at K;
//   (d) kept candidates append themselves (rank by popcount) to the kept list and to the segment's output.
// LARGE = false: queued segments with 33..NMS_SORT_SMEM_KEYS candidates (sorted in shared memory); LARGE = true: the separate
// queue of larger segments (dense scores), which need the select / in-place sort machinery and many more registers.
template <bool DECODED, bool LARGE>
__global__ void __launch_bounds__(NMS_THREADS) nms_kernel(
    unsigned long long* __restrict__ cand, long long capc, KeyFormat fmt, const int* __restrict__ seg_count,
    const CodeView codes, const float4* __restrict__ anchors, long long A,
    long long nseg, int C, int K, float iou_thr, float4* __restrict__ seg_box, float* __restrict__ seg_score,
    int* __restrict__ seg_anchor, int* __restrict__ seg_kept, const int* __restrict__ heavy_queue,
    const int* __restrict__ heavy_count) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    __shared__ NmsBox s_tile[NMS_CH];
    __shared__ float s_tile_area[NMS_CH];
    __shared__ unsigned long long s_col[NMS_CH];
    __shared__ unsigned s_alive32[2][2];               // [chunk parity][word]: survivors of (a), set with atomicOr
    __shared__ unsigned long long s_keep;
    __shared__ unsigned s_hist[256];                   // select_best_keys scratch
    __shared__ unsigned long long s_sel64[2];
    __shared__ int s_sel32[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long* s_sort = (unsigned long long*)nms_smem;     // [NMS_SORT_SMEM_KEYS]
    NmsBox* s_kept = (NmsBox*)(s_sort + NMS_SORT_SMEM_KEYS);        // [K]
    float* s_kept_area = (float*)(s_kept + K);                      // [K]
    const float band = fabsf(iou_thr) * 3.814697265625e-06f;
    const int c = tid / NMS_TPC, q = tid % NMS_TPC;    // candidate slot in the chunk, position among its threads
    const int nheavy = *heavy_count;
    for (int item = blockIdx.x; item < nheavy; item += gridDim.x) {
        const long long seg = heavy_queue[item];
        const int n = seg_count[seg];
        const int b = (int)(seg / C);
        unsigned long long* keys = cand + (size_t)seg * capc;
        const size_t obase = (size_t)seg * K;
        int kept = 0;
        // ---- sort the segment (score descending, anchor ascending).  Up to NMS_SORT_SMEM_KEYS keys: in shared memory.
        //      Larger segments (dense scores): attempt 0 selects the best <= NMS_SORT_SMEM_KEYS keys with a radix select and
        //      sorts only those -- greedy NMS truncated at K normally finishes inside them; if it does not (kept < K with
        //      candidates left) attempt 1 sorts the whole segment in place and starts over.
      for (int attempt = 0;; ++attempt) {
        const unsigned long long* sorted = s_sort;
        int navail = n;
        __syncthreads();
        if (!LARGE) {
            for (int i = tid; i < n; i += NMS_THREADS) s_sort[i] = keys[i];
            __syncthreads();
            cta_sort_keys<1>(s_sort, n, tid, NMS_THREADS);
        } else if (attempt == 0) {
            navail = select_best_keys(keys, n, s_sort, NMS_SORT_SMEM_KEYS / 2, NMS_SORT_SMEM_KEYS, fmt.cshift,
                                      (unsigned long long)(seg % C) << fmt.cshift, tid, NMS_THREADS, s_hist, s_sel64, s_sel32);
            cta_sort_keys<1>(s_sort, navail, tid, NMS_THREADS);
        } else {
            cta_sort_keys<4>(keys, n, tid, NMS_THREADS);
            sorted = keys;
        }
        kept = 0;
        if (tid < 4) s_alive32[tid >> 1][tid & 1] = 0u;
        __syncthreads();

        // the candidate of the NEXT chunk (key -> code -> anchor: dependent global loads) is fetched while the current
        // chunk is processed
        unsigned long long nkey = 0ull;
        float4 ncode = make_float4(0.f, 0.f, 0.f, 0.f), nanc = make_float4(0.f, 0.f, 0.f, 0.f);
        auto fetch = [&](int i) {
            if (i < navail) {
                nkey = sorted[i];
                const int na = key_anchor(nkey, fmt);
                ncode = load_code(codes, b, A, na);
                if (!DECODED) nanc = anchors[na];
            }
        };
        fetch(c);
        int parity = 0;
        for (int base = 0; base < navail && kept < K; base += NMS_CH, parity ^= 1) {
            const int i = base + c;
            bool alive = i < navail;
            NmsBox box = {0.f, 0.f, 0.f, 0.f};
            float4 raw = make_float4(0.f, 0.f, 0.f, 0.f);
            float area = 0.f, score = 0.f;
            int a = 0;
            if (alive) {
                a = key_anchor(nkey, fmt);
                score = key_score(nkey, fmt);
                if (DECODED) raw = ncode;
                else raw = box_clip01(box_decode(ncode, nanc));                              // nms.py:76-77
                box.ymin = fminf(raw.x, raw.z); box.xmin = fminf(raw.y, raw.w);
                box.ymax = fmaxf(raw.x, raw.z); box.xmax = fmaxf(raw.y, raw.w);
                area = f_mul(f_sub(box.ymax, box.ymin), f_sub(box.xmax, box.xmin));
            }
            fetch(i + NMS_CH);
            // (a) against the boxes kept from earlier chunks
            int hit = 0;
            if (alive) {
                for (int j = q; j < kept; j += NMS_TPC) {
                    bool y, am;
                    nms_fast(box, area, s_kept[j], s_kept_area[j], iou_thr, band, y, am);
                    if (am) y = nms_exact(box, area, s_kept[j], s_kept_area[j], iou_thr); // rare: within 2^-18 of the threshold
                    hit |= y ? 1 : 0;
                }
            }
#pragma unroll
            for (int o = 1; o < NMS_TPC; o <<= 1) hit |= __shfl_xor_sync(0xffffffffu, hit, o);
            alive = alive && !hit;
            if (q == 0) {
                s_tile[c] = box;
                s_tile_area[c] = area;
                if (alive) atomicOr(&s_alive32[parity][c >> 5], 1u << (c & 31));
            }
            if (tid < 2) s_alive32[parity ^ 1][tid] = 0u;      // the buffer of the next chunk (last read two barriers ago)
            __syncthreads();
            const unsigned long long alive_mask = ((unsigned long long)s_alive32[parity][1] << 32) | s_alive32[parity][0];
            // (b) column of the chunk's suppression matrix
            unsigned long long col = 0ull;
            if (alive && (alive_mask & (alive_mask - 1ull))) {               // at least two survivors in the chunk
                for (int t = q; t < c; t += NMS_TPC) {
                    if ((alive_mask >> t) & 1ull) {
                        bool y, am;
                        nms_fast(box, area, s_tile[t], s_tile_area[t], iou_thr, band, y, am);
                        if (am) y = nms_exact(box, area, s_tile[t], s_tile_area[t], iou_thr);
                        if (y) col |= 1ull << t;
                    }
                }
            }
#pragma unroll
            for (int o = 1; o < NMS_TPC; o <<= 1) col |= __shfl_xor_sync(0xffffffffu, col, o);
            if (q == 0) s_col[c] = col;
            __syncthreads();
            // (c) greedy order by resolution rounds (warp 0: lane l owns candidates l and l + 32)
            if (warp == 0) {
                const unsigned long long col_lo = s_col[lane], col_hi = s_col[lane + 32];
                unsigned long long keep = 0ull, gone = ~alive_mask, und = alive_mask;
                while (und) {
                    int st_lo = 0, st_hi = 0;                                // 1 = kept, 2 = removed
                    if ((und >> lane) & 1ull) st_lo = (col_lo & keep) ? 2 : ((col_lo & ~gone) == 0ull ? 1 : 0);
                    if ((und >> (lane + 32)) & 1ull) st_hi = (col_hi & keep) ? 2 : ((col_hi & ~gone) == 0ull ? 1 : 0);
                    const unsigned long long k_new = ((unsigned long long)__ballot_sync(0xffffffffu, st_hi == 1) << 32) | __ballot_sync(0xffffffffu, st_lo == 1);
                    const unsigned long long g_new = ((unsigned long long)__ballot_sync(0xffffffffu, st_hi == 2) << 32) | __ballot_sync(0xffffffffu, st_lo == 2);
                    keep |= k_new;
                    gone |= g_new;
                    und &= ~(k_new | g_new);
                }
                int extra = __popcll(keep) - (K - kept);                      // at most K in total: drop the lowest-scored
                while (extra > 0) { keep &= ~(1ull << (63 - __clzll((long long)keep))); --extra; }
                if (lane == 0) s_keep = keep;
            }
            __syncthreads();
            const unsigned long long keep = s_keep;
            // (d) append the kept candidates in score order
            if (q == 0 && ((keep >> c) & 1ull)) {
                const int pos = kept + __popcll(keep & ((1ull << c) - 1ull));
                s_kept[pos] = box;
                s_kept_area[pos] = area;
                seg_box[obase + pos] = raw;
                seg_score[obase + pos] = score;
                seg_anchor[obase + pos] = a;
            }
            kept += __popcll(keep);
            __syncthreads();
        }
        if (!LARGE || navail == n || kept >= K) break;          // uniform over the CTA; otherwise: full sort and start over
      }
        if (tid == 0) seg_kept[seg] = kept;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- 4. pack
#define PACK_SLOTS_PER_BLOCK 1024
__global__ void __launch_bounds__(256) pack_kernel(const float4* __restrict__ seg_box, const float* __restrict__ seg_score,
                                                   const int* __restrict__ seg_anchor, const int* __restrict__ seg_kept,
                                                   int C, int K, float4* __restrict__ out_boxes, float* __restrict__ out_scores,
                                                   int* __restrict__ out_classes, int* __restrict__ out_num,
                                                   int* __restrict__ out_anchor, const float4* __restrict__ box_scaler,
                                                   float final_thr, const int* __restrict__ large_count, int* hint_out) {
    extern __shared__ int s_off[];   // [C+1] exclusive prefix sums of the per-class kept counts, then [C] the counts
    int* kept = s_off + C + 1;
    const int b = blockIdx.y;
    // density hint for the NEXT call: number of segments that needed the large-segment path, posted to mapped host memory
    if (hint_out && blockIdx.x == 0 && b == 0 && threadIdx.x == 0) *hint_out = *large_count;
    // Post-path consumers folded in (model.py:67-68, inference/detector.py:54-58): boxes /= box_scaler[b], and a final
    // `scores > final_thr` filter.  Inside a class the kept scores are descending, so the survivors of that filter are a
    // prefix of every class segment and the class-major order is preserved, exactly as the reference's boolean mask does.
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        int k = seg_kept[(size_t)b * C + c];
        if (final_thr > -INFINITY) {
            const float* sc = seg_score + ((size_t)b * C + c) * K;
            while (k > 0 && !(sc[k - 1] > final_thr)) --k;
        }
        kept[c] = k;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int per = (C + 31) / 32;
        const int c0 = min(lane * per, C), c1 = min(c0 + per, C);
        int mine = 0;
        for (int c = c0; c < c1; ++c) mine += kept[c];
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int run = incl - mine;
        for (int c = c0; c < c1; ++c) { s_off[c] = run; run += kept[c]; }
        if (lane == 31) {
            s_off[C] = incl;
            if (blockIdx.x == 0) out_num[b] = incl;                  // nms.py:81
        }
    }
    __syncthreads();
    const int total = s_off[C];
    const size_t M = (size_t)C * K;
    float4* ob = out_boxes + b * M;
    float* os = out_scores + b * M;
    int* oc = out_classes + b * M;
    int* oa = out_anchor ? out_anchor + b * M : nullptr;
    const int lo = blockIdx.x * PACK_SLOTS_PER_BLOCK;
    const int hi = min(lo + PACK_SLOTS_PER_BLOCK, C * K);
    for (int idx = lo + threadIdx.x; idx < hi; idx += blockDim.x) {
        const int c = idx / K, j = idx - c * K;
        if (j < kept[c]) {                                            // selected entries, class-major (nms.py:42-44)
            const size_t src = ((size_t)b * C + c) * K + j;
            const int dst = s_off[c] + j;
            float4 bx = seg_box[src];
            if (box_scaler) {
                const float4 sc4 = box_scaler[b];
                bx = make_float4(f_div(bx.x, sc4.x), f_div(bx.y, sc4.y), f_div(bx.z, sc4.z), f_div(bx.w, sc4.w));
            }
            ob[dst] = bx;
            os[dst] = seg_score[src];
            oc[dst] = c;
            if (oa) oa[dst] = seg_anchor[src];
        }
        if (idx >= total) {                                           // zero padding (nms.py:84-89)
            ob[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
            os[idx] = 0.f;
            oc[idx] = 0;
            if (oa) oa[idx] = -1;
        }
    }
}

// ---------------------------------------------------------------------------------------------- host side
static int bits_for(long long n) {   // bits needed to represent values in [0, n)
    int b = 1;
    while ((1ll << b) < n) ++b;
    return b;
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

static int postprocess_impl(ssdk_ctx* ctx, const HeadGeom* head, const float* codes, const float* anchors, const float* scores, int flags,
                            int B, int64_t A, int C, double score_threshold, double iou_threshold, int K,
                            const float* box_scaler, double final_score_threshold,
                            float* out_boxes, float* out_scores, int32_t* out_classes, int32_t* out_num,
                            int32_t* out_anchor_idx) {
    SSDK_TRY(ssdk_ctx_enter(ctx));
    SSDK_REQUIRE(B >= 0 && A >= 0 && C > 0 && K > 0, SSDK_ERR_ARG, "ssdk_postprocess: bad sizes (B=%d A=%lld C=%d K=%d)", B,
                 (long long)A, C, K);
    SSDK_REQUIRE(B <= 65535, SSDK_ERR_SHAPE, "ssdk_postprocess: batch %d > 65535", B);
    // OP_REQUIRES of TensorFlow's NonMaxSuppressionV3 kernel (the op behind detector/utils/nms.py:33)
    SSDK_REQUIRE(iou_threshold >= 0.0 && iou_threshold <= 1.0, SSDK_ERR_ARG, "ssdk_postprocess: iou_threshold must be in [0, 1] (got %g)",
                 iou_threshold);
    if (B == 0) return SSDK_OK;
    SSDK_REQUIRE(out_boxes && out_scores && out_classes && out_num, SSDK_ERR_ARG, "ssdk_postprocess: null output");
    const bool decoded = (flags & SSDK_BOXES_DECODED) != 0;
    const bool is_logits = (flags & SSDK_INPUT_LOGITS) != 0;
    SSDK_REQUIRE(A == 0 || head || (codes && scores), SSDK_ERR_ARG, "ssdk_postprocess: null input");
    SSDK_REQUIRE(A == 0 || decoded || anchors, SSDK_ERR_ARG, "ssdk_postprocess: null anchors");
    SSDK_REQUIRE(!(head && decoded), SSDK_ERR_ARG, "ssdk_head_detect: head tensors hold encoded boxes");
    SSDK_REQUIRE(aligned16(codes) && aligned16(anchors) && aligned16(out_boxes), SSDK_ERR_SHAPE,
                 "ssdk_postprocess: box arrays must be 16-byte aligned");
    SSDK_REQUIRE(((uintptr_t)scores & 3) == 0, SSDK_ERR_SHAPE, "ssdk_postprocess: scores must be 4-byte aligned");
    const long long per_image = (long long)A * C;
    SSDK_REQUIRE(per_image < (1ll << 31), SSDK_ERR_SHAPE, "ssdk_postprocess: A*C must be < 2^31");
    SSDK_REQUIRE((size_t)K * 20 <= 160 * 1024, SSDK_ERR_SHAPE,
                 "ssdk_postprocess: max_boxes_per_class %d too large", K);
    KeyFormat fmt;
    fmt.abits = bits_for(A > 1 ? A : 2);
    fmt.cshift = 32 + fmt.abits;
    SSDK_REQUIRE(fmt.cshift + bits_for(C > 1 ? C : 2) <= 64, SSDK_ERR_SHAPE, "ssdk_postprocess: A=%lld x C=%d does not fit the key",
                 (long long)A, C);

    // workspace: one candidate region of A keys per (image, class) segment (a class cannot have more candidates than
    // anchors, so the regions cannot overflow; only the filled prefixes are ever touched), one counter per segment
    // (zeroed every call), per-segment NMS results
    const long long capc = A > 0 ? A : 1;
    const long long nseg = (long long)B * C;
    SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_cand, (size_t)nseg * capc * sizeof(unsigned long long)));
    const size_t n_int = 4 + 4 * (size_t)nseg;                         // queue counts (+pad), seg_count, seg_kept, heavy queue, large queue
    SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_counts, n_int * sizeof(int)));
    const size_t seg_elems = (size_t)nseg * K;
    SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_seg, seg_elems * (sizeof(float4) + sizeof(float) + sizeof(int))));
    unsigned long long* cand = (unsigned long long*)ctx->ws_cand.p;
    int* heavy_count = (int*)ctx->ws_counts.p;
    int* seg_count = heavy_count + 4;
    int* seg_kept = seg_count + nseg;
    int* heavy_queue = seg_kept + nseg;
    float4* seg_box = (float4*)ctx->ws_seg.p;
    float* seg_score = (float*)(seg_box + seg_elems);
    int* seg_anchor = (int*)(seg_score + seg_elems);
    SSDK_CHECK_CUDA(cudaMemsetAsync(heavy_count, 0, (4 + (size_t)nseg) * sizeof(int), ctx->stream));
    if (per_image == 0) SSDK_CHECK_CUDA(cudaMemsetAsync(seg_kept, 0, (size_t)nseg * sizeof(int), ctx->stream));

    const float thr = (float)score_threshold;
    if (per_image > 0) {
        // 1. filter
        float x_lo = -INFINITY;
        if (is_logits) {
            if (score_threshold >= 1.0) x_lo = INFINITY;
            else if (score_threshold > 0.0) {
                const double lg = log(score_threshold / (1.0 - score_threshold));
                x_lo = (float)(lg - 1e-3 * (1.0 + fabs(lg)));
            }
        }
        if (head) {
            long long most = 0;                                           // floats of the largest level
            for (int l = 0; l < head->num_levels; ++l) {
                const long long cnt = (long long)B * head->per_loc * C * head->hw[l];
                if (cnt > most) most = cnt;
            }
            long long chunks = (most / 4 + FILTER_THREADS * FILTER_UNROLL - 1) / (FILTER_THREADS * FILTER_UNROLL);
            long long gx = (long long)ctx->num_sms * 16;
            if (gx > chunks) gx = chunks;
            if (gx < 1) gx = 1;
            const dim3 hgrid_f((unsigned)gx, head->num_levels);
            SSDK_KERNEL(ctx, SSDK_K_FILTER,
                if (is_logits)
                    head_filter_kernel<true><<<hgrid_f, FILTER_THREADS, 0, ctx->stream>>>(*head, B, thr, x_lo, fmt, cand, capc, seg_count);
                else
                    head_filter_kernel<false><<<hgrid_f, FILTER_THREADS, 0, ctx->stream>>>(*head, B, thr, x_lo, fmt, cand, capc, seg_count));
        } else {
            long long chunks = (per_image / 4 + FILTER_THREADS * FILTER_UNROLL - 1) / (FILTER_THREADS * FILTER_UNROLL);
            long long gx = ((long long)ctx->num_sms * 16 + B - 1) / B;
            if (gx > chunks) gx = chunks;
            if (gx < 1) gx = 1;
            const dim3 fgrid((unsigned)gx, B);
            // dense scores last time on this context (hint posted by pack_kernel; stale or missing is fine, both kernels are
            // correct for any input): CTA-aggregated append
            const bool dense = ctx->hint_host && ((volatile int*)ctx->hint_host)[0] > 0 && C <= 4096 && per_image < (1ll << 31) - 8;
            if (getenv("SSDK_FILTER_DENSE") ? atoi(getenv("SSDK_FILTER_DENSE")) != 0 : dense) {
                const size_t dsmem = 2 * (size_t)C * sizeof(int);
                SSDK_KERNEL(ctx, SSDK_K_FILTER,
                    if (is_logits)
                        filter_dense_kernel<true><<<fgrid, FILTER_THREADS, dsmem, ctx->stream>>>(scores, per_image, C, thr, x_lo, fmt, cand, capc, seg_count);
                    else
                        filter_dense_kernel<false><<<fgrid, FILTER_THREADS, dsmem, ctx->stream>>>(scores, per_image, C, thr, x_lo, fmt, cand, capc, seg_count));
            } else
            SSDK_KERNEL(ctx, SSDK_K_FILTER,
                if (is_logits)
                    filter_kernel<true><<<fgrid, FILTER_THREADS, 0, ctx->stream>>>(scores, per_image, C, thr, x_lo, fmt, cand, capc, seg_count);
                else
                    filter_kernel<false><<<fgrid, FILTER_THREADS, 0, ctx->stream>>>(scores, per_image, C, thr, x_lo, fmt, cand, capc, seg_count));
        }

        // 2.-3. NMS: one warp per small segment (<= 32 candidates, sorted with shuffles), then one CTA per queued large
        //       segment (sorted by the CTA)
        const size_t nms_smem = (size_t)NMS_SORT_SMEM_KEYS * sizeof(unsigned long long) + (size_t)K * (sizeof(NmsBox) + sizeof(float));
        const int sgrid_nms = ceil_div_i(nseg, NMS_SMALL_WARPS);
        long long hgrid = (long long)ctx->num_sms * 4;
        if (hgrid > nseg) hgrid = nseg;
        CodeView c4;
        c4.flat = head ? nullptr : (const float4*)codes;
        if (head) c4.head = *head;
        else memset(&c4.head, 0, sizeof(c4.head));
        const float4* a4 = (const float4*)anchors;
        const float iou_f = (float)iou_threshold;
        const int nms_slot = ctx->profiling ? ssdk_prof_begin(ctx, SSDK_K_NMS) : -1;
        // segments with more than NMS_SORT_SMEM_KEYS candidates can only exist when A is that large
        const bool may_be_large = A > NMS_SORT_SMEM_KEYS;
        long long lgrid = (long long)ctx->num_sms;
        if (lgrid > nseg) lgrid = nseg;
#define SSDK_LAUNCH_NMS(DEC)                                                                                                  \
        do {                                                                                                                  \
            SSDK_TRY(ssdk_set_max_smem(ctx, (const void*)nms_kernel<DEC, false>, (int)nms_smem));                            \
            nms_small_kernel<DEC><<<sgrid_nms, NMS_SMALL_WARPS * 32, 0, ctx->stream>>>(                                      \
                cand, capc, fmt, seg_count, c4, a4, A, nseg, C, K, iou_f, seg_box, seg_score, seg_anchor, seg_kept, heavy_queue, \
                heavy_count);                                                                                                 \
            nms_kernel<DEC, false><<<(int)hgrid, NMS_THREADS, nms_smem, ctx->stream>>>(                                      \
                cand, capc, fmt, seg_count, c4, a4, A, nseg, C, K, iou_f, seg_box, seg_score, seg_anchor, seg_kept, heavy_queue, \
                heavy_count);                                                                                                 \
            if (may_be_large) {                                                                                               \
                SSDK_TRY(ssdk_set_max_smem(ctx, (const void*)nms_kernel<DEC, true>, (int)nms_smem));                         \
                nms_kernel<DEC, true><<<(int)lgrid, NMS_THREADS, nms_smem, ctx->stream>>>(                                   \
                    cand, capc, fmt, seg_count, c4, a4, A, nseg, C, K, iou_f, seg_box, seg_score, seg_anchor, seg_kept,       \
                    heavy_queue + nseg, heavy_count + 1);                                                                     \
                ctx->launches++;                                                                                              \
            }                                                                                                                 \
        } while (0)
        if (decoded) SSDK_LAUNCH_NMS(true);
        else SSDK_LAUNCH_NMS(false);
#undef SSDK_LAUNCH_NMS
        if (nms_slot >= 0) ssdk_prof_end(ctx, nms_slot);
        ctx->launches++;
        SSDK_CHECK_LAUNCH(ctx);
    }
    // 4. pack
    SSDK_KERNEL(ctx, SSDK_K_PACK,
                pack_kernel<<<dim3(ceil_div_i((long long)C * K, PACK_SLOTS_PER_BLOCK), B), 256, (size_t)(2 * C + 1) * sizeof(int), ctx->stream>>>(
                    seg_box, seg_score, seg_anchor, seg_kept, C, K, (float4*)out_boxes, out_scores, out_classes, out_num,
                    out_anchor_idx, (const float4*)box_scaler, (float)final_score_threshold, heavy_count + 1,
                    (per_image > 0 && !head) ? ctx->hint_dev : nullptr));
    return SSDK_OK;
}

extern "C" int ssdk_postprocess(ssdk_ctx* ctx, const float* codes, const float* anchors, const float* scores, int flags,
                                int B, int64_t A, int C, double score_threshold, double iou_threshold, int K,
                                float* out_boxes, float* out_scores, int32_t* out_classes, int32_t* out_num,
                                int32_t* out_anchor_idx) {
    return postprocess_impl(ctx, nullptr, codes, anchors, scores, flags, B, A, C, score_threshold, iou_threshold, K, nullptr,
                            -INFINITY, out_boxes, out_scores, out_classes, out_num, out_anchor_idx);
}

extern "C" int ssdk_detect(ssdk_ctx* ctx, const float* codes, const float* anchors, const float* scores, int flags, int B,
                           int64_t A, int C, double score_threshold, double iou_threshold, int K, const float* box_scaler,
                           double final_score_threshold, float* out_boxes, float* out_scores, int32_t* out_classes,
                           int32_t* out_num) {
    SSDK_REQUIRE(box_scaler == nullptr || aligned16(box_scaler), SSDK_ERR_SHAPE, "ssdk_detect: box_scaler must be 16-byte aligned");
    return postprocess_impl(ctx, nullptr, codes, anchors, scores, flags, B, A, C, score_threshold, iou_threshold, K, box_scaler,
                            final_score_threshold, out_boxes, out_scores, out_classes, out_num, nullptr);
}

extern "C" int ssdk_head_detect(ssdk_ctx* ctx, const ssdk_head* head, const float* anchors, int flags, int B, int64_t A, int C,
                                double score_threshold, double iou_threshold, int K, const float* box_scaler,
                                double final_score_threshold, float* out_boxes, float* out_scores, int32_t* out_classes,
                                int32_t* out_num, int32_t* out_anchor_idx) {
    SSDK_REQUIRE(box_scaler == nullptr || aligned16(box_scaler), SSDK_ERR_SHAPE, "ssdk_head_detect: box_scaler must be 16-byte aligned");
    HeadGeom G;
    SSDK_TRY(ssdk_head_geom(head, B, A, C, true, true, &G));
    return postprocess_impl(ctx, &G, nullptr, anchors, nullptr, flags & SSDK_INPUT_LOGITS, B, A, C, score_threshold, iou_threshold, K,
                            box_scaler, final_score_threshold, out_boxes, out_scores, out_classes, out_num, out_anchor_idx);
}
