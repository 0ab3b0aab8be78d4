// Inference post-processing: score threshold -> per-(image, class) candidate lists -> sort -> greedy NMS -> pack.
// Replaces detector/utils/nms.py:48-102 (batch_multiclass_non_max_suppression), :6-45
// (multiclass_non_max_suppression), the sigmoid of detector/ssd.py:60 and TensorFlow 1.12's NonMaxSuppressionV3
// (external C++ kernel called at nms.py:33; semantics restated in oracle/nms.py and pinned to TensorFlow's own unit-test
// vectors, tests/golden/tf_nms_vectors.py).
//
// Pipeline (all on the context's stream, no host synchronisation, the same launches whatever the data):
//   1. filter_kernel      streams the [B,A,C] scores (or logits) once with 128-bit no-allocate loads and appends a
//                         packed 64-bit key per (anchor, class) with score > threshold to the candidate list of its
//                         (image, class) SEGMENT: a BOUNDED region of SEG_CAP keys and one atomic counter per segment.
//                         This is the HBM-bound kernel.
//                         key = class | ~order(score) | anchor  ->  ascending u64 order == score descending, anchor
//                         index ascending inside a segment.
//                         The reference's `is_confident` anchor pre-filter (nms.py:71-74, '>=') only removes
//                         anchors that have no candidate at all (candidates need '>'), so it cannot change any
//                         output and is not materialised.
//   2. nms_small_kernel   one WARP per segment with at most 32 candidates (the vast majority): the keys are sorted with a
//                         shuffle bitonic network, decoded (box_utils.py:114-142) and clipped (nms.py:77), a 32x32
//                         suppression bit matrix is built and walked greedily.  Larger segments are queued.
//   3. nms_kernel         one CTA per queued segment that fits its region: bitonic sort of the keys in shared memory, then
//                         candidates are taken 64 at a time in sorted order (eight threads per candidate), tested against the
//                         boxes kept so far (shared memory) and against the chunk's earlier candidates (64x64 bit matrix);
//                         the greedy order is resolved with ballots; stops at K.
//   4. nms_rounds_kernel  segments with MORE candidates than a region holds (dense scores) -- the region's content is then
//                         an arbitrary subset and is discarded.  One persistent grid works in rounds separated by grid-wide
//                         barriers: a streaming pass histograms the not yet processed keys of every such segment over its
//                         current key range, a planning step picks the best-scored prefix of bins that fits a region
//                         (or narrows the range when a single bin is too big: a radix select over streaming passes), a
//                         second streaming pass collects exactly those keys, and the segment's greedy NMS CONTINUES over
//                         them with the boxes kept so far; repeat until K boxes are kept or no candidate is left.  Greedy
//                         NMS truncated at K normally ends inside the first ~1000 keys, i.e. after one round; the rounds
//                         make the result exact for ANY input with bounded memory.  Exits at once when no segment
//                         overflowed (the normal, sparse case).
//   5. pack_kernel        class-major concatenation, zero padding to C*K and num_boxes (nms.py:83-93).
// There is no separate sort pass and nothing of the size of a whole image is ever sorted: the filter buckets by class,
// each segment is sorted by the warp / CTA that runs its NMS.  Workspace: 32 KB per segment (0.08 of the logits at 90 classes
// and 107k anchors), independent of the number of anchors.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

#define FILTER_THREADS 256
#define FILTER_UNROLL 4
#define SEG_CAP 4096                 // keys per (image, class) region == keys a CTA sorts in shared memory (32 KB)
#define NMS_THREADS 512
#define NMS_CH 64                    // candidates per chunk
#define NMS_TPC (NMS_THREADS / NMS_CH)  // threads per candidate (a power of two <= 32)
#define ROUND_NB 64                  // histogram bins per round
#define ROUND_WANT 512               // a round collects at least this many keys (if that many are left) and at most SEG_CAP
#define ROUND_SMEM_MAX_C 320         // classes up to which the rounds keep their per-class tables / histograms in shared memory

struct KeyFormat {
    int abits;       // bits for the anchor index
    int cshift;      // 32 + abits
};

__host__ __device__ __forceinline__ unsigned order_desc_bits(unsigned u) {
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending total order of floats
    return ~u;                                        // descending
}
__device__ __forceinline__ unsigned order_desc(float s) { return order_desc_bits(__float_as_uint(s)); }
__device__ __forceinline__ float score_of_order(unsigned o) {
    unsigned u = ~o;
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    return __uint_as_float(u);
}
__device__ __forceinline__ float key_score(unsigned long long key, KeyFormat f) {
    return score_of_order((unsigned)((key >> f.abits) & 0xFFFFFFFFull));
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
// atomic; otherwise one atomic per candidate -- candidates are rare unless the scores are dense.  The counter keeps counting
// past the region's capacity (that is how an overflow is seen), but a CTA that has SEEN a segment past its capacity stops
// adding to it (one bit per segment in shared memory: `sat`, `sat_bits` of them, indexed seg - sat_seg0): nms_rounds_kernel
// re-derives such a segment from the scores, and with dense scores these atomics would otherwise all hit the same few
// counters (13.5 ms per 8 images at the stress configuration, measured in round 1; re-reading the counters instead of
// caching the verdict still took 2.6 ms: 70 M loads of the same twenty cache lines).
__device__ __forceinline__ void append_key(unsigned long long* __restrict__ cand, int* __restrict__ seg_count, long long seg,
                                           unsigned long long key, unsigned* sat, long long sat_seg0, int sat_bits) {
    const long long sb = seg - sat_seg0;
    const bool cached = sb >= 0 && sb < sat_bits;
    if (cached) {
        if ((sat[sb >> 5] >> (sb & 31)) & 1u) return;
    } else if (__ldcg(seg_count + seg) > SEG_CAP) return;
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    const long long seg_l = __shfl_sync(m, seg, leader);
    int pos, n = 1;
    if (__all_sync(m, seg == seg_l)) {
        int first = 0;
        n = __popc(m);
        if (lane == leader) first = atomicAdd(seg_count + seg, n);
        first = __shfl_sync(m, first, leader);
        pos = first + __popc(m & ((1u << lane) - 1u));
        if (first + n > SEG_CAP && cached && lane == leader) atomicOr(&sat[sb >> 5], 1u << (sb & 31));
    } else {
        pos = atomicAdd(seg_count + seg, 1);
        if (pos >= SEG_CAP && cached) atomicOr(&sat[sb >> 5], 1u << (sb & 31));
    }
    if (pos < SEG_CAP) cand[(size_t)seg * SEG_CAP + pos] = key;
}
#define FILTER_SAT_WORDS 2048          // 65536 segment bits (8 KB of shared memory) cached per CTA

// One level of a HeadGeom by value, and the index arithmetic that recovers (image, anchor, class) from a flat element index.
// Exact division of 32-bit numerators by a runtime constant: q = mulhi64(n, floor(2^64 / d) + 1) (round-up method: exact for
// every n < 2^32).  The candidate paths recover (anchor, class) from a flat element index with these; with dense scores
// every element is a candidate, and a 64-bit hardware-emulated division per element was half of the filter's instructions.
struct FastDiv {
    unsigned d;
    unsigned long long m;
};
__host__ __device__ __forceinline__ FastDiv fast_div(unsigned d) {
    FastDiv f;
    f.d = d ? d : 1u;
    f.m = f.d == 1u ? 0ull : (~0ull / f.d) + 1ull;
    return f;
}
__device__ __forceinline__ unsigned div_fast(unsigned n, const FastDiv f) {
    return f.d == 1u ? n : (unsigned)__umul64hi((unsigned long long)n, f.m);
}

struct LevelGeom {        // one level of a HeadGeom, by value
    int hw, per_loc, C, channels_first, anchor_off;
    FastDiv d_hw, d_nc, d_c, d_img;     // by hw, per_loc * C, C, per_loc * C * hw
};
__host__ __device__ __forceinline__ LevelGeom level_geom(const HeadGeom& G, int l) {
    LevelGeom g;
    g.hw = G.hw[l]; g.per_loc = G.per_loc; g.C = G.C; g.channels_first = G.channels_first; g.anchor_off = G.anchor_off[l];
    g.d_hw = fast_div((unsigned)g.hw); g.d_nc = fast_div((unsigned)(g.per_loc * g.C)); g.d_c = fast_div((unsigned)g.C);
    g.d_img = fast_div((unsigned)(g.per_loc * g.C * g.hw));
    return g;
}
// element r of ONE image's block of a level -> (anchor, class)
__device__ __forceinline__ void level_decompose(const LevelGeom g, int r, int& a, int& c) {
    const int hw = g.hw, n = g.per_loc, C = g.C;
    int loc, q;
    if (g.channels_first) { q = (int)div_fast((unsigned)r, g.d_hw); loc = r - q * hw; }
    else { loc = (int)div_fast((unsigned)r, g.d_nc); q = r - loc * (n * C); }
    const int k = (int)div_fast((unsigned)q, g.d_c);
    c = q - k * C;
    a = g.anchor_off + loc * n + k;
}
// element e of a level's whole tensor [B, ...] -> (image, anchor, class); the tensor holds < 2^31 elements per image and the
// image index is found with one 64-bit division only when the tensor itself is larger than 2^32 elements
__device__ __forceinline__ void head_decompose(const LevelGeom g, long long e, int& b, int& a, int& c) {
    const long long per_image = (long long)g.per_loc * g.C * g.hw;
    if (e < (1ll << 32)) b = (int)div_fast((unsigned)e, g.d_img);
    else b = (int)(e / per_image);
    level_decompose(g, (int)(e - (long long)b * per_image), a, c);
}

// Everything a candidate needs to find its segment, by value in shared memory (one copy per CTA).
#ifndef FILTER_MIN_CTAS
#define FILTER_MIN_CTAS 4              // at most 64 registers
#endif
#define FD_U 4
#define FD_NB 32                       // bins of filter_dense_kernel's per-segment value histograms
#define FD_RANGE 19.0f                 // logits: sigmoid(lim + 19) rounds to 1 for every usual threshold; scores: (thr, 1] is a subset
struct FilterEmit {
    unsigned long long* cand;
    int* seg_count;
    long long seg0;            // anchor-major: first segment of the CTA's image
    long long sat_seg0;
    int sat_bits;
    int C;
    int head;                  // 1: `g` describes the level scanned (element -> image, anchor, class through head_decompose)
    KeyFormat fmt;
    LevelGeom g;
    float thr, x_lo;
};

// Four consecutive elements starting at element e0, at least one of which may be a candidate.
template <bool IS_LOGITS, bool HEAD>
__device__ __forceinline__ void filter_group(const FilterEmit* __restrict__ E, unsigned* sat, const float4 v, long long e0) {
    const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float sc;
        if (!is_candidate<IS_LOGITS>(vals[j], E->thr, E->x_lo, &sc)) continue;
        const long long e = e0 + j;
        int a, c;
        long long seg;
        if (HEAD) {
            int b;
            head_decompose(E->g, e, b, a, c);
            seg = (long long)b * E->C + c;
        } else {
            a = (int)((unsigned)e / (unsigned)E->C);                          // per image: e < 2^31
            c = (int)e - a * E->C;
            seg = E->seg0 + c;
        }
        append_key(E->cand, E->seg_count, seg, make_key(c, sc, a, E->fmt), sat, E->sat_seg0, E->sat_bits);
    }
}

// One tile of the streaming scan: SCAN_U 128-bit no-allocate loads per thread (all in flight together), then every group of
// four elements whose maximum passes the (cheap) pre-test goes to filter_group.  FULL: the whole tile lies inside the array.
#ifndef SCAN_U
#define SCAN_U 6                       // anchor-major scan: 96 bytes in flight per thread, 64 registers, four CTAs per SM (profiles/r2b_filter_variants.txt)
#endif
#ifndef SCAN_U_HEAD
#define SCAN_U_HEAD 4                  // head-layout scan (its candidate path needs more registers: four-way index decomposition)
#endif
template <bool IS_LOGITS, bool HEAD, bool FULL, int U>
__device__ __forceinline__ void scan_tile(const float4* __restrict__ ptr, long long i0, long long nbody4, long long head, float lim,
                                          const FilterEmit* __restrict__ E, unsigned* sat, int* dense_flag, int* cta_stop, bool& stop) {
    float4 v[U];
    bool inb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        inb[u] = FULL || (i0 + u * FILTER_THREADS + threadIdx.x < nbody4);
        v[u] = inb[u] ? ld_stream_f4(ptr + u * FILTER_THREADS) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    // ONE test and branch per tile on the common path: the maximum over everything the thread loaded (out-of-bounds groups hold
    // -inf); only a thread that holds a possible candidate looks at its groups one by one
    float mx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) mx[u] = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
    float mall = mx[0];
#pragma unroll
    for (int u = 1; u < U; ++u) mall = fmaxf(mall, mx[u]);
    if (__builtin_expect(!(mall > lim), 1)) return;                       // (the hint moves the candidate code out of the loop body)
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (!(mx[u] > lim)) continue;
        if (!HEAD && dense_flag) {
            if (stop) continue;
            if (u == 0) {
                // a thread ALL of whose groups (4 KB apart) hold a possible candidate sits in a dense image: with 5 %
                // of the elements above the threshold that happens to 0.1 % of the threads, with 20 % to every tenth --
                // and with a few confident anchors per image never (their classes cannot line up in all the 4-class
                // windows).  Checked BEFORE anything is appended: in a dense image every thread would otherwise start
                // with a burst of contended atomics.  One shared-memory exchange per warp, one global one per CTA
                // (hundreds of thousands of stores to one word take 0.3 ms by themselves).
                bool all = true;
#pragma unroll
                for (int w = 1; w < U; ++w) all = all && mx[w] > lim;
                if (all) {
                    if ((int)(threadIdx.x & 31) == __ffs(__activemask()) - 1 && atomicExch(cta_stop, 1) == 0)
                        cta_stop[1] = atomicExch(dense_flag, 1) == 0 ? 1 : 0;        // the first CTA of the image to notice
                    stop = true;
                    continue;
                }
            }
            if (*(volatile int*)cta_stop || __ldcg(dense_flag)) { stop = true; continue; }
        }
        filter_group<IS_LOGITS, HEAD>(E, sat, v[u], head + ((i0 + u * FILTER_THREADS + threadIdx.x) << 2));
    }
}

// The streaming scan shared by both layouts: `count` floats at `base` are read once with 128-bit no-allocate loads;
// every group of four elements whose maximum passes the (cheap) pre-test goes to filter_group.  The candidate code is written
// fully unrolled and inlined on purpose: the compiler then lays the rare path out of the way of the streaming loop.  Measured
// on a B200 (cfg3, graph replay of the inference sub-path, profiles/round2_filter_variants.txt): this form 0.228 ms (scan
// 0.190 ms); the candidate code as an out-of-line function 0.248 ms (0.246 / 0.268 ms when the call forces spills at 48 / 40
// registers); values parked in shared memory and walked by a rolled loop 0.335 ms; a rolled 4-iteration loop 4.3 vs 6.2 TB/s
// (round 1).  The scan sits at the memory system's limit: same box, same run (profiles/r2b_filter_variants.txt), bounds tests
// on every load 0.1933 ms, full tiles without them and one branch per tile 0.1910 ms, six loads per thread 0.1881 ms, eight
// 0.190-0.192 ms, five CTAs per SM at 48 registers 0.198 ms -- 6.4-6.6 TB/s; box-to-box variation is larger (0.188-0.206 ms).
// `dense_flag` (anchor-major only): set by the first thread that finds possible candidates in all four of its groups; every
// thread that comes across a possible candidate afterwards stops its scan -- filter_dense_kernel redoes such an image from
// scratch.  Both checks sit on the candidate path, the streaming loop itself does not know about them (a flag load and a
// ballot per iteration cost 15 % of the scan's bandwidth).
template <bool IS_LOGITS, bool HEAD, int U>
__device__ __forceinline__ void scan_candidates(const float* __restrict__ base, long long count, const FilterEmit* __restrict__ E,
                                                unsigned* sat, int* dense_flag, int* cta_stop /*[2]: stop, won*/,
                                                long long first_tile /*of this CTA*/, long long tile_stride /*CTAs sharing the array*/) {
    const int lane = threadIdx.x & 31;
    const float lim = IS_LOGITS ? E->x_lo : E->thr;
    // peel to 16-byte alignment: head scalars | body float4 | tail scalars
    const unsigned mis = (unsigned)(((uintptr_t)base >> 2) & 3);
    long long head = mis ? (4 - mis) : 0;
    if (head > count) head = count;
    const long long nbody4 = (count - head) >> 2;
    const long long tail0 = head + (nbody4 << 2);
    const float4* body = (const float4*)(base + head);

    if (first_tile == 0 && threadIdx.x < 32) {
        // head + tail elements (< 8 in total), one lane each
        long long e = -1;
        if (lane < head) e = lane;
        else if (lane - head < count - tail0) e = tail0 + (lane - head);
        if (e >= 0 && base[e] > lim)
            filter_group<IS_LOGITS, HEAD>(E, sat, make_float4(base[e], -INFINITY, -INFINITY, -INFINITY), e);
    }

    // Full tiles (U x 256 float4 = U x 4 KB per CTA and iteration) are read without any bounds test; the one partial
    // tile of the array (the last one, met by a single CTA) goes through the bounded form of the same code.
    const long long tile = (long long)FILTER_THREADS * U;
    const long long stride = tile_stride * tile;
    bool stop = false;
    // the thread's load address advances by a constant (kept as a pointer: recomputing it from the loop index put eight
    // dependent integer instructions in front of every iteration's first load)
    const float4* ptr = body + first_tile * tile + threadIdx.x;
    long long i0 = first_tile * tile;
    for (; i0 + tile <= nbody4 && !stop; i0 += stride, ptr += stride)
        scan_tile<IS_LOGITS, HEAD, true, U>(ptr, i0, nbody4, head, lim, E, sat, dense_flag, cta_stop, stop);
    if (i0 < nbody4 && !stop) scan_tile<IS_LOGITS, HEAD, false, U>(ptr, i0, nbody4, head, lim, E, sat, dense_flag, cta_stop, stop);
}

// anchor-major layout: grid (gx, B), image blockIdx.y is the [A,C] array scanned
template <bool IS_LOGITS>
__global__ void __launch_bounds__(FILTER_THREADS, FILTER_MIN_CTAS) filter_kernel(
    const float* __restrict__ scores, long long per_image /*A*C*/, int C, float thr, float x_lo, KeyFormat fmt,
    unsigned long long* __restrict__ cand, int* __restrict__ seg_count /*[B*C]*/, int* __restrict__ img_dense /*[B] or NULL*/,
    int* __restrict__ seg_count_dense, unsigned* __restrict__ vhist) {
    extern __shared__ unsigned s_sat[];                                 // one bit per class of this image: segment seen past its capacity
    __shared__ FilterEmit s_emit;
    __shared__ int s_stop[2];
    pdl_launch_dependents();
    const int b = blockIdx.y;
    const int sat_bits = min(C, FILTER_SAT_WORDS * 32);
    for (int i = threadIdx.x; i < (sat_bits + 31) / 32; i += FILTER_THREADS) s_sat[i] = 0u;
    if (threadIdx.x == 0) {
        s_stop[0] = s_stop[1] = 0;
        FilterEmit e;
        e.cand = cand; e.seg_count = seg_count; e.seg0 = (long long)b * C; e.sat_seg0 = e.seg0; e.sat_bits = sat_bits;
        e.C = C; e.head = 0; e.fmt = fmt; e.thr = thr; e.x_lo = x_lo;
        memset(&e.g, 0, sizeof(e.g));
        s_emit = e;
    }
    __syncthreads();
    scan_candidates<IS_LOGITS, false, SCAN_U>(scores + (size_t)b * per_image, per_image, &s_emit, s_sat, img_dense ? img_dense + b : nullptr, s_stop,
                                               (long long)blockIdx.x, (long long)gridDim.x);
    if (img_dense) {
        // the CTA that declared the image dense prepares filter_dense_kernel's counters and value histograms for it (they are
        // not part of the per-call memset: zeroing 128 bytes per segment on every call cost 2-3 us of the sparse path)
        __syncthreads();
        if (s_stop[1]) {
            for (int i = threadIdx.x; i < C * FD_NB; i += FILTER_THREADS) vhist[(size_t)b * C * FD_NB + i] = 0u;
            for (int c = threadIdx.x; c < C; c += FILTER_THREADS) seg_count_dense[(size_t)b * C + c] = 0;
        }
    }
}

// Head layout (head-layout fusion, see head.cu): level l's class tensor [B, n*C, h, w] (or [B, h, w, n*C]) is scanned as ONE flat
// array by the CTAs [cta0[l], cta0[l + 1]) of a one-dimensional grid; only a candidate pays for the index arithmetic that
// recovers (image, anchor, class) from its flat position.  The levels differ fourfold in size from one to the next, so every
// level gets CTAs in proportion to its tiles (round 1 launched one full grid row per level: three quarters of the 11,840 CTAs
// found no tile or one; the scan then took 1.07-1.12 times as long as the anchor-major scan of the same bytes on the same box,
// now 1.04 times: profiles/r2q_bench.json vs r2e / r2p).
struct HeadCtas {
    int cta0[SSDK_MAX_LEVELS + 1];
};
template <bool IS_LOGITS>
__global__ void __launch_bounds__(FILTER_THREADS, FILTER_MIN_CTAS) head_filter_kernel(const HeadGeom G, int B, float thr, float x_lo, KeyFormat fmt,
                                                                    unsigned long long* __restrict__ cand,
                                                                    int* __restrict__ seg_count, const HeadCtas H) {
    extern __shared__ unsigned s_sat[];                                 // one bit per (image, class) segment (the first 65536 of them)
    __shared__ FilterEmit s_emit;
    pdl_launch_dependents();
    int l = 0;
    while (l + 1 < G.num_levels && (int)blockIdx.x >= H.cta0[l + 1]) ++l;
    const LevelGeom g = level_geom(G, l);
    const long long count = (long long)B * g.per_loc * g.C * g.hw;
    const int sat_bits = (int)min((long long)B * g.C, (long long)FILTER_SAT_WORDS * 32);
    for (int i = threadIdx.x; i < (sat_bits + 31) / 32; i += FILTER_THREADS) s_sat[i] = 0u;
    if (threadIdx.x == 0) {
        FilterEmit e;
        e.cand = cand; e.seg_count = seg_count; e.seg0 = 0; e.sat_seg0 = 0; e.sat_bits = sat_bits;
        e.C = g.C; e.head = 1; e.fmt = fmt; e.g = g; e.thr = thr; e.x_lo = x_lo;
        s_emit = e;
    }
    __syncthreads();
    scan_candidates<IS_LOGITS, true, SCAN_U_HEAD>(G.cls[l], count, &s_emit, s_sat, nullptr, nullptr, (long long)((int)blockIdx.x - H.cta0[l]),
                                                  (long long)(H.cta0[l + 1] - H.cta0[l]));
}

// ---------------------------------------------------------------------------------------------- 1b. dense images
// An image declared dense by filter_kernel is redone here from scratch (its own counters: seg_count_dense), by the whole
// grid row of that image; every other CTA exits at once.  A CTA works on tiles of 4096 consecutive elements:
//   phase 1  every element above the pre-test limit is counted in its class's VALUE histogram (shared memory; FD_NB bins over
//            [lim, lim + FD_RANGE) of the raw value -- logit or score --, no sigmoid); unless the CTA has seen the class's
//            segment past its capacity, the exact candidate test follows and a candidate takes a rank inside (tile, class);
//   phase 2  one global atomic per class and tile reserves the tile's slots (13.5 ms -> 0.5 ms against one atomic per
//            candidate in round 1: same-address atomics retire at ~8 M/s);
//   phase 3  the keys are written.
// At the end the value histograms of the saturated classes are added to the global ones: nms_rounds_kernel uses them to GUESS
// the value above which a region's worth of the best candidates lies, so that its first round needs no histogram pass of its
// own; the guess only has to be good, not exact (the collect pass counts).
__device__ __forceinline__ int value_bin(float v, float lim, float inv_w) {      // bin 0 = the best values
    const int b = FD_NB - 1 - (int)((v - lim) * inv_w);
    return b < 0 ? 0 : (b > FD_NB - 1 ? FD_NB - 1 : b);
}
template <bool IS_LOGITS>
__global__ void __launch_bounds__(FILTER_THREADS) filter_dense_kernel(
    const float* __restrict__ scores, long long per_image /*A*C*/, int C, float thr, float x_lo, KeyFormat fmt,
    unsigned long long* __restrict__ cand, int* __restrict__ seg_count_dense, const int* __restrict__ img_dense,
    unsigned* __restrict__ vhist /*[B*C][FD_NB]*/) {
    extern __shared__ int s_dense[];
    const int b = blockIdx.y, tid = threadIdx.x;
    pdl_wait();                                             // img_dense is written by filter_kernel
    pdl_launch_dependents();
    if (!img_dense[b]) return;
    int* s_cnt = s_dense;                                   // [C]
    int* s_base = s_dense + C;                              // [C]
    unsigned* s_satd = (unsigned*)(s_dense + 2 * C);        // [(C + 31) / 32]
    unsigned* s_vh = s_satd + (C + 31) / 32;                // [C][FD_NB]
    const float* base = scores + (size_t)b * per_image;
    const long long seg0 = (long long)b * C;
    const float lim = IS_LOGITS ? x_lo : thr;
    const float inv_w = (float)FD_NB / FD_RANGE;
    const FastDiv dc = fast_div((unsigned)C);
    for (int c = tid; c < C; c += FILTER_THREADS) s_cnt[c] = 0;
    for (int i = tid; i < (C + 31) / 32; i += FILTER_THREADS) s_satd[i] = 0u;
    for (int i = tid; i < C * FD_NB; i += FILTER_THREADS) s_vh[i] = 0u;
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
        if (e >= 0 && base[e] > lim) {
            const int a = (int)div_fast((unsigned)e, dc), c = (int)e - a * C;
            atomicAdd(&s_vh[c * FD_NB + value_bin(base[e], lim, inv_w)], 1u);
            if (is_candidate<IS_LOGITS>(base[e], thr, x_lo, &s)) {
                const int pos = atomicAdd(seg_count_dense + seg0 + c, 1);
                if (pos < SEG_CAP) cand[(size_t)(seg0 + c) * SEG_CAP + pos] = make_key(c, s, a, fmt);
            }
        }
    }
    const long long ntiles = (nbody4 + FILTER_THREADS * FD_U - 1) / (FILTER_THREADS * FD_U);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        float sc[4 * FD_U];
        unsigned short rk[4 * FD_U];
        unsigned hit = 0u;
        // phase 1
#pragma unroll
        for (int u = 0; u < FD_U; ++u) {
            const long long i4 = tile * (FILTER_THREADS * FD_U) + u * FILTER_THREADS + tid;
            const float4 v = i4 < nbody4 ? ld_stream_f4(body + i4) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            const float vals[4] = {v.x, v.y, v.z, v.w};
            const unsigned e0 = (unsigned)(head + (i4 << 2));
            unsigned a = div_fast(e0, dc);
            int c = (int)(e0 - a * (unsigned)C);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (vals[j] > lim) {
                    atomicAdd(&s_vh[c * FD_NB + value_bin(vals[j], lim, inv_w)], 1u);
                    if (!((s_satd[c >> 5] >> (c & 31)) & 1u) && is_candidate<IS_LOGITS>(vals[j], thr, x_lo, &sc[u * 4 + j])) {
                        rk[u * 4 + j] = (unsigned short)atomicAdd(&s_cnt[c], 1);
                        hit |= 1u << (u * 4 + j);
                    }
                }
                if (++c == C) c = 0;                                    // next element: next class (next anchor at the wrap)
            }
        }
        if (!__syncthreads_or((int)hit)) continue;                      // (every class saturated: the normal case after the first tile)
        // phase 2: one global atomic per class reserves the tile's slots
        for (int c = tid; c < C; c += FILTER_THREADS) {
            const int h = s_cnt[c];
            if (h) {
                const int first = atomicAdd(seg_count_dense + seg0 + c, h);
                s_base[c] = first;
                s_cnt[c] = 0;
                if (first + h > SEG_CAP) atomicOr(&s_satd[c >> 5], 1u << (c & 31));
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
                    const unsigned e = e0 + j, a = div_fast(e, dc), c = e - a * (unsigned)C;
                    const int pos = s_base[c] + rk[u * 4 + j];
                    if (pos < SEG_CAP) cand[(size_t)(seg0 + c) * SEG_CAP + pos] = make_key((int)c, sc[u * 4 + j], (int)a, fmt);
                }
            }
        }
        // no barrier here: the next tile's phase 2 (the only writer of s_base) comes after its phase-1 barrier, which every
        // thread reaches only after this phase 3
    }
    __syncthreads();
    for (int i = tid; i < C * FD_NB; i += FILTER_THREADS) {
        const int c = i / FD_NB;
        const unsigned h = s_vh[i];
        if (h && ((s_satd[c >> 5] >> (c & 31)) & 1u)) atomicAdd(&vhist[(size_t)(seg0 + c) * FD_NB + (i - c * FD_NB)], h);
    }
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

// CTA-wide bitonic sort of keys[0, n) in shared memory, ascending, for ANY n: the network is the all-ascending formulation
// (first step of a merge compares i with its mirror image, the others i with i + j); positions >= n behave as +infinity,
// never move, and their compare-exchanges are simply skipped.
__device__ __forceinline__ void sort_step(unsigned long long* keys, int n, int halfP, int tid, int nthreads, int lk, int j /*0: flip step*/) {
    const int k = 1 << lk, half = k >> 1;
    for (int t = tid; t < halfP; t += nthreads) {
        int ii, ll;
        if (j == 0) {
            const int blk = t >> (lk - 1), o = t & (half - 1);
            ii = (blk << lk) + o;
            ll = (blk << lk) + (k - 1 - o);
        } else {
            ii = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            ll = ii | j;
        }
        if (ll < n) {
            const unsigned long long x = keys[ii], y = keys[ll];
            if (x > y) { keys[ii] = y; keys[ll] = x; }
        }
    }
}
__device__ __forceinline__ void cta_sort_keys(unsigned long long* keys, int n, int tid, int nthreads) {
    int P = 1, logP = 0;
    while (P < n) { P <<= 1; ++logP; }
    for (int lk = 1; lk <= logP; ++lk) {
        sort_step(keys, n, P >> 1, tid, nthreads, lk, 0);
        __syncthreads();
        for (int j = 1 << (lk - 2 >= 0 ? lk - 2 : 0); lk >= 2 && j > 0; j >>= 1) {
            sort_step(keys, n, P >> 1, tid, nthreads, lk, j);
            __syncthreads();
        }
    }
}

// Rank sort for n <= nthreads keys (all distinct: the anchor index is part of the key): a key's rank is the number of smaller
// keys; `parts` = nthreads / P threads share one key's count (P = n rounded up to a power of two) and add up with shuffles.
// Two barriers instead of the bitonic network's log^2: a third of nms_kernel's stall samples sat in that network
// (ncu source view, profiles/r2m_ncu_summary.txt era build) for segments of 60-120 keys.  dst must not overlap src.
__device__ __forceinline__ void cta_rank_sort_keys(const unsigned long long* __restrict__ src, unsigned long long* __restrict__ dst, int n,
                                                   int tid, int nthreads) {
    int P = 32;
    while (P < n) P <<= 1;
    const int parts = nthreads / P;                                   // a power of two in [1, 16] (n > 32, nthreads <= 512)
    const int i = tid / parts, part = tid - i * parts;
    const int span = P / parts;
    int rank = 0;
    if (i < n) {
        const unsigned long long mine = src[i];
        const int j1 = min(n, (part + 1) * span);
        for (int j = part * span; j < j1; ++j) rank += src[j] < mine ? 1 : 0;
    }
    for (int o = 1; o < parts; o <<= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    if (i < n && part == 0) dst[rank] = src[i];
    __syncthreads();
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
__device__ __forceinline__ NmsBox nms_box_of(const float4 raw, float& area) {
    NmsBox box;
    box.ymin = fminf(raw.x, raw.z); box.xmin = fminf(raw.y, raw.w);
    box.ymax = fmaxf(raw.x, raw.z); box.xmax = fmaxf(raw.y, raw.w);
    area = f_mul(f_sub(box.ymax, box.ymin), f_sub(box.xmax, box.xmin));
    return box;
}

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

// header ints of the counter workspace (zeroed every call)
enum { H_HEAVY = 0, H_PEND, H_BAR, H_ERR, H_NIMG_C, H_NHIST0, H_NHIST1, H_NIMG_H0, H_NIMG_H1, H_WORDS = 16 };

// Candidate count of a segment: from filter_kernel's counters, or from filter_dense_kernel's when the image was redone by it.
struct SegCounts {
    const int* sparse;
    const int* dense;          // NULL when the dense-image path is not in use
    const int* img_dense;
};
__device__ __forceinline__ int seg_candidates(const SegCounts SC, long long seg, int C) {
    if (SC.dense && SC.img_dense[seg / C]) return SC.dense[seg];
    return SC.sparse[seg];
}

// Segments with at most 32 candidates (the vast majority: background classes) are resolved by ONE WARP each, one
// candidate per lane: decode, a 32x32 suppression bit matrix (one column per lane), a greedy walk with shuffles.
// Larger segments are pushed to a queue: `heavy` (fits its region: nms_kernel spends a whole CTA on each) or `pending`
// (overflowed its region: nms_rounds_kernel).
#define NMS_SMALL_WARPS 8
template <bool DECODED>
__global__ void __launch_bounds__(NMS_SMALL_WARPS * 32) nms_small_kernel(
    const unsigned long long* __restrict__ cand, KeyFormat fmt, const SegCounts SC,
    const CodeView codes, const float4* __restrict__ anchors, long long A,
    long long nseg, int C, int K, float iou_thr, float4* __restrict__ seg_box, float* __restrict__ seg_score,
    int* __restrict__ seg_anchor, int* __restrict__ seg_kept, int* __restrict__ heavy_queue, int* __restrict__ pend_queue,
    int* __restrict__ hdr) {
    __shared__ NmsBox s_tile[NMS_SMALL_WARPS][32];
    __shared__ float s_tile_area[NMS_SMALL_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long seg = (long long)blockIdx.x * NMS_SMALL_WARPS + warp;
    pdl_wait();                                             // candidates and counters come from the filter kernels
    pdl_launch_dependents();
    if (seg >= nseg) return;
    const int n = seg_candidates(SC, seg, C);
    if (n <= 0) {
        if (lane == 0) seg_kept[seg] = 0;
        return;
    }
    if (n > 32) {
        if (lane == 0) {
            if (n > SEG_CAP) pend_queue[atomicAdd(hdr + H_PEND, 1)] = (int)seg;
            else heavy_queue[atomicAdd(hdr + H_HEAVY, 1)] = (int)seg;
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
    const unsigned long long key = warp_sort_keys(alive ? cand[(size_t)seg * SEG_CAP + lane] : ~0ull, lane);
    if (alive) {
        a = key_anchor(key, fmt);
        score = key_score(key, fmt);
        raw = load_code(codes, b, A, a);
        if (!DECODED) raw = box_clip01(box_decode(raw, anchors[a]));                         // nms.py:76-77
        box = nms_box_of(raw, area);
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

// Greedy NMS of ONE segment over `navail` keys sorted ascending (score descending), by a whole CTA, NMS_CH candidates per
// chunk in sorted order, NMS_TPC threads per candidate; CONTINUES a list of `kept` boxes already in sh.kept (0 for a fresh
// segment):
//   (a) every candidate is decoded + clipped and tested against the boxes kept so far (shared memory); the threads of a
//       candidate split the kept list and OR their verdicts with shuffles;
//   (b) column c of the chunk's suppression bit matrix (bit t set <=> t < c, t survived (a), IoU(t, c) > threshold) is
//       built the same way, the threads splitting t;
//   (c) warp 0 resolves the greedy order WITHOUT walking it: a candidate is removed as soon as one of its suppressors is
//       known to be kept, and kept as soon as all of them are known to be removed; every round decides at least the first
//       undecided candidate, in practice a handful of rounds of ballots decide all 64;  the result is cut at K;
//   (d) kept candidates append themselves (rank by popcount) to the kept list and to the segment's output.
// Returns the new number of kept boxes (CTA-uniform).
struct NmsShared {
    NmsBox tile[NMS_CH];
    float tile_area[NMS_CH];
    unsigned long long col[NMS_CH];
    unsigned alive32[2][2];                // [chunk parity][word]: survivors of (a), set with atomicOr
    unsigned long long keep;
};
struct NmsSegArgs {
    KeyFormat fmt;
    CodeView codes;
    const float4* anchors;
    long long A;
    int C, K;
    float iou_thr;
    float4* seg_box; float* seg_score; int* seg_anchor;
};

template <bool DECODED, int THREADS = NMS_THREADS>
__device__ __forceinline__ int nms_sorted_segment(const NmsSegArgs& N, NmsShared& sh, NmsBox* s_kept, float* s_kept_area,
                                                  const unsigned long long* sorted, int navail, long long seg, int kept) {
    constexpr int TPC = THREADS / NMS_CH;              // threads per candidate (a power of two <= 32)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = tid / TPC, q = tid % TPC;            // candidate slot in the chunk, position among its threads
    const int b = (int)(seg / N.C), K = N.K;
    const size_t obase = (size_t)seg * K;
    const float iou_thr = N.iou_thr, band = fabsf(iou_thr) * 3.814697265625e-06f;
    if (tid < 4) sh.alive32[tid >> 1][tid & 1] = 0u;
    __syncthreads();
    // the candidate of the NEXT chunk (key -> code -> anchor: dependent global loads) is fetched while the current
    // chunk is processed
    unsigned long long nkey = 0ull;
    float4 ncode = make_float4(0.f, 0.f, 0.f, 0.f), nanc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch = [&](int i) {
        if (i < navail) {
            nkey = sorted[i];
            const int na = key_anchor(nkey, N.fmt);
            ncode = load_code(N.codes, b, N.A, na);
            if (!DECODED) nanc = N.anchors[na];
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
            a = key_anchor(nkey, N.fmt);
            score = key_score(nkey, N.fmt);
            if (DECODED) raw = ncode;
            else raw = box_clip01(box_decode(ncode, nanc));                              // nms.py:76-77
            box = nms_box_of(raw, area);
        }
        fetch(i + NMS_CH);
        // (a) against the boxes kept so far
        int hit = 0;
        if (alive) {
            for (int j = q; j < kept; j += TPC) {
                bool y, am;
                nms_fast(box, area, s_kept[j], s_kept_area[j], iou_thr, band, y, am);
                if (am) y = nms_exact(box, area, s_kept[j], s_kept_area[j], iou_thr); // rare: within 2^-18 of the threshold
                hit |= y ? 1 : 0;
            }
        }
#pragma unroll
        for (int o = 1; o < TPC; o <<= 1) hit |= __shfl_xor_sync(0xffffffffu, hit, o);
        alive = alive && !hit;
        if (q == 0) {
            sh.tile[c] = box;
            sh.tile_area[c] = area;
            if (alive) atomicOr(&sh.alive32[parity][c >> 5], 1u << (c & 31));
        }
        if (tid < 2) sh.alive32[parity ^ 1][tid] = 0u;      // the buffer of the next chunk (last read two barriers ago)
        __syncthreads();
        const unsigned long long alive_mask = ((unsigned long long)sh.alive32[parity][1] << 32) | sh.alive32[parity][0];
        // (b) column of the chunk's suppression matrix
        unsigned long long col = 0ull;
        if (alive && (alive_mask & (alive_mask - 1ull))) {               // at least two survivors in the chunk
            for (int t = q; t < c; t += TPC) {
                if ((alive_mask >> t) & 1ull) {
                    bool y, am;
                    nms_fast(box, area, sh.tile[t], sh.tile_area[t], iou_thr, band, y, am);
                    if (am) y = nms_exact(box, area, sh.tile[t], sh.tile_area[t], iou_thr);
                    if (y) col |= 1ull << t;
                }
            }
        }
#pragma unroll
        for (int o = 1; o < TPC; o <<= 1) col |= __shfl_xor_sync(0xffffffffu, col, o);
        if (q == 0) sh.col[c] = col;
        __syncthreads();
        // (c) greedy order by resolution rounds (warp 0: lane l owns candidates l and l + 32)
        if (warp == 0) {
            const unsigned long long col_lo = sh.col[lane], col_hi = sh.col[lane + 32];
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
            if (lane == 0) sh.keep = keep;
        }
        __syncthreads();
        const unsigned long long keep = sh.keep;
        // (d) append the kept candidates in score order
        if (q == 0 && ((keep >> c) & 1ull)) {
            const int pos = kept + __popcll(keep & ((1ull << c) - 1ull));
            s_kept[pos] = box;
            s_kept_area[pos] = area;
            N.seg_box[obase + pos] = raw;
            N.seg_score[obase + pos] = score;
            N.seg_anchor[obase + pos] = a;
        }
        kept += __popcll(keep);
        __syncthreads();
    }
    return kept;
}

// One CTA per queued (image, class) segment with 33..SEG_CAP candidates: sort in shared memory, then nms_sorted_segment.
// NMS_HEAVY_THREADS: 512 threads (eight per candidate).  256 threads at 45 registers (a CTA that fits next to five resident CTAs
// of the fused training-step kernel) was measured for the two-stream step: 0.2353 vs 0.2295 ms for the inference sub-path alone and no
// gain for the step (0.355 vs 0.351 ms, profiles/r2k_nms256.txt) -- with both sub-paths streaming, the step is bound by the aggregate
// HBM bandwidth (2.05 GB in 0.351 ms = 5.85 TB/s), whatever the split of the SMs.
#ifndef NMS_HEAVY_THREADS
#define NMS_HEAVY_THREADS 512
#endif
template <bool DECODED, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 5 : 2) nms_kernel(const unsigned long long* __restrict__ cand, const SegCounts SC,
                                                          const NmsSegArgs N, int* __restrict__ seg_kept,
                                                          const int* __restrict__ heavy_queue, const int* __restrict__ hdr) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    __shared__ NmsShared sh;
    const int tid = threadIdx.x;
    unsigned long long* s_sort = (unsigned long long*)nms_smem;     // [SEG_CAP]
    NmsBox* s_kept = (NmsBox*)(s_sort + SEG_CAP);                   // [K]
    float* s_kept_area = (float*)(s_kept + N.K);                    // [K]
    pdl_wait();                                                     // the queue comes from nms_small_kernel
    pdl_launch_dependents();
    const int nheavy = hdr[H_HEAVY];
    for (int item = blockIdx.x; item < nheavy; item += gridDim.x) {
        const long long seg = heavy_queue[item];
        const int n = min(seg_candidates(SC, seg, N.C), SEG_CAP);
        const unsigned long long* keys = cand + (size_t)seg * SEG_CAP;
        __syncthreads();
        for (int i = tid; i < n; i += THREADS) s_sort[i] = keys[i];
        __syncthreads();
        const unsigned long long* sorted = s_sort;              // score descending, anchor ascending
        if (n <= THREADS) {
            cta_rank_sort_keys(s_sort, s_sort + SEG_CAP / 2, n, tid, THREADS);
            sorted = s_sort + SEG_CAP / 2;
        } else {
            cta_sort_keys(s_sort, n, tid, THREADS);
        }
        const int kept = nms_sorted_segment<DECODED, THREADS>(N, sh, s_kept, s_kept_area, sorted, n, seg, 0);
        if (tid == 0) seg_kept[seg] = kept;
    }
}

// ---------------------------------------------------------------------------------------------- 4. rounds (overflowing segments)
// State of a pending segment, in k-space: k = key without its class bits = order(score) << abits | anchor; every candidate has
// k < k_thr; smaller k = better.  lo: every key < lo has been through the segment's NMS.  [dlo, dhi): the key range the
// current histogram resolves (bin = (k - dlo) >> sh; keys in [lo, dlo) count into bin 0, keys >= dhi are not looked at, their
// number is `beyond`).  mid: the round collects the keys in [lo, mid).
enum { ST_HIST = 1, ST_COLLECT = 2, ST_DONE = 3 };
struct RoundState {
    int* hdr;                      // H_* words
    int* pend_queue;               // [nseg]
    int* st; int* sh; int* fill; int* rem; int* beyond;              // [nseg]
    unsigned long long* lo; unsigned long long* mid; unsigned long long* dlo; unsigned long long* dhi;   // [nseg]
    unsigned* hist;                // [nseg][ROUND_NB]
    int* stamp_h; int* stamp_c;    // [B]: image listed for the histogram / collect pass of round number = stamp
    int* list_h;                   // [2][B]
    int* list_c;                   // [B]
    unsigned long long k_thr, dlo0;
    float* vcut;                   // [nseg]: collect pass: no key < mid has a raw value (logit / score) below this
    const unsigned* vhist;         // [nseg][FD_NB] value histograms of filter_dense_kernel, or NULL
    const int* img_dense;          // [B]
    float vlim, vwidth;            // value-histogram geometry: bin FD_NB - 1 - q holds [vlim + q * vwidth, vlim + (q + 1) * vwidth)
    int* err;                      // the context's sticky asynchronous-error word
    unsigned long long* times;     // [ROUND_TIMES]: [0] = rounds run, then %globaltimer (ns) at the start and after every phase of the first rounds
};
#define ROUND_TIMES 32
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Lower bound of the raw value of any element whose key is < mid: such keys have an order word <= mid's, i.e. a score >=
// score_of_order(mid >> abits); for logits the bound goes back through the sigmoid with a margin.
template <bool IS_LOGITS>
__device__ __forceinline__ float value_bound_of_key(unsigned long long mid, int abits) {
    const unsigned long long om = mid >> abits;
    if (om > 0xFFFFFFFFull) return -INFINITY;
    const float s_mid = score_of_order((unsigned)om);
    if (!IS_LOGITS) return s_mid;
    if (s_mid >= 1.0f) return 15.0f;                                     // sigmoid rounds to 1 only above 16.6
    if (!(s_mid > 0.0f)) return -INFINITY;
    const float lg = logf(s_mid / (1.0f - s_mid));
    return lg - 1e-3f * (1.0f + fabsf(lg));
}

__device__ __forceinline__ int ceil_log2_per_bin(unsigned long long width) {   // smallest s with ROUND_NB << s >= width
    const unsigned long long per = (width + ROUND_NB - 1) / ROUND_NB;
    return per <= 1ull ? 0 : 64 - __clzll((long long)(per - 1ull));
}

// all CTAs of the (co-resident) grid; `epoch` counts the barriers passed.  Gives up after ~10 s (a CTA that was never scheduled).
__device__ __forceinline__ bool grid_barrier(int* hdr, int& epoch) {
    __shared__ int s_ok;
    ++epoch;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(hdr + H_BAR, 1);
        const int target = epoch * (int)gridDim.x;
        const long long t0 = clock64();
        int ok = 1;
        while (*((volatile int*)(hdr + H_BAR)) < target) {
            if (*((volatile int*)(hdr + H_ERR)) || clock64() - t0 > 20000000000ll) { atomicExch(hdr + H_ERR, 1); ok = 0; break; }
            __nanosleep(64);
        }
        __threadfence();
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

// CTA-cooperative streaming scan of `count` floats at `base` (peeled to 16-byte alignment); emit(r, v) for every element with
// v > lim (r = element index inside the block).  Same structure as scan_candidates: values that may matter are parked in the
// thread's shared-memory slot and walked in a rolled loop, so that `emit` is instantiated twice, not seventeen times.
template <typename Emit>
__device__ __forceinline__ void cta_scan(const float* __restrict__ base, int count, float lim, float (*slot)[NMS_THREADS], Emit emit) {
    const int tid = threadIdx.x;
    const unsigned mis = (unsigned)(((uintptr_t)base >> 2) & 3);
    int head = mis ? (int)(4 - mis) : 0;
    if (head > count) head = count;
    const int nbody4 = (count - head) >> 2;
    const int tail0 = head + (nbody4 << 2);
    const float4* body = (const float4*)(base + head);
    if (tid < 32) {
        int r = -1;
        if (tid < head) r = tid;
        else if (tid - head < count - tail0) r = tail0 + (tid - head);
        if (r >= 0 && base[r] > lim) emit(r, base[r]);
    }
    for (int i0 = 0; i0 < nbody4; i0 += NMS_THREADS * FILTER_UNROLL) {
        float4 v[FILTER_UNROLL];
#pragma unroll
        for (int u = 0; u < FILTER_UNROLL; ++u) {
            const int i = i0 + u * NMS_THREADS + tid;
            v[u] = i < nbody4 ? ld_stream_f4(body + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < FILTER_UNROLL; ++u) mx = fmaxf(mx, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
        if (!(mx > lim)) continue;
#pragma unroll
        for (int u = 0; u < FILTER_UNROLL; ++u) {
            slot[4 * u + 0][tid] = v[u].x; slot[4 * u + 1][tid] = v[u].y; slot[4 * u + 2][tid] = v[u].z; slot[4 * u + 3][tid] = v[u].w;
        }
#pragma unroll 1
        for (int k = 0; k < 4 * FILTER_UNROLL; ++k) {
            const float x = slot[k][tid];
            if (x > lim) emit(head + ((i0 + (k >> 2) * NMS_THREADS + tid) << 2) + (k & 3), x);
        }
    }
}

// The same scan for a SPARSE survivor set (the collect pass: ~1 % of a dense image lies above the class-independent bound):
// handling the survivors where they are found means a ~100-instruction path (index arithmetic, sigmoid, 64-bit key compares,
// the append) executed by one or two lanes of a warp, several times per iteration -- 0.54 ms per pass over the stress
// configuration.  Here every warp compacts its survivors (value, element index) into a small shared-memory queue and handles
// 32 of them at a time with all lanes busy.
#define SCAN_QUEUE 128                 // entries per warp; an iteration adds at most 512, normally < 16
struct ScanQueueEntry { float v; int r; };
template <typename Emit>
__device__ __forceinline__ void cta_scan_compact(const float* __restrict__ base, int count, float lim, ScanQueueEntry* queue /*[SCAN_QUEUE] of this warp*/,
                                                 Emit emit) {
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned mis = (unsigned)(((uintptr_t)base >> 2) & 3);
    int head = mis ? (int)(4 - mis) : 0;
    if (head > count) head = count;
    const int nbody4 = (count - head) >> 2;
    const int tail0 = head + (nbody4 << 2);
    const float4* body = (const float4*)(base + head);
    if (tid < 32) {
        int r = -1;
        if (tid < head) r = tid;
        else if (tid - head < count - tail0) r = tail0 + (tid - head);
        if (r >= 0 && base[r] > lim) emit(r, base[r]);
    }
    int qn = 0;                                                            // warp-uniform
    for (int i0 = 0; i0 < nbody4; i0 += NMS_THREADS * FILTER_UNROLL) {      // trip count uniform over the CTA
        float4 v[FILTER_UNROLL];
#pragma unroll
        for (int u = 0; u < FILTER_UNROLL; ++u) {
            const int i = i0 + u * NMS_THREADS + tid;
            v[u] = i < nbody4 ? ld_stream_f4(body + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < FILTER_UNROLL; ++u) mx = fmaxf(mx, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
        if (!__any_sync(0xffffffffu, mx > lim)) continue;
        const float vals[4 * FILTER_UNROLL] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w,
                                                v[2].x, v[2].y, v[2].z, v[2].w, v[3].x, v[3].y, v[3].z, v[3].w};
        static_assert(FILTER_UNROLL == 4, "vals[] spells out four float4");
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < 4 * FILTER_UNROLL; ++k) cnt += vals[k] > lim ? 1 : 0;
        int incl = cnt;                                                    // inclusive prefix sum over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (qn + total > SCAN_QUEUE) {
            // (never in practice) more survivors than the queue takes: handle this iteration's where they are
#pragma unroll
            for (int k = 0; k < 4 * FILTER_UNROLL; ++k)
                if (vals[k] > lim) emit(head + ((i0 + (k >> 2) * NMS_THREADS + tid) << 2) + (k & 3), vals[k]);
            continue;
        }
        int pos = qn + incl - cnt;
#pragma unroll
        for (int k = 0; k < 4 * FILTER_UNROLL; ++k) {
            if (vals[k] > lim) {
                queue[pos].v = vals[k];
                queue[pos].r = head + ((i0 + (k >> 2) * NMS_THREADS + tid) << 2) + (k & 3);
                ++pos;
            }
        }
        qn += total;
        __syncwarp();
        while (qn >= 32) {                                                 // 32 survivors, one per lane
            const ScanQueueEntry e = queue[qn - 32 + lane];
            qn -= 32;
            __syncwarp();
            emit(e.r, e.v);
        }
    }
    __syncwarp();
    if (lane < qn) {
        const ScanQueueEntry e = queue[lane];
        emit(e.r, e.v);
    }
    __syncwarp();
}

// per-class tables of one image during a streaming pass (shared memory when C <= ROUND_SMEM_MAX_C, else read from global)
struct ClassTables {
    unsigned char* st;             // [C]
    unsigned char* shift;          // [C]
    float* vmin;                   // [C]   collect pass: conservative lower bound of the value (logit or score) of a key < mid
    float* vmin_all;               // [1]   collect pass: the smallest vmin over the image's collecting classes
    unsigned long long* lo;        // [C]
    unsigned long long* a;         // [C]   histogram pass: dlo;  collect pass: mid
    unsigned long long* b;         // [C]   histogram pass: dhi
    unsigned* hist;                // [C][ROUND_NB]
};

template <bool DECODED, bool IS_LOGITS>
__global__ void __launch_bounds__(NMS_THREADS, 2) nms_rounds_kernel(const HeadGeom G, int B, float thr, float x_lo, unsigned long long* __restrict__ cand,
                                                                 const NmsSegArgs N, int* __restrict__ seg_kept, const RoundState R) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    __shared__ NmsShared sh;
    __shared__ union {                                               // never in use at the same time
        float slot[4 * FILTER_UNROLL][NMS_THREADS];                  // histogram pass (cta_scan)
        ScanQueueEntry queue[NMS_THREADS / 32][SCAN_QUEUE];          // collect pass (cta_scan_compact)
    } s_scan;
    pdl_wait();                                                      // queues, counters and kept boxes come from the kernels before
    pdl_launch_dependents();
    const int npend = R.hdr[H_PEND];
    if (npend == 0) {                                                // the normal case: no segment overflowed its region
        if (blockIdx.x == 0 && threadIdx.x == 0) R.times[0] = 0ull;
        return;
    }
    const int tid = threadIdx.x, lane = tid & 31;
    const int C = N.C, K = N.K;
    const KeyFormat fmt = N.fmt;
    const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
    unsigned long long* s_sort = (unsigned long long*)nms_smem;     // [SEG_CAP]
    NmsBox* s_kept = (NmsBox*)(s_sort + SEG_CAP);                   // [K]
    float* s_kept_area = (float*)(s_kept + K);                      // [K]
    const bool tables_in_smem = C <= ROUND_SMEM_MAX_C;
    ClassTables T;
    {
        unsigned char* p = (unsigned char*)(s_kept_area + K);
        p = (unsigned char*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
        T.lo = (unsigned long long*)p; p += (size_t)C * 8;
        T.a = (unsigned long long*)p; p += (size_t)C * 8;
        T.b = (unsigned long long*)p; p += (size_t)C * 8;
        T.hist = (unsigned*)p; p += (size_t)C * ROUND_NB * 4;
        T.vmin = (float*)p; p += (size_t)C * 4;
        T.vmin_all = (float*)p; p += 16;
        T.st = p; p += C;
        T.shift = p;
    }
    const unsigned long long kmask = (1ull << fmt.cshift) - 1ull;
    const bool flat_geom = G.num_levels == 1 && !G.channels_first && G.per_loc == 1;   // the anchor-major tensor: class = element % C
    int epoch = 0, stamp = 1;
    auto mark = [&]() { if (cta == 0 && tid == 0 && stamp < ROUND_TIMES) R.times[stamp++] = global_timer_ns(); };
    mark();

    // ---- init.  A segment of an image that filter_dense_kernel redid comes with a VALUE histogram: the value above which
    //      about ROUND_WANT of its best candidates lie is read off it, turned into a key bound `mid` (for logits through the
    //      sigmoid, 8 ulps up: expf is good to 2 ulps and the division is exact, so no element with a smaller logit can have a
    //      score above that) and the first round collects [0, mid) right away; the collect pass counts exactly, and a guess that
    //      turns out too generous falls back to the exact key-space histogram.  Every other segment starts with that histogram
    //      over [dlo0, k_thr).
    static_assert(FD_NB == 32 && ROUND_NB == 64, "one value-histogram bin and two key-histogram bins per lane");
    for (int p = cta * (NMS_THREADS / 32) + (tid >> 5); p < npend; p += grid * (NMS_THREADS / 32)) {     // one warp per segment
        const int seg = R.pend_queue[p];
        const int b = seg / C;
        R.hist[(size_t)seg * ROUND_NB + lane] = 0u;
        R.hist[(size_t)seg * ROUND_NB + 32 + lane] = 0u;
        bool guessed = false;
        float v_cut = 0.0f, s_cut = 0.0f;
        if (R.vhist && R.img_dense[b]) {
            long long incl = R.vhist[(size_t)seg * FD_NB + lane];            // inclusive prefix sums: bin 0 = the best values
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const unsigned reach = __ballot_sync(0xffffffffu, incl >= ROUND_WANT + ROUND_WANT / 4);
            const int jb = reach ? __ffs(reach) - 1 : FD_NB;
            const long long cum = jb < FD_NB ? __shfl_sync(0xffffffffu, incl, jb) : 0;
            if (jb < FD_NB - 1 && cum <= SEG_CAP - SEG_CAP / 4) {
                v_cut = R.vlim + (float)(FD_NB - 1 - jb) * R.vwidth;
                s_cut = v_cut;
                if (IS_LOGITS) s_cut = __uint_as_float(__float_as_uint(f_div(1.0f, f_add(1.0f, expf(-v_cut)))) + 8u);
                guessed = s_cut < 1.0f && s_cut > thr;
            }
        }
        if (lane != 0) continue;
        R.lo[seg] = 0ull; R.beyond[seg] = 0; R.rem[seg] = 0; R.fill[seg] = 0;
        seg_kept[seg] = 0;
        if (guessed) {
            R.mid[seg] = (unsigned long long)order_desc(s_cut) << fmt.abits;             // keys < mid <=> score > s_cut
            R.vcut[seg] = v_cut;
            R.rem[seg] = 0x3fffffff;                                                     // unknown: assume there is more
            R.st[seg] = ST_COLLECT;
            if (atomicExch(R.stamp_c + b, 1) != 1) R.list_c[atomicAdd(R.hdr + H_NIMG_C, 1)] = b;
        } else {
            R.st[seg] = ST_HIST; R.dlo[seg] = R.dlo0; R.dhi[seg] = R.k_thr;
            R.sh[seg] = ceil_log2_per_bin(R.k_thr - R.dlo0);
            atomicAdd(R.hdr + H_NHIST0 + 1, 1);
            if (atomicExch(R.stamp_h + b, 1) != 1) R.list_h[(size_t)B + atomicAdd(R.hdr + H_NIMG_H0 + 1, 1)] = b;   // round 1: list_h[1]
        }
    }
    if (!grid_barrier(R.hdr, epoch)) { if (tid == 0) *R.err = SSDK_ASYNC_ROUNDS_TIMEOUT; return; }
    mark();

    // streaming decomposition of one pass: unit u -> (image list[u / group], slice u % group of every level's block)
    auto for_units = [&](const int* list, int n_img, auto&& per_unit) {
        const int group = n_img >= grid ? 1 : grid / n_img;
        const long long units = (long long)n_img * group;
        for (long long u = cta; u < units; u += grid) per_unit(list[u / group], (int)(u % group), group);
    };
    auto load_tables = [&](int b, int want_state) {
        // per-class view of image b's segments for this pass; returns through T (shared memory) when it fits
        for (int c = tid; c < C; c += NMS_THREADS) {
            const int seg = b * C + c;
            const int st = R.st[seg];
            T.st[c] = (unsigned char)(st == want_state ? 1 : 0);
            if (st != want_state) continue;
            T.lo[c] = R.lo[seg];
            if (want_state == ST_HIST) {
                T.a[c] = R.dlo[seg]; T.b[c] = R.dhi[seg]; T.shift[c] = (unsigned char)R.sh[seg];
            } else {
                T.a[c] = R.mid[seg];
                T.vmin[c] = R.vcut[seg];
            }
        }
        if (want_state == ST_HIST)
            for (int i = tid; i < C * ROUND_NB; i += NMS_THREADS) T.hist[i] = 0u;
        __syncthreads();
        if (want_state == ST_COLLECT) {
            if (tid == 0) {
                float m = INFINITY;
                for (int c = 0; c < C; ++c)
                    if (T.st[c]) m = fminf(m, T.vmin[c]);
                T.vmin_all[0] = m;
            }
            __syncthreads();
        }
    };

    for (int round = 1;; ++round) {
        const int par = round & 1;
        const int n_img_h = R.hdr[H_NIMG_H0 + par];
        const int n_hist = R.hdr[H_NHIST0 + par];
        if (n_hist == 0 && (round > 1 || R.hdr[H_NIMG_C] == 0)) break;   // uniform: written before the last barrier
        if (cta == 0 && tid == 0) R.times[0] = (unsigned long long)round;
        mark();
        if (cta == 0 && tid == 0) {                                  // counters the NEXT round reads
            R.hdr[H_NIMG_H0 + (par ^ 1)] = 0;
            R.hdr[H_NHIST0 + (par ^ 1)] = 0;
        }
        // ---- (1) histogram pass
        for_units(R.list_h + (size_t)par * B, n_img_h, [&](int b, int slice, int group) {
            if (tables_in_smem) load_tables(b, ST_HIST);
            for (int l = 0; l < G.num_levels; ++l) {
                const LevelGeom g = level_geom(G, l);
                const int blk = g.per_loc * g.C * g.hw;
                const int r0 = (int)((long long)blk * slice / group), r1 = (int)((long long)blk * (slice + 1) / group);
                const float* base = G.cls[l] + (size_t)b * blk;
                cta_scan(base + r0, r1 - r0, IS_LOGITS ? x_lo : thr, s_scan.slot, [&](int r, float v) {
                    int a, c;
                    if (flat_geom) { a = (int)div_fast((unsigned)(r0 + r), g.d_c); c = r0 + r - a * C; }
                    else level_decompose(g, r0 + r, a, c);
                    const int seg = b * C + c;
                    if (tables_in_smem ? !T.st[c] : (R.st[seg] != ST_HIST)) return;
                    float s;
                    if (!is_candidate<IS_LOGITS>(v, thr, x_lo, &s)) return;
                    const unsigned long long k = ((unsigned long long)order_desc(s) << fmt.abits) | (unsigned long long)a;
                    const unsigned long long lo = tables_in_smem ? T.lo[c] : R.lo[seg];
                    const unsigned long long dlo = tables_in_smem ? T.a[c] : R.dlo[seg];
                    const unsigned long long dhi = tables_in_smem ? T.b[c] : R.dhi[seg];
                    if (k < lo || k >= dhi) return;
                    const int shf = tables_in_smem ? (int)T.shift[c] : R.sh[seg];
                    unsigned long long bin = k <= dlo ? 0ull : ((k - dlo) >> shf);
                    if (bin > ROUND_NB - 1) bin = ROUND_NB - 1;
                    if (tables_in_smem) atomicAdd(&T.hist[c * ROUND_NB + (int)bin], 1u);
                    else atomicAdd(&R.hist[(size_t)seg * ROUND_NB + (int)bin], 1u);
                });
            }
            if (tables_in_smem) {
                __syncthreads();
                for (int i = tid; i < C * ROUND_NB; i += NMS_THREADS) {
                    const unsigned h = T.hist[i];
                    if (h) atomicAdd(&R.hist[(size_t)b * C * ROUND_NB + i], h);
                }
                __syncthreads();
            }
        });
        if (!grid_barrier(R.hdr, epoch)) { if (tid == 0) *R.err = SSDK_ASYNC_ROUNDS_TIMEOUT; return; }
        mark();

        // ---- (2) plan: one warp per pending segment (lane l owns bins l and l + 32)
        static_assert(ROUND_NB == 64, "the planning step maps two bins to every lane");
        for (int p = cta * (NMS_THREADS / 32) + (tid >> 5); p < npend; p += grid * (NMS_THREADS / 32)) {
            const int seg = R.pend_queue[p];
            if (R.st[seg] != ST_HIST) continue;                      // warp-uniform
            unsigned* h = R.hist + (size_t)seg * ROUND_NB;
            const long long h0 = h[lane], h1 = h[lane + 32];
            h[lane] = 0u;                                            // zero for the next histogram pass
            h[lane + 32] = 0u;
            long long c0 = h0, c1 = h1;                              // inclusive prefix sums over the 64 bins
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long v0 = __shfl_up_sync(0xffffffffu, c0, o), v1 = __shfl_up_sync(0xffffffffu, c1, o);
                if (lane >= o) { c0 += v0; c1 += v1; }
            }
            const long long sum0 = __shfl_sync(0xffffffffu, c0, 31);
            c1 += sum0;
            const long long total = __shfl_sync(0xffffffffu, c1, 31);
            const unsigned ne0 = __ballot_sync(0xffffffffu, h0 != 0), ne1 = __ballot_sync(0xffffffffu, h1 != 0);
            const int f = ne0 ? __ffs(ne0) - 1 : (ne1 ? 32 + __ffs(ne1) - 1 : -1);
            // first bin whose inclusive sum reaches ROUND_WANT / exceeds SEG_CAP
            const unsigned w0 = __ballot_sync(0xffffffffu, c0 >= ROUND_WANT), w1 = __ballot_sync(0xffffffffu, c1 >= ROUND_WANT);
            const unsigned x0 = __ballot_sync(0xffffffffu, c0 > SEG_CAP), x1 = __ballot_sync(0xffffffffu, c1 > SEG_CAP);
            const int j_want = w0 ? __ffs(w0) - 1 : (w1 ? 32 + __ffs(w1) - 1 : ROUND_NB);
            const int j_cap = x0 ? __ffs(x0) - 1 : (x1 ? 32 + __ffs(x1) - 1 : ROUND_NB);
            int j = j_want + 1 < j_cap ? j_want + 1 : j_cap;          // bins [f, j) are collected
            if (j > ROUND_NB) j = ROUND_NB;
            const long long cum_j = j == 0 ? 0 : (j <= 32 ? __shfl_sync(0xffffffffu, c0, j - 1) : __shfl_sync(0xffffffffu, c1, j - 33));
            const long long h_f = f < 0 ? 0 : (f < 32 ? __shfl_sync(0xffffffffu, h0, f) : __shfl_sync(0xffffffffu, h1, f - 32));
            if (lane != 0) continue;
            const unsigned long long lo = R.lo[seg], dlo = R.dlo[seg], dhi = R.dhi[seg];
            const int shf = R.sh[seg], b = seg / C;
            const long long beyond = R.beyond[seg];
            if (f < 0) {
                // nothing left inside the range
                if (beyond == 0) R.st[seg] = ST_DONE;
                else {                                              // (cannot happen after a refinement, whose bin is not empty; kept for safety)
                    R.lo[seg] = dhi; R.dlo[seg] = dhi; R.dhi[seg] = R.k_thr; R.sh[seg] = ceil_log2_per_bin(R.k_thr - dhi); R.beyond[seg] = 0;
                    atomicAdd(R.hdr + H_NHIST0 + (par ^ 1), 1);
                    if (atomicExch(R.stamp_h + b, round + 1) != round + 1) R.list_h[(size_t)(par ^ 1) * B + atomicAdd(R.hdr + H_NIMG_H0 + (par ^ 1), 1)] = b;
                }
            } else if (h_f > SEG_CAP) {
                // the best non-empty bin alone does not fit a region: narrow the range to that bin (the bins before it are empty)
                const unsigned long long nlo = f == 0 ? lo : dlo + ((unsigned long long)f << shf);
                unsigned long long nhi = f == ROUND_NB - 1 ? dhi : dlo + ((unsigned long long)(f + 1) << shf);
                if (nhi > dhi) nhi = dhi;
                R.lo[seg] = nlo; R.dlo[seg] = nlo; R.dhi[seg] = nhi; R.sh[seg] = ceil_log2_per_bin(nhi - nlo);
                R.beyond[seg] = (int)(beyond + total - h_f);
                atomicAdd(R.hdr + H_NHIST0 + (par ^ 1), 1);
                if (atomicExch(R.stamp_h + b, round + 1) != round + 1) R.list_h[(size_t)(par ^ 1) * B + atomicAdd(R.hdr + H_NIMG_H0 + (par ^ 1), 1)] = b;
            } else {
                unsigned long long mid = j >= ROUND_NB ? dhi : dlo + ((unsigned long long)j << shf);
                if (mid > dhi) mid = dhi;
                R.mid[seg] = mid;
                R.vcut[seg] = value_bound_of_key<IS_LOGITS>(mid, fmt.abits);
                R.rem[seg] = (int)(total - cum_j + beyond);
                R.fill[seg] = 0;
                R.st[seg] = ST_COLLECT;
                if (atomicExch(R.stamp_c + b, round) != round) R.list_c[atomicAdd(R.hdr + H_NIMG_C, 1)] = b;
            }
        }
        if (!grid_barrier(R.hdr, epoch)) { if (tid == 0) *R.err = SSDK_ASYNC_ROUNDS_TIMEOUT; return; }
        mark();

        // ---- (3) collect pass: the keys in [lo, mid) of every collecting segment -> its region
        const int n_img_c = R.hdr[H_NIMG_C];
        for_units(R.list_c, n_img_c, [&](int b, int slice, int group) {
            if (tables_in_smem) load_tables(b, ST_COLLECT);
            // nothing below the smallest per-class bound can be collected: one compare rejects ~99 % of a dense image
            const float lim_c = tables_in_smem ? fmaxf(IS_LOGITS ? x_lo : thr, T.vmin_all[0] - 1e-6f * fabsf(T.vmin_all[0])) : (IS_LOGITS ? x_lo : thr);
            for (int l = 0; l < G.num_levels; ++l) {
                const LevelGeom g = level_geom(G, l);
                const int blk = g.per_loc * g.C * g.hw;
                const int r0 = (int)((long long)blk * slice / group), r1 = (int)((long long)blk * (slice + 1) / group);
                const float* base = G.cls[l] + (size_t)b * blk;
                cta_scan_compact(base + r0, r1 - r0, lim_c, s_scan.queue[tid >> 5], [&](int r, float v) {
                    int a, c;
                    if (flat_geom) { a = (int)div_fast((unsigned)(r0 + r), g.d_c); c = r0 + r - a * C; }
                    else level_decompose(g, r0 + r, a, c);
                    const int seg = b * C + c;
                    if (tables_in_smem) {
                        if (!T.st[c] || v < T.vmin[c]) return;
                    } else if (R.st[seg] != ST_COLLECT || v < R.vcut[seg]) return;
                    float s;
                    if (!is_candidate<IS_LOGITS>(v, thr, x_lo, &s)) return;
                    const unsigned long long k = ((unsigned long long)order_desc(s) << fmt.abits) | (unsigned long long)a;
                    const unsigned long long lo = tables_in_smem ? T.lo[c] : R.lo[seg];
                    const unsigned long long mid = tables_in_smem ? T.a[c] : R.mid[seg];
                    if (k < lo || k >= mid) return;
                    const int pos = atomicAdd(R.fill + seg, 1);
                    if (pos < SEG_CAP) cand[(size_t)seg * SEG_CAP + pos] = ((unsigned long long)c << fmt.cshift) | k;
                });
            }
            __syncthreads();
        });
        if (!grid_barrier(R.hdr, epoch)) { if (tid == 0) *R.err = SSDK_ASYNC_ROUNDS_TIMEOUT; return; }
        mark();

        // ---- (4) the segments' NMS continues over the collected keys
        if (cta == 0 && tid == 0) R.hdr[H_NIMG_C] = 0;               // consumed by (3); produced again by the next round's (2)
        for (int p = cta; p < npend; p += grid) {
            const int seg = R.pend_queue[p];
            if (R.st[seg] != ST_COLLECT) continue;                   // CTA-uniform
            if (R.fill[seg] > SEG_CAP) {
                // (only after a guessed bound) more keys below `mid` than a region holds: the exact histogram over [lo, mid) decides
                if (tid == 0) {
                    const int b = seg / C;
                    const unsigned long long lo = R.lo[seg], mid = R.mid[seg];
                    R.st[seg] = ST_HIST;
                    R.dlo[seg] = lo; R.dhi[seg] = mid; R.sh[seg] = ceil_log2_per_bin(mid - lo); R.beyond[seg] = 1;   // >= 1 key lies beyond
                    atomicAdd(R.hdr + H_NHIST0 + (par ^ 1), 1);
                    if (atomicExch(R.stamp_h + b, round + 1) != round + 1) R.list_h[(size_t)(par ^ 1) * B + atomicAdd(R.hdr + H_NIMG_H0 + (par ^ 1), 1)] = b;
                }
                continue;
            }
            const int n = min(R.fill[seg], SEG_CAP);
            int kept = seg_kept[seg];
            __syncthreads();
            for (int i = tid; i < n; i += NMS_THREADS) s_sort[i] = cand[(size_t)seg * SEG_CAP + i] & kmask;
            for (int i = tid; i < kept; i += NMS_THREADS) {          // the boxes kept in earlier rounds
                float area;
                s_kept[i] = nms_box_of(N.seg_box[(size_t)seg * K + i], area);
                s_kept_area[i] = area;
            }
            __syncthreads();
            cta_sort_keys(s_sort, n, tid, NMS_THREADS);
            kept = nms_sorted_segment<DECODED>(N, sh, s_kept, s_kept_area, s_sort, n, seg, kept);
            if (tid == 0) {
                seg_kept[seg] = kept;
                const unsigned long long mid = R.mid[seg];
                if (kept >= K || R.rem[seg] <= 0 || mid >= R.k_thr) R.st[seg] = ST_DONE;
                else {
                    const int b = seg / C;
                    R.st[seg] = ST_HIST;
                    R.lo[seg] = mid; R.dlo[seg] = mid; R.dhi[seg] = R.k_thr; R.sh[seg] = ceil_log2_per_bin(R.k_thr - mid); R.beyond[seg] = 0;
                    atomicAdd(R.hdr + H_NHIST0 + (par ^ 1), 1);
                    if (atomicExch(R.stamp_h + b, round + 1) != round + 1) R.list_h[(size_t)(par ^ 1) * B + atomicAdd(R.hdr + H_NIMG_H0 + (par ^ 1), 1)] = b;
                }
            }
        }
        if (!grid_barrier(R.hdr, epoch)) { if (tid == 0) *R.err = SSDK_ASYNC_ROUNDS_TIMEOUT; return; }
        mark();
    }
}

// ---------------------------------------------------------------------------------------------- 5. pack
// Optional extra outputs of pack_kernel: the rows of the COCO results json that inference/evaluate_on_COCO.ipynb (cell 10) builds
// from the detector's output: boxes * [height, width, height, width] (float32), then x, y = int(xmin), int(ymin) and
// w, h = int(xmax - xmin), int(ymax - ymin) (Python int(): truncation towards zero), category_id through the notebook's
// integer_to_coco_id table, image_id.
struct PackCoco {
    const float2* image_sizes;     // [B] (height, width) in pixels; NULL = no COCO outputs
    const int* image_ids;          // [B] or NULL (then the image's index in the batch)
    const int* category_ids;       // [C] or NULL (then the class index)
    int4* out_xywh;                // [B, C*K]
    int* out_category;             // [B, C*K]
    int* out_image;                // [B, C*K]
};

#define PACK_SLOTS_PER_BLOCK 1024
__global__ void __launch_bounds__(256) pack_kernel(const float4* __restrict__ seg_box, const float* __restrict__ seg_score,
                                                   const int* __restrict__ seg_anchor, const int* __restrict__ seg_kept,
                                                   int C, int K, float4* __restrict__ out_boxes, float* __restrict__ out_scores,
                                                   int* __restrict__ out_classes, int* __restrict__ out_num,
                                                   int* __restrict__ out_anchor, const float4* __restrict__ box_scaler,
                                                   float final_thr, const PackCoco coco) {
    extern __shared__ int s_off[];   // [C+1] exclusive prefix sums of the per-class kept counts, then [C] the counts
    int* kept = s_off + C + 1;
    const int b = blockIdx.y;
    pdl_wait();                      // the per-segment results come from the NMS kernels
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
            if (coco.image_sizes) {
                const float2 hw = coco.image_sizes[b];
                const float ymin = f_mul(bx.x, hw.x), xmin = f_mul(bx.y, hw.y), ymax = f_mul(bx.z, hw.x), xmax = f_mul(bx.w, hw.y);
                coco.out_xywh[b * M + dst] = make_int4((int)xmin, (int)ymin, (int)f_sub(xmax, xmin), (int)f_sub(ymax, ymin));
                coco.out_category[b * M + dst] = coco.category_ids ? coco.category_ids[c] : c;
                coco.out_image[b * M + dst] = coco.image_ids ? coco.image_ids[b] : b;
            }
        }
        if (idx >= total) {                                           // zero padding (nms.py:84-89)
            ob[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
            os[idx] = 0.f;
            oc[idx] = 0;
            if (oa) oa[idx] = -1;
            if (coco.image_sizes) {
                coco.out_xywh[b * M + idx] = make_int4(0, 0, 0, 0);
                coco.out_category[b * M + idx] = -1;
                coco.out_image[b * M + idx] = -1;
            }
        }
    }
}

// Label-major packing: the per-label detection lists the reference's evaluator keeps (metrics.py:113-123, add_detections:
// `self.detections[label].append(get_box(box, image_name, score))` for every detection of every image in evaluation order).
// For label c the records are image 0's kept boxes of that class (descending score), then image 1's, ...  One CTA per label;
// the images' counts are scanned in chunks of 256.  out_* are [C, B*K] (the most a label can have), out_counts [C].
struct PackByLabel {
    float4* out_boxes; float* out_scores; int* out_image; int* out_counts;
    const int* image_ids;          // [B] or NULL
};
__global__ void __launch_bounds__(256) pack_by_label_kernel(const float4* __restrict__ seg_box, const float* __restrict__ seg_score,
                                                            const int* __restrict__ seg_kept, int B, int C, int K,
                                                            const float4* __restrict__ box_scaler, float final_thr, const PackByLabel P) {
    __shared__ int s_scan[256];
    __shared__ int s_base;
    const int c = blockIdx.x, tid = threadIdx.x;
    const size_t cap = (size_t)B * K;
    pdl_wait();
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < B; b0 += 256) {
        const int b = b0 + tid;
        int k = 0;
        if (b < B) {
            k = seg_kept[(size_t)b * C + c];
            if (final_thr > -INFINITY) {
                const float* sc = seg_score + ((size_t)b * C + c) * K;
                while (k > 0 && !(sc[k - 1] > final_thr)) --k;
            }
        }
        s_scan[tid] = k;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {                                  // inclusive scan (Hillis-Steele)
            const int v = tid >= o ? s_scan[tid - o] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int off = s_base + s_scan[tid] - k;
        for (int j = 0; j < k; ++j) {
            const size_t src = ((size_t)b * C + c) * K + j;
            float4 bx = seg_box[src];
            if (box_scaler) {
                const float4 sc4 = box_scaler[b];
                bx = make_float4(f_div(bx.x, sc4.x), f_div(bx.y, sc4.y), f_div(bx.z, sc4.z), f_div(bx.w, sc4.w));
            }
            P.out_boxes[(size_t)c * cap + off + j] = bx;
            P.out_scores[(size_t)c * cap + off + j] = seg_score[src];
            P.out_image[(size_t)c * cap + off + j] = P.image_ids ? P.image_ids[b] : b;
        }
        __syncthreads();
        if (tid == 255) s_base += s_scan[255];
        __syncthreads();
    }
    if (tid == 0) P.out_counts[c] = s_base;
}

// ---------------------------------------------------------------------------------------------- host side
static int bits_for(long long n) {   // bits needed to represent values in [0, n)
    int b = 1;
    while ((1ll << b) < n) ++b;
    return b;
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// shared memory of the filter kernels' saturation cache: one bit per segment, at most FILTER_SAT_WORDS words
static size_t sat_bytes(long long segments) {
    const long long bits = segments < (long long)FILTER_SAT_WORDS * 32 ? segments : (long long)FILTER_SAT_WORDS * 32;
    return (size_t)((bits + 31) / 32) * 4 + 4;
}

HeadGeom ssdk_flat_geom(const float* logits, const float* codes, int64_t A, int C);

static unsigned order_of_float(float s) {
    unsigned u;
    memcpy(&u, &s, 4);
    return order_desc_bits(u);
}

static int postprocess_impl(ssdk_ctx* ctx, const HeadGeom* head, const float* codes, const float* anchors, const float* scores, int flags,
                            int B, int64_t A, int C, double score_threshold, double iou_threshold, int K,
                            const float* box_scaler, double final_score_threshold,
                            float* out_boxes, float* out_scores, int32_t* out_classes, int32_t* out_num,
                            int32_t* out_anchor_idx, const PackCoco* coco = nullptr, const PackByLabel* by_label = nullptr) {
    SSDK_ENTER(ctx);
    ctx->round_times = nullptr;
    SSDK_REQUIRE(B >= 0 && A >= 0 && C > 0 && K > 0, SSDK_ERR_ARG, "ssdk_postprocess: bad sizes (B=%d A=%lld C=%d K=%d)", B,
                 (long long)A, C, K);
    SSDK_REQUIRE(B <= 65535, SSDK_ERR_SHAPE, "ssdk_postprocess: batch %d > 65535", B);
    // OP_REQUIRES of TensorFlow's NonMaxSuppressionV3 kernel (the op behind detector/utils/nms.py:33)
    SSDK_REQUIRE(iou_threshold >= 0.0 && iou_threshold <= 1.0, SSDK_ERR_ARG, "ssdk_postprocess: iou_threshold must be in [0, 1] (got %g)",
                 iou_threshold);
    if (B == 0) return SSDK_OK;
    SsdkWsGuard ws_guard(ctx, SSDK_WS_POST);
    const bool decoded = (flags & SSDK_BOXES_DECODED) != 0;
    const bool is_logits = (flags & SSDK_INPUT_LOGITS) != 0;
    const bool scan_only = (flags & SSDK_POST_SCAN_ONLY) != 0, finish_only = (flags & SSDK_POST_FINISH_ONLY) != 0;
    SSDK_REQUIRE(!(scan_only && finish_only), SSDK_ERR_ARG, "ssdk_postprocess: SSDK_POST_SCAN_ONLY and SSDK_POST_FINISH_ONLY exclude each other");
    SSDK_REQUIRE(scan_only || by_label || (out_boxes && out_scores && out_classes && out_num), SSDK_ERR_ARG, "ssdk_postprocess: null output");
    SSDK_REQUIRE(A == 0 || head || (codes && scores), SSDK_ERR_ARG, "ssdk_postprocess: null input");
    SSDK_REQUIRE(A == 0 || decoded || anchors, SSDK_ERR_ARG, "ssdk_postprocess: null anchors");
    SSDK_REQUIRE(!(head && decoded), SSDK_ERR_ARG, "ssdk_head_detect: head tensors hold encoded boxes");
    SSDK_REQUIRE(aligned16(codes) && aligned16(anchors) && aligned16(out_boxes), SSDK_ERR_SHAPE,
                 "ssdk_postprocess: box arrays must be 16-byte aligned");
    SSDK_REQUIRE(((uintptr_t)scores & 3) == 0, SSDK_ERR_SHAPE, "ssdk_postprocess: scores must be 4-byte aligned");
    const long long per_image = (long long)A * C;
    SSDK_REQUIRE(per_image < (1ll << 31), SSDK_ERR_SHAPE, "ssdk_postprocess: A*C must be < 2^31");
    SSDK_REQUIRE((size_t)K * 20 <= 96 * 1024, SSDK_ERR_SHAPE, "ssdk_postprocess: max_boxes_per_class %d too large", K);
    KeyFormat fmt;
    fmt.abits = bits_for(A > 1 ? A : 2);
    fmt.cshift = 32 + fmt.abits;
    SSDK_REQUIRE(fmt.cshift + bits_for(C > 1 ? C : 2) <= 64, SSDK_ERR_SHAPE, "ssdk_postprocess: A=%lld x C=%d does not fit the key",
                 (long long)A, C);
    const long long nseg = (long long)B * C;
    SSDK_REQUIRE(nseg < (1ll << 31), SSDK_ERR_SHAPE, "ssdk_postprocess: B*C must be < 2^31");

    // workspace: one BOUNDED candidate region of min(A, SEG_CAP) keys per (image, class) segment, one counter per segment
    // (zeroed every call), the queues, the per-segment NMS results, and the state of the rounds (dense segments)
    const bool may_overflow = A > SEG_CAP;                               // a segment can only overflow when there are that many anchors
    SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_cand, (size_t)nseg * SEG_CAP * sizeof(unsigned long long)));
    // dense images (anchor-major tensors only): redone by filter_dense_kernel with its own counters and value histograms
    const bool use_dense = may_overflow && !head && C <= 1024;
    // int words zeroed by ONE memset per call: header | round stamps [2B] | dense-image flags [B] | seg_count [nseg];
    // seg_count_dense [nseg] and the value histograms [nseg][FD_NB] of a dense image are zeroed by the CTA that flags it
    const size_t n_stamp = may_overflow ? 2 * (size_t)B : 0;
    const size_t n_dense = use_dense ? (size_t)nseg + (size_t)nseg * FD_NB : 0;
    const size_t n_zero = H_WORDS + n_stamp + (use_dense ? (size_t)B : 0) + (size_t)nseg;
    const size_t n_int = n_zero + n_dense + 3 * (size_t)nseg + (may_overflow ? 6 * (size_t)nseg + 3 * (size_t)B + 4 : 0);
    const size_t n_u64 = may_overflow ? 4 * (size_t)nseg + ROUND_TIMES : 0;
    const size_t n_hist = may_overflow ? (size_t)nseg * ROUND_NB : 0;
    SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_counts, 16 + n_u64 * 8 + (n_int + n_hist) * 4));
    const size_t seg_elems = (size_t)nseg * K;
    SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_seg, seg_elems * (sizeof(float4) + sizeof(float) + sizeof(int))));
    unsigned long long* cand = (unsigned long long*)ctx->ws_cand.p;
    unsigned long long* w64 = (unsigned long long*)ctx->ws_counts.p;
    int* hdr = (int*)(w64 + n_u64);
    int* stamps = hdr + H_WORDS;
    int* img_dense = use_dense ? stamps + n_stamp : nullptr;
    int* seg_count = stamps + n_stamp + (use_dense ? B : 0);
    int* seg_count_dense = use_dense ? seg_count + nseg : nullptr;
    unsigned* vhist = use_dense ? (unsigned*)(seg_count_dense + nseg) : nullptr;
    int* seg_kept = hdr + n_zero + n_dense;
    int* heavy_queue = seg_kept + nseg;
    int* pend_queue = heavy_queue + nseg;
    float4* seg_box = (float4*)ctx->ws_seg.p;
    float* seg_score = (float*)(seg_box + seg_elems);
    int* seg_anchor = (int*)(seg_score + seg_elems);
    if (!finish_only) {
        SSDK_CHECK_CUDA(cudaMemsetAsync(hdr, 0, n_zero * sizeof(int), ctx->stream));
        if (per_image == 0) SSDK_CHECK_CUDA(cudaMemsetAsync(seg_kept, 0, (size_t)nseg * sizeof(int), ctx->stream));
    }

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
        if (finish_only) {
            // the scan of this very call was enqueued by the SSDK_POST_SCAN_ONLY call before
        } else if (head) {
            // CTAs per level in proportion to its tiles, num_sms * 16 in total (at least one per level)
            long long tiles[SSDK_MAX_LEVELS], total = 0;
            for (int l = 0; l < head->num_levels; ++l) {
                const long long cnt = (long long)B * head->per_loc * C * head->hw[l];
                tiles[l] = (cnt / 4 + FILTER_THREADS * SCAN_U_HEAD - 1) / (FILTER_THREADS * SCAN_U_HEAD);
                if (tiles[l] < 1) tiles[l] = 1;
                total += tiles[l];
            }
            long long budget = (long long)ctx->num_sms * 16;
            if (budget > total) budget = total;
            HeadCtas HC;
            int next = 0;
            for (int l = 0; l <= SSDK_MAX_LEVELS; ++l) HC.cta0[l] = 0;
            for (int l = 0; l < head->num_levels; ++l) {
                long long n = (tiles[l] * budget + total - 1) / total;
                if (n < 1) n = 1;
                if (n > tiles[l]) n = tiles[l];
                HC.cta0[l] = next;
                next += (int)n;
            }
            for (int l = head->num_levels; l <= SSDK_MAX_LEVELS; ++l) HC.cta0[l] = next;
            const dim3 hgrid_f((unsigned)next);
            const size_t sat_head = sat_bytes(nseg);
            SSDK_KERNEL(ctx, SSDK_K_FILTER,
                if (is_logits) head_filter_kernel<true><<<hgrid_f, FILTER_THREADS, sat_head, ctx->stream>>>(*head, B, thr, x_lo, fmt, cand, seg_count, HC);
                else head_filter_kernel<false><<<hgrid_f, FILTER_THREADS, sat_head, ctx->stream>>>(*head, B, thr, x_lo, fmt, cand, seg_count, HC));
        } else {
            long long chunks = (per_image / 4 + FILTER_THREADS * SCAN_U - 1) / (FILTER_THREADS * SCAN_U);
            long long gx = ((long long)ctx->num_sms * 16 + B - 1) / B;
            if (gx > chunks) gx = chunks;
            if (gx < 1) gx = 1;
            const dim3 fgrid((unsigned)gx, B);
            const size_t sat_img = sat_bytes(C);
            SSDK_KERNEL(ctx, SSDK_K_FILTER,
                if (is_logits) filter_kernel<true><<<fgrid, FILTER_THREADS, sat_img, ctx->stream>>>(scores, per_image, C, thr, x_lo, fmt, cand, seg_count, img_dense, seg_count_dense, vhist);
                else filter_kernel<false><<<fgrid, FILTER_THREADS, sat_img, ctx->stream>>>(scores, per_image, C, thr, x_lo, fmt, cand, seg_count, img_dense, seg_count_dense, vhist));
            if (use_dense) {
                // images that filter_kernel found dense are redone with per-tile aggregation; a no-op otherwise (a small
                // persistent grid: six CTAs per SM in total, each walks its share of the image's tiles)
                long long dgx = ((long long)ctx->num_sms * 6 + B - 1) / B;
                if (dgx > chunks) dgx = chunks;
                if (dgx < 1) dgx = 1;
                const dim3 dgrid((unsigned)dgx, B);
                const size_t dsmem = ((size_t)2 * C + (size_t)(C + 31) / 32 + (size_t)C * FD_NB) * sizeof(int);
                SSDK_TRY(ssdk_set_max_smem(ctx, is_logits ? (const void*)filter_dense_kernel<true> : (const void*)filter_dense_kernel<false>, (int)dsmem));
                SSDK_KERNEL(ctx, SSDK_K_SORT,
                    if (is_logits) ssdk_launch(ctx, true, filter_dense_kernel<true>, dgrid, dim3(FILTER_THREADS), dsmem, scores, per_image, C, thr, x_lo, fmt, cand, seg_count_dense, (const int*)img_dense, vhist);
                    else ssdk_launch(ctx, true, filter_dense_kernel<false>, dgrid, dim3(FILTER_THREADS), dsmem, scores, per_image, C, thr, x_lo, fmt, cand, seg_count_dense, (const int*)img_dense, vhist));
            }
        }

        if (scan_only) return SSDK_OK;

        // 2.-3. NMS: one warp per small segment (<= 32 candidates, sorted with shuffles), then one CTA per queued heavy
        //       segment (sorted by the CTA)
        const size_t nms_smem = (size_t)SEG_CAP * sizeof(unsigned long long) + (size_t)K * (sizeof(NmsBox) + sizeof(float));
        const int sgrid_nms = ceil_div_i(nseg, NMS_SMALL_WARPS);
        long long hgrid = (long long)ctx->num_sms * 4;
        if (hgrid > nseg) hgrid = nseg;
        NmsSegArgs N;
        memset(&N, 0, sizeof(N));
        N.fmt = fmt;
        N.codes.flat = head ? nullptr : (const float4*)codes;
        if (head) N.codes.head = *head;
        N.anchors = (const float4*)anchors;
        N.A = A; N.C = C; N.K = K;
        N.iou_thr = (float)iou_threshold;
        N.seg_box = seg_box; N.seg_score = seg_score; N.seg_anchor = seg_anchor;
        SegCounts SC;
        SC.sparse = seg_count; SC.dense = seg_count_dense; SC.img_dense = img_dense;
        const int nms_slot = ctx->profiling ? ssdk_prof_begin(ctx, SSDK_K_NMS) : -1;
#define SSDK_LAUNCH_NMS(DEC)                                                                                                  \
        do {                                                                                                                  \
            SSDK_TRY(ssdk_set_max_smem(ctx, (const void*)nms_kernel<DEC, NMS_HEAVY_THREADS>, (int)nms_smem));                                   \
            ssdk_launch(ctx, true, nms_small_kernel<DEC>, dim3(sgrid_nms), dim3(NMS_SMALL_WARPS * 32), 0,                    \
                        (const unsigned long long*)cand, fmt, SC, N.codes, N.anchors, (long long)A, nseg, C, K, N.iou_thr,    \
                        seg_box, seg_score, seg_anchor, seg_kept, heavy_queue, pend_queue, hdr);                              \
            ssdk_launch(ctx, true, nms_kernel<DEC, NMS_HEAVY_THREADS>, dim3((unsigned)hgrid), dim3(NMS_HEAVY_THREADS), nms_smem,                      \
                        (const unsigned long long*)cand, SC, N, seg_kept, (const int*)heavy_queue, (const int*)hdr);          \
        } while (0)
        if (decoded) SSDK_LAUNCH_NMS(true);
        else SSDK_LAUNCH_NMS(false);
#undef SSDK_LAUNCH_NMS
        if (nms_slot >= 0) ssdk_prof_end(ctx, nms_slot);
        ctx->launches++;
        SSDK_CHECK_LAUNCH(ctx);

        // 4. rounds: segments that overflowed their region (dense scores); exits at once when there is none
#ifdef SSDK_VARIANT_NO_ROUNDS
        if (false) {
#else
        if (may_overflow) {
#endif
            RoundState R;
            memset(&R, 0, sizeof(R));
            R.hdr = hdr;
            R.pend_queue = pend_queue;
            int* ip = pend_queue + nseg;
            R.st = ip; ip += nseg;
            R.sh = ip; ip += nseg;
            R.fill = ip; ip += nseg;
            R.rem = ip; ip += nseg;
            R.beyond = ip; ip += nseg;
            R.vcut = (float*)ip; ip += nseg;
            R.vhist = vhist;
            R.img_dense = img_dense;
            R.vlim = is_logits ? x_lo : thr;
            R.vwidth = FD_RANGE / (float)FD_NB;
            R.stamp_h = stamps;
            R.stamp_c = R.stamp_h + B;
            R.list_h = ip; ip += 2 * (size_t)B;
            R.list_c = ip; ip += B;
            R.hist = (unsigned*)(((uintptr_t)ip + 15) & ~(uintptr_t)15);
            R.lo = w64; R.mid = w64 + nseg; R.dlo = w64 + 2 * nseg; R.dhi = w64 + 3 * nseg;
            R.times = w64 + 4 * nseg;
            ctx->round_times = R.times;
            const unsigned o_thr = order_of_float(thr), o_one = order_of_float(1.0f);
            R.k_thr = (unsigned long long)o_thr << fmt.abits;            // candidates: score > thr <=> order word < o_thr
            R.err = ctx->dev_err;
            R.dlo0 = o_one < o_thr ? ((unsigned long long)o_one << fmt.abits) : 0ull;   // scores are normally <= 1: resolve [1.0, thr)
            const HeadGeom G = head ? *head : ssdk_flat_geom(scores, codes, A, C);
            const bool tables = C <= ROUND_SMEM_MAX_C;
            size_t rsmem = nms_smem + 16;
            if (tables) rsmem += (size_t)C * (3 * 8 + ROUND_NB * 4 + 4 + 2) + 32;
            const void* fn = decoded ? (is_logits ? (const void*)nms_rounds_kernel<true, true> : (const void*)nms_rounds_kernel<true, false>)
                                     : (is_logits ? (const void*)nms_rounds_kernel<false, true> : (const void*)nms_rounds_kernel<false, false>);
            SSDK_TRY(ssdk_set_max_smem(ctx, fn, (int)rsmem));
            int occ = 0;
            SSDK_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, NMS_THREADS, rsmem));
            SSDK_REQUIRE(occ >= 1, SSDK_ERR_SHAPE, "ssdk_postprocess: %d classes x %d boxes per class do not fit the dense-segment kernel", C, K);
            if (occ > 2) occ = 2;
            const int rgrid = ctx->num_sms * occ;                        // all CTAs must be co-resident (grid-wide barriers)
            SSDK_KERNEL(ctx, SSDK_K_NMS_ROUNDS,
                if (decoded && is_logits) ssdk_launch(ctx, true, nms_rounds_kernel<true, true>, dim3(rgrid), dim3(NMS_THREADS), rsmem, G, B, thr, x_lo, cand, N, seg_kept, R);
                else if (decoded) ssdk_launch(ctx, true, nms_rounds_kernel<true, false>, dim3(rgrid), dim3(NMS_THREADS), rsmem, G, B, thr, x_lo, cand, N, seg_kept, R);
                else if (is_logits) ssdk_launch(ctx, true, nms_rounds_kernel<false, true>, dim3(rgrid), dim3(NMS_THREADS), rsmem, G, B, thr, x_lo, cand, N, seg_kept, R);
                else ssdk_launch(ctx, true, nms_rounds_kernel<false, false>, dim3(rgrid), dim3(NMS_THREADS), rsmem, G, B, thr, x_lo, cand, N, seg_kept, R));
        }
    }
    if (scan_only) return SSDK_OK;
    // 5. pack
    if (by_label) {
        SSDK_KERNEL(ctx, SSDK_K_PACK,
                    ssdk_launch(ctx, per_image > 0, pack_by_label_kernel, dim3(C), dim3(256), 0, (const float4*)seg_box, (const float*)seg_score,
                                (const int*)seg_kept, B, C, K, (const float4*)box_scaler, (float)final_score_threshold, *by_label));
        return SSDK_OK;
    }
    PackCoco pc;
    memset(&pc, 0, sizeof(pc));
    if (coco) pc = *coco;
    SSDK_KERNEL(ctx, SSDK_K_PACK,
                ssdk_launch(ctx, per_image > 0, pack_kernel, dim3(ceil_div_i((long long)C * K, PACK_SLOTS_PER_BLOCK), B), dim3(256),
                            (size_t)(2 * C + 1) * sizeof(int), (const float4*)seg_box, (const float*)seg_score, (const int*)seg_anchor,
                            (const int*)seg_kept, C, K, (float4*)out_boxes, out_scores, out_classes, out_num, out_anchor_idx,
                            (const float4*)box_scaler, (float)final_score_threshold, pc));
    return SSDK_OK;
}

// Phase timestamps of the dense-segment rounds of the LAST post-processing call on this context (synchronises): out[0] = rounds
// run (0: no segment overflowed), out[1..] = nanosecond timestamps: kernel start, after init, then per round: start, after the
// histogram pass, after planning, after the collect pass, after the NMS phase.
extern "C" int ssdk_ctx_round_times(ssdk_ctx* ctx, int64_t* out, int n) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(out && n >= 1, SSDK_ERR_ARG, "ssdk_ctx_round_times: bad arguments");
    for (int i = 0; i < n; ++i) out[i] = 0;
    if (!ctx->round_times) return SSDK_OK;
    SSDK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    SSDK_CHECK_CUDA(cudaMemcpy(out, ctx->round_times, sizeof(int64_t) * (size_t)(n < ROUND_TIMES ? n : ROUND_TIMES), cudaMemcpyDeviceToHost));
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

extern "C" int ssdk_detect_coco(ssdk_ctx* ctx, const float* codes, const float* anchors, const float* scores, int flags, int B,
                                int64_t A, int C, double score_threshold, double iou_threshold, int K, const float* box_scaler,
                                double final_score_threshold, const float* image_sizes, const int32_t* image_ids,
                                const int32_t* category_ids, float* out_boxes, float* out_scores, int32_t* out_classes,
                                int32_t* out_num, int32_t* out_bbox_xywh, int32_t* out_category_id, int32_t* out_image_id) {
    SSDK_REQUIRE(box_scaler == nullptr || aligned16(box_scaler), SSDK_ERR_SHAPE, "ssdk_detect_coco: box_scaler must be 16-byte aligned");
    SSDK_REQUIRE(image_sizes && out_bbox_xywh && out_category_id && out_image_id, SSDK_ERR_ARG, "ssdk_detect_coco: null pointer");
    SSDK_REQUIRE(((uintptr_t)image_sizes & 7) == 0 && aligned16(out_bbox_xywh), SSDK_ERR_SHAPE,
                 "ssdk_detect_coco: image_sizes must be 8-byte aligned, out_bbox_xywh 16-byte aligned");
    PackCoco pc;
    pc.image_sizes = (const float2*)image_sizes; pc.image_ids = image_ids; pc.category_ids = category_ids;
    pc.out_xywh = (int4*)out_bbox_xywh; pc.out_category = out_category_id; pc.out_image = out_image_id;
    return postprocess_impl(ctx, nullptr, codes, anchors, scores, flags, B, A, C, score_threshold, iou_threshold, K, box_scaler,
                            final_score_threshold, out_boxes, out_scores, out_classes, out_num, nullptr, &pc, nullptr);
}

extern "C" int ssdk_detect_by_label(ssdk_ctx* ctx, const float* codes, const float* anchors, const float* scores, int flags, int B,
                                    int64_t A, int C, double score_threshold, double iou_threshold, int K, const float* box_scaler,
                                    double final_score_threshold, const int32_t* image_ids, float* out_boxes, float* out_scores,
                                    int32_t* out_image, int32_t* out_counts) {
    SSDK_REQUIRE(box_scaler == nullptr || aligned16(box_scaler), SSDK_ERR_SHAPE, "ssdk_detect_by_label: box_scaler must be 16-byte aligned");
    SSDK_REQUIRE(out_boxes && out_scores && out_image && out_counts, SSDK_ERR_ARG, "ssdk_detect_by_label: null output");
    SSDK_REQUIRE(aligned16(out_boxes), SSDK_ERR_SHAPE, "ssdk_detect_by_label: out_boxes must be 16-byte aligned");
    PackByLabel pl;
    pl.out_boxes = (float4*)out_boxes; pl.out_scores = out_scores; pl.out_image = out_image; pl.out_counts = out_counts;
    pl.image_ids = image_ids;
    return postprocess_impl(ctx, nullptr, codes, anchors, scores, flags, B, A, C, score_threshold, iou_threshold, K, box_scaler,
                            final_score_threshold, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &pl);
}
