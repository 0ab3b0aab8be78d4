// The training-side hot path as ONE kernel launch: target assignment (detector/ssd.py:84 -> training_target_creation.py:5-176),
// focal + smooth-L1 losses with their weights (ssd.py:89-117, losses.py:4-50), the matched count and the normalisation
// (ssd.py:121-133) and -- on several GPUs -- the all-reduce of the three sums over NVLink peer memory.
//
// Round 1 ran this as matcher || flat pass on two streams, then head_rows_kernel, then a finalize (and a comm) launch: five
// nodes with fork / join events, and the co-running matcher finished AFTER the flat pass (0.155 ms per 16 images, 0.69 of the
// HBM roofline).  Here the CTAs of one persistent grid take ROLES:
//   * CTAs [0, n_match) start as matchers (matcher.cuh): ALU-bound work (IEEE divides, argmax bookkeeping; < 1 % DRAM) that
//     runs in the issue slots the streaming CTAs of the same SM leave idle.  The loss contributions of the few matched /
//     ignored anchors (smooth-L1, count, the positive-class and ignore corrections of head.cu) are added right where the match
//     is decided -- the [B,A] matches array is never read back, head_rows_kernel is gone from this path; the CTA that finishes
//     an image applies the forced matches (training_target_creation.py:105-126) and books the resulting CHANGE of the sums.
//   * all other CTAs stream the class logits (flat.cuh: 16 KB chunks, 128-bit no-allocate loads, e^3 h(e) fast path).  By
//     default the chunks are handed out DYNAMICALLY (one atomic per four chunks): the matcher CTAs start claiming the moment
//     their matching is done, so every CTA slot of the GPU streams until the end whatever the ratio of the two kinds of work,
//     and the flat sum is added in 2^-32 fixed point (integer addition is associative), so it is bit-identical from run to run
//     although the chunk -> CTA assignment is not.  Option SSDK_OPT_TRAIN_DYNAMIC_CHUNKS = 0: the static split (the matcher CTAs
//     take a fixed share of the chunk list, double accumulation in a fixed order).
//   * every CTA writes its partial sums; the CTA that draws the last ticket adds them in a fixed order, performs the peer-memory
//     all-reduce (comm.cuh) when asked to, and writes sums and losses.  It also leaves the workspace zeroed for the next launch,
//     so the step is exactly one graph node.
#include <stdlib.h>

#include "comm.cuh"
#include "flat.cuh"
#include "matcher.cuh"

struct TrainStepArgs {
    MatchArgs M;
    FlatSegs S;
    HeadGeom G;
    int B;
    int n_match;                  // CTAs [0, n_match) start in the matcher role
    int gx;                       // matcher work items per image
    long long rounds_all;         // flat rounds in which every CTA takes a chunk (the matcher CTAs' share of the streaming)
    float gamma, alpha;
    double* partials;             // [grid][4]: flat sum, sum loc, class corrections, matched count
    double* img_partials;         // [B][3]: change of (sum loc, class corrections, matched count) by the forced matches; zero before the launch
    unsigned* ticket;             // zero before the launch
    unsigned long long* flat_counter;   // DYN: next unclaimed chunk of the flat pass; zero before the launch
    long long* fx_partials;       // DYN: [grid] fixed-point flat sums
    double* out_sums;             // [3]
    float* out_losses;            // [2] or NULL
    CommPeers P;
    int use_comm;
};

// ---------------------------------------------------------------------------------------------- matched / ignored anchors
// Contribution of ONE matched anchor (ssd.py:89-117): .x = smooth-L1 against its target, .y = the positive-class correction --
// its logit is summed as a negative by the flat pass, so the correction is alpha * pos(x) - (1 - alpha) * neg(x).
template <int GAMMA_MODE>
__device__ __noinline__ float2 matched_contribution(const HeadGeom& G, int b, int a, float4 target, int tc, float gamma, float alpha) {
    const int l = head_level_of(G, a);
    const int r = a - G.anchor_off[l];
    const int n = G.per_loc, loc = r / n, k = r - loc * n, hw = G.hw[l], cf = G.channels_first;
    const long long es = cf ? hw : 1;
    const float* pb = G.box[l] + head_elem(cf, b, n * 4, hw, k * 4, loc);
    const float4 p = make_float4(__ldg(pb), __ldg(pb + es), __ldg(pb + 2 * es), __ldg(pb + 3 * es));
    float2 out;
    out.x = smooth_l1_4(p, target);                                            // losses.py:4-19
    out.y = 0.0f;
    if (tc >= 0 && tc < G.C) {                                                 // one_hot(cls, C+1)[1:] (ssd.py:96-100)
        const float x = __ldg(G.cls[l] + head_elem(cf, b, n * G.C, hw, k * G.C + tc, loc));
        out.y = alpha * focal_positive<GAMMA_MODE>(x, gamma) - (1.0f - alpha) * focal_negative<GAMMA_MODE>(x, gamma);
    }
    return out;
}

// Sum over all classes of the negative-class term of ONE ignored anchor (matches == -2: weight 0, ssd.py:103; the flat pass
// summed every class).  `nlanes` = 32: warp-cooperative (lanes over classes, fixed shuffle tree, every lane returns the sum);
// `nlanes` = 1: by the calling thread alone (forced-match bookkeeping: a handful of anchors per image at most).
template <int GAMMA_MODE>
__device__ __noinline__ float ignored_row_sum(const HeadGeom& G, int b, int a, float gamma, int nlanes) {
    const int lane = nlanes == 32 ? (int)(threadIdx.x & 31) : 0;
    const int l = head_level_of(G, a);
    const int r = a - G.anchor_off[l];
    const int n = G.per_loc, loc = r / n, k = r - loc * n, hw = G.hw[l], cf = G.channels_first;
    const long long es = cf ? hw : 1;
    const float* px = G.cls[l] + head_elem(cf, b, n * G.C, hw, k * G.C, loc);
    float sub = 0.0f;
    for (int c = lane; c < G.C; c += nlanes) sub += focal_negative<GAMMA_MODE>(__ldg(px + c * es), gamma);
    if (nlanes == 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sub += __shfl_xor_sync(0xffffffffu, sub, o);
    }
    return sub;
}

// Shared-memory scratch of the forced-match bookkeeping: per ground-truth box g of the image being finished, the change of
// (sum loc, class corrections, matched count) its forced match causes.  Lives in the tail of MatchSmem::box, which
// force_match_image uses only for its first GT_CHUNK bytes (s_ok).
struct ForcedDeltas {
    float loc[GT_CHUNK], fix[GT_CHUNK], cnt[GT_CHUNK];
};
static_assert(sizeof(ForcedDeltas) + 1024 <= sizeof(float4) * GT_CHUNK, "forced-match scratch must fit behind s_ok in MatchSmem::box");

template <int GAMMA_MODE>
struct LossHook {
    const TrainStepArgs& T;
    double acc_loc, acc_fix;                  // this thread's share of the matched / ignored anchors' contributions
    int acc_cnt;
    ForcedDeltas* fd;
    double* s_scratch;                        // [MATCH_THREADS / 32][3] shared

    __device__ __forceinline__ LossHook(const TrainStepArgs& t, MatchSmem& sm, double* scratch)
        : T(t), acc_loc(0.0), acc_fix(0.0), acc_cnt(0), fd((ForcedDeltas*)((char*)sm.box + 1024)), s_scratch(scratch) {}

    // warp-collective: every lane of the warp calls with its own anchor
    __device__ __forceinline__ void anchors(int b, int a, bool valid, int m, const float4&, const float4& target, int label1) {
        const unsigned special = __ballot_sync(0xffffffffu, valid && m != -1);
        if (special == 0u) return;                                             // background anchors: the overwhelming majority
        if (valid && m >= 0) {
            const float2 c = matched_contribution<GAMMA_MODE>(T.G, b, a, target, label1 - 1, T.gamma, T.alpha);
            acc_loc += (double)c.x;
            acc_fix += (double)c.y;
            acc_cnt += 1;
        }
        for (unsigned todo = __ballot_sync(0xffffffffu, valid && m == -2); todo;) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const int aj = __shfl_sync(0xffffffffu, a, j);
            const float sub = ignored_row_sum<GAMMA_MODE>(T.G, b, aj, T.gamma, 32);
            if ((int)(threadIdx.x & 31) == j) acc_fix -= (double)((1.0f - T.alpha) * sub);
        }
    }

    __device__ __forceinline__ void forced_init(int g) { fd->loc[g] = 0.0f; fd->fix[g] = 0.0f; fd->cnt[g] = 0.0f; }

    // one thread: anchor a of image b goes from m_old (threshold result) to the forced match g
    __device__ __forceinline__ void forced(int b, int a, int m_old, int g, const float4& anc) {
        const MatchArgs& M = T.M;
        float loc = 0.0f, fix = 0.0f, cnt = 1.0f;
        if (m_old >= 0) {                                                     // undo what anchors() added for the old match
            const float4 t_old = box_encode(M.gt_boxes[(size_t)b * M.Gmax + m_old], anc);
            const float2 c = matched_contribution<GAMMA_MODE>(T.G, b, a, t_old, M.gt_labels[(size_t)b * M.Gmax + m_old], T.gamma, T.alpha);
            loc = -c.x; fix = -c.y; cnt = 0.0f;
        } else if (m_old == -2) {
            fix = (1.0f - T.alpha) * ignored_row_sum<GAMMA_MODE>(T.G, b, a, T.gamma, 1);
        }
        const float4 t_new = box_encode(M.gt_boxes[(size_t)b * M.Gmax + g], anc);
        const float2 c = matched_contribution<GAMMA_MODE>(T.G, b, a, t_new, M.gt_labels[(size_t)b * M.Gmax + g], T.gamma, T.alpha);
        fd->loc[g] = loc + c.x;
        fd->fix[g] = fix + c.y;
        fd->cnt[g] = cnt;
    }

    // all threads of the CTA that finished image b (N boxes): the image's forced-match deltas -> img_partials[b] (fixed order)
    __device__ __forceinline__ void image_done(int b, int N) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        double v[3] = {0.0, 0.0, 0.0};
        __syncthreads();
        for (int g = threadIdx.x; g < N; g += MATCH_THREADS) { v[0] += (double)fd->loc[g]; v[1] += (double)fd->fix[g]; v[2] += (double)fd->cnt[g]; }
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
        if (lane == 0) { s_scratch[warp * 3 + 0] = v[0]; s_scratch[warp * 3 + 1] = v[1]; s_scratch[warp * 3 + 2] = v[2]; }
        __syncthreads();
        if (threadIdx.x < 3) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < MATCH_THREADS / 32; ++w) t += s_scratch[w * 3 + threadIdx.x];
            T.img_partials[(size_t)b * 3 + threadIdx.x] = t;
        }
    }
};

// ---------------------------------------------------------------------------------------------- the kernel
// 40 registers: six resident CTAs per SM; the few spills this costs lie around the calls of the matched-anchor functions and at
// the role switch, not in the matching or streaming loops (checked in the SASS)
#ifndef TRAIN_MIN_CTAS
#define TRAIN_MIN_CTAS 6
#endif
union TrainSmem {
    MatchSmem match;
    double fin[FLAT_THREADS][4];
};

// DYN: the chunks of the flat pass are handed out dynamically (FLAT_CLAIM at a time, one atomic per claim, the next claim in
// flight while the current one is worked off) instead of by the static split: whoever is free takes the next chunks -- the
// matcher CTAs as soon as their matching is done, and CTAs that found their SM occupied by another kernel when the grid
// started (a second sub-path running next to this one) simply take fewer.  The flat sum is accumulated in fixed point
// (FlatAccFixed), so it is still bit-identical from run to run.
#define FLAT_CLAIM 4
template <int GAMMA_MODE, bool DYN>
__global__ void __launch_bounds__(FLAT_THREADS, TRAIN_MIN_CTAS) train_step_kernel(const __grid_constant__ TrainStepArgs T) {
    static_assert(FLAT_THREADS == MATCH_THREADS, "both roles use the same CTA shape");
    __shared__ TrainSmem sm;
    __shared__ double s_red[FLAT_THREADS / 32][4];
    __shared__ double s_hook[(MATCH_THREADS / 32) * 3];
    __shared__ int s_last;
    __shared__ long long s_claim[2];
    __shared__ long long s_fx[FLAT_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
    const bool matcher = cta < T.n_match;
    const int n_flat = grid - T.n_match;
    double acc_flat = 0.0;
    long long acc_fx = 0;
    LossHook<GAMMA_MODE> hook(T, sm.match, s_hook);

    // ---- role 1: target assignment + the matched / ignored anchors' contributions
    if (matcher) {
        const int items = T.gx * T.B;
        for (int w = cta; w < items; w += T.n_match) match_work_item<true>(T.M, sm.match, w / T.gx, w % T.gx, T.gx, hook);
    }

    // ---- role 2: the flat pass.  DYN: claims of FLAT_CLAIM chunks.  Static split: rounds [0, rounds_all): every CTA takes chunk
    //      round * grid + cta; later rounds: only the streaming CTAs, chunk rounds_all * grid + (round - rounds_all) * n_flat + (cta - n_match).
    if (DYN) {
        const long long total = T.S.chunk0[T.S.nseg];
        FlatAccFixed acc;
        long long next = 0;
        if (tid == 0) s_claim[0] = (long long)atomicAdd(T.flat_counter, (unsigned long long)FLAT_CLAIM);
        __syncthreads();
        int buf = 0;
        while (true) {
            const long long g0 = s_claim[buf];
            if (g0 >= total) break;
            if (tid == 0) next = (long long)atomicAdd(T.flat_counter, (unsigned long long)FLAT_CLAIM);   // used only after this claim's work
            long long g1 = g0 + FLAT_CLAIM;
            if (g1 > total) g1 = total;
            flat_sum_range<GAMMA_MODE>(T.S, acc, g0, g1, 1, T.gamma, tid);
            if (tid == 0) s_claim[buf ^ 1] = next;
            __syncthreads();
            buf ^= 1;
        }
        acc_flat = acc.big;
        acc_fx = acc.fx;
    } else {
        const long long total = T.S.chunk0[T.S.nseg];
        long long all_end = T.rounds_all * grid;
        if (all_end > total) all_end = total;
#pragma unroll 1
        for (int phase = 0; phase < (matcher ? 1 : 2); ++phase) {
            const long long g0 = phase == 0 ? (long long)cta : all_end + (cta - T.n_match);
            const long long g1 = phase == 0 ? all_end : total;
            const long long st = phase == 0 ? (long long)grid : (long long)n_flat;
            acc_flat += flat_sum_range<GAMMA_MODE>(T.S, g0, g1, st, T.gamma, tid);
        }
    }

    // ---- CTA reduction (fixed order) -> partials[cta]
    double v[4] = {acc_flat, hook.acc_loc, hook.acc_fix, (double)hook.acc_cnt};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    if (DYN) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc_fx += __shfl_xor_sync(0xffffffffu, acc_fx, o);
        if (lane == 0) s_fx[warp] = acc_fx;
    }
    if (lane == 0)
        for (int j = 0; j < 4; ++j) s_red[warp][j] = v[j];
    __syncthreads();
    if (DYN && tid == 4) {
        long long t = 0;
#pragma unroll
        for (int w = 0; w < FLAT_THREADS / 32; ++w) t += s_fx[w];
        T.fx_partials[cta] = t;
        __threadfence();
    }
    if (tid < 4) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < FLAT_THREADS / 32; ++w) t += s_red[w][tid];
        T.partials[(size_t)cta * 4 + tid] = t;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(T.ticket, 1u) == (unsigned)grid - 1u);
    __syncthreads();
    if (!s_last) return;

    // ---- the last CTA: every partial in a fixed order (thread-strided, then a fixed tree), the forced-match deltas in image
    //      order, the exchange, the losses; and the workspace is left zeroed
    __threadfence();
    {
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = tid; i < grid; i += FLAT_THREADS)
            for (int j = 0; j < 4; ++j) t[j] += __ldcg(&T.partials[(size_t)i * 4 + j]);
        for (int b = tid; b < T.B; b += FLAT_THREADS) {
            for (int j = 0; j < 3; ++j) {
                t[j + 1] += __ldcg(&T.img_partials[(size_t)b * 3 + j]);
                T.img_partials[(size_t)b * 3 + j] = 0.0;
            }
        }
        for (int j = 0; j < 4; ++j) sm.fin[tid][j] = t[j];
    }
    __syncthreads();
    for (int o = FLAT_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o)
            for (int j = 0; j < 4; ++j) sm.fin[tid][j] += sm.fin[tid + o][j];
        __syncthreads();
    }
    const double fin_flat = sm.fin[0][0], fin_loc = sm.fin[0][1], fin_fix = sm.fin[0][2], fin_cnt = sm.fin[0][3];
    double flat_total = fin_flat;
    if (DYN) {
        // the fixed-point flat sums of all CTAs: integer addition, exact in any order
        __syncthreads();                                                                  // everyone has read sm.fin[0][*]
        long long* fxfin = (long long*)&sm.fin[0][0];
        long long t = 0;
        for (int i = tid; i < grid; i += FLAT_THREADS) t += __ldcg(&T.fx_partials[i]);
        fxfin[tid] = t;
        __syncthreads();
        for (int o = FLAT_THREADS / 2; o > 0; o >>= 1) {
            if (tid < o) fxfin[tid] += fxfin[tid + o];
            __syncthreads();
        }
        flat_total = fin_flat + (double)fxfin[0] * 2.3283064365386963e-10;               // 2^-32
    }
    if (tid == 0) {
        T.out_sums[0] = fin_loc;                                                          // sum loc_losses
        T.out_sums[1] = (double)(1.0f - T.alpha) * flat_total + fin_fix;                  // sum cls_losses
        T.out_sums[2] = fin_cnt;                                                          // num_matches
        *T.ticket = 0u;
        if (DYN) *T.flat_counter = 0ull;
    }
    __syncthreads();
    if (T.use_comm) comm_all_reduce(T.P, T.out_sums, 3);                                  // sums over all image shards (ssd.py:121-122)
    if (tid == 0 && T.out_losses) {
        const double norm = fmax(T.out_sums[2], 1.0);                                     // ssd.py:123
        T.out_losses[0] = (float)(T.out_sums[0] / norm);                                  // ssd.py:131-133
        T.out_losses[1] = (float)(T.out_sums[1] / norm);
    }
}

// ---------------------------------------------------------------------------------------------- host side
int ssdk_match_impl(ssdk_ctx* ctx, const float* anchors, int64_t A, const float* gt_boxes, const int32_t* gt_labels,
                    const int32_t* num_boxes, int B, int Gmax, double pos_thr, double neg_thr, int force,
                    float* out_reg, int32_t* out_cls, int32_t* out_matches, double* out_count = nullptr);
int ssdk_head_loss_core(ssdk_ctx* ctx, const HeadGeom& G, const float* reg_targets, const int32_t* cls_targets,
                        const int32_t* matches, int B, int64_t A, int C, double gamma, double alpha, const double* num_matches,
                        const float* upstream, double* out_sums, const struct ssdk_head_grads* grads, bool with_grad, int phases);
void ssdk_flat_segments(const HeadGeom& G, int B, int C, FlatSegs* S);

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// Targets + losses of a batch.  out_losses may be NULL; flags & SSDK_STEP_ALL_REDUCE: the sums are all-reduced over the
// connected peers inside the kernel (out_sums then holds the global sums on every rank).
int ssdk_train_step_impl(ssdk_ctx* ctx, const HeadGeom& G, const float* anchors, const float* gt_boxes, const int32_t* gt_labels,
                         const int32_t* num_boxes, int B, int64_t A, int C, int Gmax, double pos_thr, double neg_thr, double gamma,
                         double alpha, int flags, double* out_sums, float* out_losses, float* out_reg, int32_t* out_cls,
                         int32_t* out_matches) {
    const size_t NA = (size_t)B * (size_t)A;
    SsdkWsGuard ws_guard(ctx, SSDK_WS_TRAIN);
    SSDK_REQUIRE(out_sums != nullptr, SSDK_ERR_ARG, "targets_and_loss: out_sums is NULL");
    SSDK_REQUIRE(pos_thr >= neg_thr, SSDK_ERR_ARG, "positives_threshold (%g) must be >= negatives_threshold (%g)", pos_thr, neg_thr);
    SSDK_REQUIRE(B >= 0 && A >= 0 && Gmax >= 0 && C > 0, SSDK_ERR_ARG, "targets_and_loss: bad sizes");
    SSDK_REQUIRE(A < (1ll << 31) && B <= 65535, SSDK_ERR_SHAPE, "targets_and_loss: A must be < 2^31 and B <= 65535");
    SSDK_REQUIRE(Gmax <= 4096, SSDK_ERR_SHAPE, "match: at most 4096 ground-truth boxes per image (got %d)", Gmax);
    CommPeers P;
    memset(&P, 0, sizeof(P));
    const bool use_comm = (flags & SSDK_STEP_ALL_REDUCE) != 0;
    if (use_comm) SSDK_REQUIRE(ssdk_comm_peers(ctx, &P), SSDK_ERR_ARG, "SSDK_STEP_ALL_REDUCE: the context is not connected to its peers (ssdk_comm_connect)");
    if (!out_reg) { SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_reg, NA * 16 + 16)); out_reg = (float*)ctx->ws_reg.p; }
    if (!out_cls) { SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_cls, NA * 4 + 16)); out_cls = (int32_t*)ctx->ws_cls.p; }
    if (!out_matches) { SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_matches, NA * 4 + 16)); out_matches = (int32_t*)ctx->ws_matches.p; }

    const bool fusable = NA > 0 && Gmax > 0 && Gmax <= GT_CHUNK && ctx->fused_train_step;
    if (!fusable) {
        // no ground truth at all, more boxes per image than one staging chunk holds, or fusion switched off: the same result
        // from separate launches (matcher, flat pass + matched / ignored anchors, exchange / finalize)
        if (NA == 0) SSDK_CHECK_CUDA(cudaMemsetAsync(out_sums, 0, 3 * sizeof(double), ctx->stream));
        else {
            SSDK_TRY(ssdk_match_impl(ctx, anchors, A, gt_boxes, gt_labels, num_boxes, B, Gmax, pos_thr, neg_thr, 1, out_reg, out_cls,
                                     out_matches));
            SSDK_TRY(ssdk_head_loss_core(ctx, G, out_reg, out_cls, out_matches, B, A, C, gamma, alpha, nullptr, nullptr, out_sums,
                                         nullptr, false, 3));
        }
        if (use_comm && out_losses) return ssdk_comm_loss_finalize(ctx, out_sums, out_losses);
        if (use_comm) return ssdk_comm_all_reduce_sum(ctx, out_sums, 3);
        if (out_losses) return ssdk_loss_finalize(ctx, out_sums, out_losses);
        return SSDK_OK;
    }
    SSDK_REQUIRE(anchors && gt_boxes && gt_labels, SSDK_ERR_ARG, "targets_and_loss: null pointer");
    SSDK_REQUIRE(aligned16(anchors) && aligned16(gt_boxes) && aligned16(out_reg), SSDK_ERR_SHAPE,
                 "targets_and_loss: box arrays must be 16-byte aligned");

    TrainStepArgs T;
    memset(&T, 0, sizeof(T));
    T.G = G;
    T.B = B;
    T.gamma = (float)gamma;
    T.alpha = (float)alpha;
    ssdk_flat_segments(G, B, C, &T.S);
    const long long chunks = T.S.chunk0[T.S.nseg];

    // grid: exactly the co-resident CTAs (persistent); roles and the static split of the chunk list
    static int occ_cache[2] = {0, 0};
    const int variant = gamma == 2.0 ? 0 : 1;
    if (occ_cache[variant] == 0) {
        int occ = 0;
        SSDK_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, variant == 0 ? (const void*)train_step_kernel<0, false> : (const void*)train_step_kernel<1, false>, FLAT_THREADS, 0));
        occ_cache[variant] = occ > 0 ? occ : 1;
    }
    int occ = occ_cache[variant];
    if (ctx->train_ctas_per_sm > 0 && ctx->train_ctas_per_sm < occ) occ = ctx->train_ctas_per_sm;   // leave room for a co-running sub-path
    const int nchunks_img = ceil_div_i(A, MATCH_THREADS);
    int gx = (ctx->num_sms * 6 + B - 1) / B;                               // matcher work items per image (about 6 per SM in total)
    if (gx > nchunks_img) gx = nchunks_img;
    if (gx < 1) gx = 1;
    long long grid = (long long)ctx->num_sms * occ;
    // Role split.  Automatic (the options' defaults) from the two kernels' costs per (image, anchor), measured alone on a B200:
    // matching 13.8 ps + 0.494 ps per ground-truth box (23.7 ps at G = 20, 162 ps at G = 300; latency-bound per CTA: with m
    // matcher CTAs per SM instead of six it takes ~5.7 / m times as long), streaming 0.646 ps per class (4C bytes at 0.96 of the
    // HBM peak, and four streaming CTAs per SM still reach it).  m grows with r = matching / streaming cost; the matcher CTAs are
    // done after tm_busy and then stream for the rest of the kernel: rho = 1 - tm_busy / total is their share of the chunk list
    // relative to a streaming CTA's.  Fitted to the sweep profiles/r2d_split_sweep.json (cfg2: m = 2, rho ~ 0.1 -> 0.1167 ms
    // = 0.95 of the roofline; 100 boxes per image: m = 4; 20 classes: m = 4; stress configuration: m = 4, rho = 0).
    const double t_match = 13.8 + 0.494 * Gmax, t_flat = 0.646 * C;
    bool small_batch = false;
    int m = ctx->match_ctas_per_sm;
    if (m <= 0) {
        const double r = t_match / t_flat;
        m = (int)(6.0 * r / (r + 0.6) + 0.5);
        if (m > occ - 2) m = occ - 2;                                       // two streaming CTAs per SM at least (five matchers: 3-14 % slower)
        if (m < 1) m = 1;
        // Small batches: the matcher's critical path (its chunks one after the other in every matcher CTA, then the forced matches
        // and the final reduction, ~10 us) is longer than the streaming, so a third matcher CTA per SM pays as long as a matcher
        // CTA has fewer than ~19 chunks to walk (cfg2: below 12 images; 13-32 % faster at 2-8 images, profiles/r2u_split_sweep_small.json)
        if (m < 3 && occ >= 5 && (long long)nchunks_img * B < 19ll * ctx->num_sms * m) {
            // (with dynamic chunks the matcher CTAs join the streaming the moment they are done, so a fourth one is free below ~8 images)
            m = (ctx->train_dynamic_chunks && occ >= 6 && (long long)nchunks_img * B < 10ll * ctx->num_sms * m) ? 4 : 3;
            small_batch = true;
        }
    }
    if (m > occ) m = occ;
    long long n_match = (long long)ctx->num_sms * m;
    // every matcher CTA should get the same number of work items: prefer an items-per-image count whose total is a multiple
    // of the matcher CTAs (at 16 images and 296 matcher CTAs: 37 per image = 2 items each, instead of 56 = 3.03)
    for (int g2 = gx; g2 >= (gx + 1) / 2 && g2 >= 1; --g2) {
        if (((long long)g2 * B) % n_match == 0) { gx = g2; break; }
    }
    const long long items2 = (long long)gx * B;
    if (n_match > items2) n_match = items2;
    if (n_match > grid - 1) n_match = grid - 1;
    if (n_match < 1) n_match = 1;
    long long n_flat = grid - n_match;
    if (n_flat > chunks) n_flat = chunks;
    if (n_flat < 1) n_flat = 1;
    grid = n_match + n_flat;
    double rho = ctx->match_flat_share_pct / 100.0;
    if (ctx->match_flat_share_pct < 0) {
        const double tm_busy = t_match * 5.7 / m;
        const double t_all = tm_busy > t_flat + 0.65 * t_match ? tm_busy : t_flat + 0.65 * t_match;
        rho = small_batch ? 0.05 : 1.0 - tm_busy / t_all;
        if (rho < 0.0) rho = 0.0;
        if (rho > 1.0) rho = 1.0;
    }
    // rounds in which every CTA takes a chunk: rho * T / (n_flat + rho * n_match)
    T.rounds_all = (long long)(rho * (double)chunks / ((double)n_flat + rho * (double)n_match));
    T.n_match = (int)n_match;
    T.gx = gx;

    // workspace: [per-CTA partials (fixed size)] [ticket | forced-match deltas | tickets | per-GT keys]: the second part is
    // zero between launches (zeroed when allocated, re-zeroed by the kernel)
    const size_t part_bytes = (size_t)ctx->num_sms * 8 * (4 * sizeof(double) + sizeof(long long));   // double partials, then the fixed-point flat sums
    const size_t zero_bytes = 16 + (size_t)B * 3 * sizeof(double) + (size_t)B * sizeof(int) + 16 + (size_t)B * Gmax * sizeof(unsigned long long);
    SSDK_REQUIRE(grid <= (long long)ctx->num_sms * 8, SSDK_ERR_CUDA, "targets_and_loss: unexpected occupancy %d", occ);
    if (ctx->ws_train.cap < part_bytes + zero_bytes) {
        SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_train, part_bytes + zero_bytes));
        SSDK_CHECK_CUDA(cudaMemsetAsync(ctx->ws_train.p, 0, ctx->ws_train.cap, ctx->stream));
    }
    char* w = (char*)ctx->ws_train.p;
    T.partials = (double*)w;
    T.fx_partials = (long long*)(w + (size_t)ctx->num_sms * 8 * 4 * sizeof(double));
    w += part_bytes;
    T.ticket = (unsigned*)w;
    T.flat_counter = (unsigned long long*)(w + 8);
    T.img_partials = (double*)(w + 16);
    int* tickets = (int*)(w + 16 + (size_t)B * 3 * sizeof(double));
    unsigned long long* best = (unsigned long long*)(((uintptr_t)(tickets + B) + 15) & ~(uintptr_t)15);

    T.M.anchors = (const float4*)anchors; T.M.A = (int)A;
    T.M.gt_boxes = (const float4*)gt_boxes; T.M.gt_labels = gt_labels; T.M.num_boxes = num_boxes; T.M.Gmax = Gmax;
    T.M.pos_thr = (float)pos_thr; T.M.neg_thr = (float)neg_thr; T.M.same_thr = (pos_thr == neg_thr) ? 1 : 0;   // compared as Python floats (:94)
    T.M.gt_best = best; T.M.tickets = tickets; T.M.img_count = nullptr; T.M.out_count = nullptr;
    T.M.matches = out_matches; T.M.reg = (float4*)out_reg; T.M.cls = out_cls; T.M.self_clean = 1;
    T.out_sums = out_sums;
    T.out_losses = out_losses;
    T.P = P;
    T.use_comm = use_comm ? 1 : 0;
    const bool dyn = ctx->train_dynamic_chunks != 0;
    SSDK_KERNEL(ctx, SSDK_K_TRAIN_STEP,
                if (variant == 0 && dyn) train_step_kernel<0, true><<<(int)grid, FLAT_THREADS, 0, ctx->stream>>>(T);
                else if (variant == 0) train_step_kernel<0, false><<<(int)grid, FLAT_THREADS, 0, ctx->stream>>>(T);
                else if (dyn) train_step_kernel<1, true><<<(int)grid, FLAT_THREADS, 0, ctx->stream>>>(T);
                else train_step_kernel<1, false><<<(int)grid, FLAT_THREADS, 0, ctx->stream>>>(T));
    return SSDK_OK;
}
