// Head-layout fusion: the SSD loss (forward, and forward + backward) computed directly on the per-level tower outputs
// [B, n*C, h_l, w_l] / [B, n*4, h_l, w_l], i.e. WITHOUT reshape_and_concatenate (detector/box_predictor.py:67-104), which in
// the reference transposes, reshapes and concatenates every logit into [B,A,C] (one full read + write) right before
// detector/ssd.py:71-133 consumes it.
//
// Why this works without a transpose: with not_ignore weights and one-hot targets (ssd.py:96-109, losses.py:34-50)
//     sum cls_losses = (1-alpha) * SUM over ALL logits of neg(x)                                   <- layout independent
//                      + SUM over matched anchors      [ alpha * pos(x_c*) - (1-alpha) * neg(x_c*) ]   <- ~0.5 % of the anchors
//                      - SUM over ignored anchors, all c [ (1-alpha) * neg(x_c) ]                    <- only when neg_thr < pos_thr
// with neg(x) = sigmoid(x)^gamma * softplus(x) and pos(x) = (1-sigmoid(x))^gamma * softplus(-x).  So
//   1. head_flat_kernel streams every level's class tensor as a flat array (128-bit no-allocate loads, register
//      double-buffering, the same e^3 g(e) fast path as ssd_loss_kernel) and needs neither targets nor geometry; with
//      WITH_GRAD it also writes k * d neg/dx for every element, in place of the same flat index of the gradient tensor;
//   2. head_rows_kernel walks matches [B,A] (4 bytes per anchor), and for the few matched / ignored anchors gathers their
//      logits and box codes through the head geometry: corrections to the sum, smooth-L1, the matched count, and (WITH_GRAD)
//      the positive-class / ignored-row gradient fix-ups and the box gradients;  the CTA that finishes last adds all
//      per-CTA partials of both kernels in a fixed order (deterministic, no float atomics).
// Algorithmic traffic per image: 4AC (+4AC gradients) + 4A (matches) + O(matched) -- the [B,A,C] copy the reference makes
// (another 8AC) is gone, and so are the 16A + 4A bytes of codes / cls_targets the anchor-major kernel streams.
//
// Also here: ssdk_head_concat (reshape_and_concatenate itself, as a tiled transpose) for callers that want the
// reference's tensors, and as the un-fused baseline in bench.py.
#include <stdlib.h>

#include "flat.cuh"

// plain stores: st.global.cs (evict-first) was measured and changes nothing (0.2751 vs 0.2725 ms for the fused step)
#define HEAD_STORE(p, v) (*(p) = (v))

struct HeadGradPtrs {
    float* cls[SSDK_MAX_LEVELS];
    float* box[SSDK_MAX_LEVELS];
};

// ---------------------------------------------------------------------------------------------- 1. flat pass
template <int GAMMA_MODE, bool WITH_GRAD>
__global__ void __launch_bounds__(FLAT_THREADS, WITH_GRAD ? 4 : 6) head_flat_kernel(const FlatSegs S, float gamma, float alpha,
                                                                  const double* __restrict__ norm_count,
                                                                  const float* __restrict__ upstream,
                                                                  double* __restrict__ partials /*[grid]*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long total = S.chunk0[S.nseg];
    float k_neg = 0.0f;
    if (WITH_GRAD) {
        const double norm = fmax(*norm_count, 1.0);                       // ssd.py:123 (global count)
        const float u_cls = upstream ? upstream[1] : 1.0f;
        k_neg = (float)((double)u_cls * (1.0 - (double)alpha) / norm);
    }
    double acc = 0.0;

    int cursor = 0;                                                       // level of the most recently loaded chunk
    auto load = [&](FlatChunk& ck, long long g) {
        if (WITH_GRAD) {
            while (g >= S.chunk0[cursor + 1]) ++cursor;
            if (cursor >= S.n) {                                          // zero-fill segment: nothing to read
                ck.lvl = cursor;
                ck.i4 = (g - S.chunk0[cursor]) * FLAT_CHUNK4 + tid;
                return;
            }
        }
        flat_load(S, ck, g, cursor, tid);
    };
    auto compute = [&](FlatChunk& ck, long long g) {
        if (!WITH_GRAD) {
            acc += (double)flat_value<GAMMA_MODE>(S, ck, g, gamma, tid);
            return;
        }
        if (ck.lvl >= S.n) {
            // box gradients: zero everywhere, head_rows_kernel then scatters the matched anchors' values
            const long long n = S.count[ck.lvl], n4 = n >> 2;
            float4* dst4 = (float4*)S.dst[ck.lvl];
#pragma unroll
            for (int u = 0; u < FLAT_U; ++u) {
                const long long i = ck.i4 + u * FLAT_THREADS;
                if (i < n4) HEAD_STORE(dst4 + i, make_float4(0.f, 0.f, 0.f, 0.f));
            }
            if (g + 1 == S.chunk0[ck.lvl + 1] && tid < (int)(n & 3)) S.dst[ck.lvl][(n & ~3ll) + tid] = 0.0f;
            return;
        }
        const long long n4 = S.count[ck.lvl] >> 2;
        float4* dst4 = (float4*)S.dst[ck.lvl];
        float s = 0.0f;
#pragma unroll
        for (int u = 0; u < FLAT_U; ++u) {
            float4 v = ck.v[u];
            float f0, f1, f2, f3;
            v.x = k_neg * focal_negative_both<GAMMA_MODE>(v.x, gamma, f0);
            v.y = k_neg * focal_negative_both<GAMMA_MODE>(v.y, gamma, f1);
            v.z = k_neg * focal_negative_both<GAMMA_MODE>(v.z, gamma, f2);
            v.w = k_neg * focal_negative_both<GAMMA_MODE>(v.w, gamma, f3);
            s += (f0 + f1) + (f2 + f3);
            const long long i = ck.i4 + u * FLAT_THREADS;
            if (i < n4) HEAD_STORE(dst4 + i, v);
        }
        // the (< 4) floats of a level beyond its last float4, handled with the level's last chunk
        if (g + 1 == S.chunk0[ck.lvl + 1]) {
            const long long n = S.count[ck.lvl];
            const int tail = (int)(n & 3);
            if (tid < tail) {
                const long long e = (n & ~3ll) + tid;
                float f;
                S.dst[ck.lvl][e] = k_neg * focal_negative_both<GAMMA_MODE>(S.src[ck.lvl][e], gamma, f);
                s += f;
            }
        }
        acc += (double)s;
    };

    // WITH_GRAD: register double-buffering, the loads of the next chunk are in flight while the current one is evaluated
    {
        FlatChunk ca, cb;
        long long g = blockIdx.x;
        const long long step = gridDim.x;
        if (!WITH_GRAD) {
            // forward: no register double-buffering -- fewer registers give six resident CTAs per SM instead of four; the
            // extra warps hide the load latency just as well (0.1504 vs 0.1600 ms for the forward sub-path).  The read+write
            // pass below does need the prefetch (0.2728 vs 0.2862 ms).
            acc += flat_sum_range<GAMMA_MODE>(S, g, total, step, gamma, tid);
        } else if (g < total) {
            load(ca, g);
            while (true) {
                const long long gb = g + step;
                const bool has_b = gb < total;
                if (has_b) load(cb, gb);
                compute(ca, g);
                if (!has_b) break;
                g = gb + step;
                const bool has_a = g < total;
                if (has_a) load(ca, g);
                compute(cb, gb);
                if (!has_a) break;
            }
        }
    }

    __shared__ double s_red[FLAT_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < FLAT_THREADS / 32; ++w) t += s_red[w];
        partials[blockIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------------------------- 2. matched / ignored anchors
#define ROWS_THREADS 256
#define ROWS_G 2                                          // 128-bit loads of `matches` in flight per lane
template <int GAMMA_MODE, bool WITH_GRAD>
__global__ void __launch_bounds__(ROWS_THREADS) head_rows_kernel(
    const HeadGeom G, const HeadGradPtrs GR, const float4* __restrict__ reg_t, const int* __restrict__ cls_t,
    const int* __restrict__ matches, int A, long long NA, float gamma, float alpha, const double* __restrict__ norm_count,
    const float* __restrict__ upstream, const double* __restrict__ flat_partials, int n_flat,
    double* __restrict__ partials /*[grid][3]*/, unsigned* __restrict__ ticket, double* __restrict__ out_sums /*[3] or NULL*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = G.C, n = G.per_loc, cf = G.channels_first;
    const float one_minus_alpha = 1.0f - alpha;
    float k_loc = 0.0f, k_pos = 0.0f;
    if (WITH_GRAD) {
        const double norm = fmax(*norm_count, 1.0);
        const float u_loc = upstream ? upstream[0] : 1.0f, u_cls = upstream ? upstream[1] : 1.0f;
        k_loc = (float)((double)u_loc / norm);
        k_pos = (float)((double)u_cls * (double)alpha / norm);
    }
    double acc_loc = 0.0, acc_fix = 0.0, acc_cnt = 0.0;

    // Each warp scans 128 consecutive anchors per iteration (one 128-bit load of `matches` per lane).  Background anchors
    // (-1, the overwhelming majority) cost nothing more.  The others -- matched anchors come in clusters around a box -- are
    // compacted into a per-warp queue and then handled ONE PER LANE, so that their dependent gathers (cls_target -> logit)
    // are in flight together instead of one after the other in the lane that happened to own four of them.
    __shared__ int s_queue[ROWS_THREADS / 32][128 * ROWS_G];
    int* queue = s_queue[warp];
    const unsigned lt_mask = (1u << lane) - 1u;
    const long long ngroups = (NA + 3) >> 2;
    const long long stride = (long long)gridDim.x * ROWS_THREADS * ROWS_G;
    for (long long q0 = (long long)blockIdx.x * ROWS_THREADS * ROWS_G; q0 < ngroups; q0 += stride) {
        // ROWS_G 128-bit loads of `matches` per lane, all issued before any is looked at; group u of the warp covers the 128
        // anchors starting at (q0 + u * ROWS_THREADS + warp * 32) * 4
        int mm[ROWS_G][4];
#pragma unroll
        for (int u = 0; u < ROWS_G; ++u) {
            const long long q = q0 + u * ROWS_THREADS + tid;
            mm[u][0] = mm[u][1] = mm[u][2] = mm[u][3] = -1;
            if (q < ngroups) {
                if ((q << 2) + 3 < NA) {
                    const int4 v = __ldg((const int4*)matches + q);
                    mm[u][0] = v.x; mm[u][1] = v.y; mm[u][2] = v.z; mm[u][3] = v.w;
                } else {
                    for (int j = 0; j < 4; ++j)
                        if ((q << 2) + j < NA) mm[u][j] = __ldg(matches + (q << 2) + j);
                }
            }
        }
        bool any_special = false;
#pragma unroll
        for (int u = 0; u < ROWS_G; ++u) any_special |= (mm[u][0] & mm[u][1] & mm[u][2] & mm[u][3]) != -1;
        if (!__any_sync(0xffffffffu, any_special)) continue;             // warp-uniform: only background anchors
        int total = 0;
#pragma unroll
        for (int u = 0; u < ROWS_G; ++u) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned bal = __ballot_sync(0xffffffffu, mm[u][j] != -1);
                if (mm[u][j] != -1) queue[total + __popc(bal & lt_mask)] = u * (ROWS_THREADS * 4) + lane * 4 + j;
                total += __popc(bal);
            }
        }
        __syncwarp();
        const long long warp_base = (q0 + (tid & ~31)) << 2;             // first anchor of this warp's group 0
        for (int r = lane; r < total; r += 32) {
            const long long i = warp_base + queue[r];
            const int m = __ldg(matches + i);                            // L1 hit
            const int b = (int)(i / A);
            const int a = (int)(i - (long long)b * A);
            const int l = head_level_of(G, a);
            const int rr = a - G.anchor_off[l];
            const int loc = rr / n, k = rr - loc * n;
            const int hw = G.hw[l];
            if (m >= 0) {
                // ---- localisation loss (ssd.py:117, losses.py:4-19) + matched count (ssd.py:121-122)
                const long long e0 = head_elem(cf, b, n * 4, hw, k * 4, loc);
                const long long es = cf ? hw : 1;
                const float* pb = G.box[l] + e0;
                const int tc = __ldg(cls_t + i) - 1;                     // one_hot(cls, C+1)[1:] (ssd.py:96-100)
                const float4 p = make_float4(__ldg(pb), __ldg(pb + es), __ldg(pb + 2 * es), __ldg(pb + 3 * es));
                const float4 t = __ldg(reg_t + i);
                acc_loc += (double)smooth_l1_4(p, t);
                acc_cnt += 1.0;
                if (WITH_GRAD) {
                    float* gb = GR.box[l] + e0;
                    gb[0] = k_loc * smooth_l1_grad(p.x, t.x);
                    gb[es] = k_loc * smooth_l1_grad(p.y, t.y);
                    gb[2 * es] = k_loc * smooth_l1_grad(p.z, t.z);
                    gb[3 * es] = k_loc * smooth_l1_grad(p.w, t.w);
                }
                // ---- the positive class: its logit was summed as a negative by the flat pass
                if (tc >= 0 && tc < C) {
                    const long long e = head_elem(cf, b, n * C, hw, k * C + tc, loc);
                    const float x = __ldg(G.cls[l] + e);
                    acc_fix += (double)(alpha * focal_positive<GAMMA_MODE>(x, gamma)) -
                               (double)(one_minus_alpha * focal_negative<GAMMA_MODE>(x, gamma));
                    if (WITH_GRAD) GR.cls[l][e] = k_pos * focal_positive_grad<GAMMA_MODE>(x, gamma);
                }
            } else {
                // ---- ignored anchor (matches == -2: weight 0, ssd.py:103): every class was summed by the flat pass.  The
                //      C gathers of a lane are independent; neighbouring lanes hold neighbouring locations (coalesced for
                //      channels_first planes)
                const long long e0 = head_elem(cf, b, n * C, hw, k * C, loc);
                const long long es = cf ? hw : 1;
                const float* px = G.cls[l] + e0;
                float sub = 0.0f;
#pragma unroll 4
                for (int c = 0; c < C; ++c) sub += focal_negative<GAMMA_MODE>(__ldg(px + c * es), gamma);
                if (WITH_GRAD) {
                    float* gx = GR.cls[l] + e0;
                    for (int c = 0; c < C; ++c) gx[c * es] = 0.0f;
                }
                acc_fix -= (double)(one_minus_alpha * sub);
            }
        }
        __syncwarp();                                                    // the queue is rewritten by the next iteration
    }

    // ---- CTA reduction (fixed order) -> partials[blockIdx.x]; the last CTA combines everything
    __shared__ double s_red[ROWS_THREADS / 32][3];
    __shared__ double s_fin[ROWS_THREADS][4];
    __shared__ int s_last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_loc += __shfl_xor_sync(0xffffffffu, acc_loc, o);
        acc_fix += __shfl_xor_sync(0xffffffffu, acc_fix, o);
        acc_cnt += __shfl_xor_sync(0xffffffffu, acc_cnt, o);
    }
    if (lane == 0) { s_red[warp][0] = acc_loc; s_red[warp][1] = acc_fix; s_red[warp][2] = acc_cnt; }
    __syncthreads();
    if (out_sums == nullptr) return;
    if (tid < 3) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < ROWS_THREADS / 32; ++w) t += s_red[w][tid];
        partials[(size_t)blockIdx.x * 3 + tid] = t;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        // every thread owns the partials tid, tid + 256, ... of both kernels: a fixed assignment and a fixed tree below, so the
        // result does not depend on which CTA happens to be last
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = tid; i < (int)gridDim.x; i += ROWS_THREADS) {
            const double p0 = __ldcg(&partials[(size_t)i * 3 + 0]), p1 = __ldcg(&partials[(size_t)i * 3 + 1]),
                         p2 = __ldcg(&partials[(size_t)i * 3 + 2]);
            t[0] += p0; t[1] += p1; t[2] += p2;
        }
        for (int i = tid; i < n_flat; i += ROWS_THREADS) t[3] += __ldcg(&flat_partials[i]);
        for (int j = 0; j < 4; ++j) s_fin[tid][j] = t[j];
    }
    __syncthreads();
    for (int o = ROWS_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o)
            for (int j = 0; j < 4; ++j) s_fin[tid][j] += s_fin[tid + o][j];
        __syncthreads();
    }
    if (tid == 0) {
        out_sums[0] = s_fin[0][0];                                                   // sum loc_losses
        out_sums[1] = (double)one_minus_alpha * s_fin[0][3] + s_fin[0][1];           // sum cls_losses
        out_sums[2] = s_fin[0][2];                                                   // num_matches
        *ticket = 0u;
    }
}

// ---------------------------------------------------------------------------------------------- reshape_and_concatenate
// channels_first level [B, CH, hw] -> out[b, (off + loc) * CH/n ... ]: for one image the level is a [CH, hw] matrix whose
// transpose [hw, CH] is the level's contiguous block of the [A, C] (or [A, 4]) output.  32x32 tiles through shared memory.
__global__ void __launch_bounds__(256) head_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int CH, int hw,
                                                             long long out_image_stride, long long out_level_off) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const float* src = in + (size_t)b * CH * hw;
    float* dst = out + (size_t)b * out_image_stride + out_level_off;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;             // 32 x 8
    const int loc0 = blockIdx.x * 32, ch0 = blockIdx.y * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int ch = ch0 + ty + j, loc = loc0 + tx;
        if (ch < CH && loc < hw) tile[ty + j][tx] = src[(size_t)ch * hw + loc];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int loc = loc0 + ty + j, ch = ch0 + tx;
        if (ch < CH && loc < hw) dst[(size_t)loc * CH + ch] = tile[tx][ty + j];
    }
}

// channels_last level [B, hw*CH] -> the level's block of every image: a strided copy
__global__ void __launch_bounds__(256) head_copy_kernel(const float* __restrict__ in, float* __restrict__ out, long long per_image,
                                                        long long out_image_stride, long long out_level_off) {
    const int b = blockIdx.y;
    const float* src = in + (size_t)b * per_image;
    float* dst = out + (size_t)b * out_image_stride + out_level_off;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < per_image; i += (long long)gridDim.x * 256) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------- host side
static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

int ssdk_head_geom(const ssdk_head* head, int B, int64_t A, int C, bool need_cls, bool need_box, HeadGeom* out) {
    SSDK_REQUIRE(head != nullptr, SSDK_ERR_ARG, "ssdk_head: null head descriptor");
    SSDK_REQUIRE(head->num_levels >= 1 && head->num_levels <= SSDK_MAX_LEVELS, SSDK_ERR_ARG, "ssdk_head: num_levels %d not in [1,%d]",
                 head->num_levels, SSDK_MAX_LEVELS);
    SSDK_REQUIRE(head->anchors_per_location >= 1, SSDK_ERR_ARG, "ssdk_head: anchors_per_location %d", head->anchors_per_location);
    SSDK_REQUIRE(head->data_format == SSDK_CHANNELS_FIRST || head->data_format == SSDK_CHANNELS_LAST, SSDK_ERR_ARG,
                 "ssdk_head: data_format %d", head->data_format);
    SSDK_REQUIRE(B >= 0 && A >= 0 && C > 0, SSDK_ERR_ARG, "ssdk_head: bad sizes (B=%d A=%lld C=%d)", B, (long long)A, C);
    SSDK_REQUIRE(A < (1ll << 31), SSDK_ERR_SHAPE, "ssdk_head: A must be < 2^31");
    HeadGeom g;
    g.num_levels = head->num_levels;
    g.per_loc = head->anchors_per_location;
    g.channels_first = head->data_format == SSDK_CHANNELS_FIRST;
    g.C = C;
    long long off = 0;
    for (int l = 0; l < SSDK_MAX_LEVELS; ++l) {
        g.anchor_off[l] = (int)off;
        g.hw[l] = 0; g.cls[l] = nullptr; g.box[l] = nullptr;
        if (l >= head->num_levels) continue;
        SSDK_REQUIRE(head->height[l] >= 0 && head->width[l] >= 0, SSDK_ERR_ARG, "ssdk_head: level %d has negative size", l);
        const long long hw = (long long)head->height[l] * head->width[l];
        SSDK_REQUIRE(hw * g.per_loc * (long long)(C > 4 ? C : 4) < (1ll << 31), SSDK_ERR_SHAPE, "ssdk_head: level %d too large", l);
        g.hw[l] = (int)hw;
        g.cls[l] = head->class_predictions[l];
        g.box[l] = head->encoded_boxes[l];
        if (hw > 0 && B > 0) {
            SSDK_REQUIRE(!need_cls || g.cls[l], SSDK_ERR_ARG, "ssdk_head: class_predictions[%d] is NULL", l);
            SSDK_REQUIRE(!need_box || g.box[l], SSDK_ERR_ARG, "ssdk_head: encoded_boxes[%d] is NULL", l);
            SSDK_REQUIRE(aligned16(g.cls[l]) && aligned16(g.box[l]), SSDK_ERR_SHAPE, "ssdk_head: level %d tensors must be 16-byte aligned", l);
        }
        off += hw * g.per_loc;
    }
    for (int l = head->num_levels; l <= SSDK_MAX_LEVELS; ++l) g.anchor_off[l] = (int)off;
    SSDK_REQUIRE(off == A, SSDK_ERR_SHAPE, "ssdk_head: levels hold %lld anchors, expected A = %lld", off, (long long)A);
    *out = g;
    return SSDK_OK;
}

// The class tensors of all levels as the chunk list of the flat pass (forward: segments [0, num_levels)).
void ssdk_flat_segments(const HeadGeom& G, int B, int C, FlatSegs* out) {
    FlatSegs S;
    const int nl = G.num_levels;
    S.n = nl;
    S.nseg = nl;
    long long chunks = 0;
    for (int l = 0; l < SSDK_MAX_LEVELS; ++l) S.src[l] = nullptr;
    for (int sg = 0; sg < 2 * SSDK_MAX_LEVELS; ++sg) { S.dst[sg] = nullptr; S.count[sg] = 0; }
    for (int sg = 0; sg < nl; ++sg) {
        S.chunk0[sg] = chunks;
        S.count[sg] = (long long)B * G.per_loc * C * G.hw[sg];
        chunks += (S.count[sg] + 4 * FLAT_CHUNK4 - 1) / (4 * FLAT_CHUNK4);
        S.src[sg] = G.cls[sg];
    }
    for (int sg = nl; sg <= 2 * SSDK_MAX_LEVELS; ++sg) S.chunk0[sg] = chunks;
    *out = S;
}

// phases: 1 = the flat pass (needs no targets), 2 = the matched / ignored anchors + final reduction, 3 = both
int ssdk_head_loss_core(ssdk_ctx* ctx, const HeadGeom& G, const float* reg_targets, const int32_t* cls_targets,
                        const int32_t* matches, int B, int64_t A, int C, double gamma, double alpha, const double* num_matches,
                        const float* upstream, double* out_sums, const ssdk_head_grads* grads, bool with_grad, int phases) {
    const long long NA = (long long)B * A;
    SsdkWsGuard ws_guard(ctx, SSDK_WS_LOSS);
    if (NA == 0) {
        if (out_sums) SSDK_CHECK_CUDA(cudaMemsetAsync(out_sums, 0, 3 * sizeof(double), ctx->stream));
        return SSDK_OK;
    }
    SSDK_REQUIRE(reg_targets && cls_targets && matches, SSDK_ERR_ARG, "ssdk_head_ssd_loss: null target pointer");
    SSDK_REQUIRE(aligned16(reg_targets), SSDK_ERR_SHAPE, "ssdk_head_ssd_loss: reg_targets must be 16-byte aligned");
    SSDK_REQUIRE(!with_grad || (grads && num_matches), SSDK_ERR_ARG, "ssdk_head_ssd_loss_forward_backward: null grads / num_matches");
    SSDK_REQUIRE(with_grad || out_sums, SSDK_ERR_ARG, "ssdk_head_ssd_loss: out_sums is NULL");

    SSDK_REQUIRE(aligned16(matches), SSDK_ERR_SHAPE, "ssdk_head_ssd_loss: matches must be 16-byte aligned");
    FlatSegs S;
    HeadGradPtrs GR;
    const int nl = G.num_levels;
    ssdk_flat_segments(G, B, C, &S);
    for (int l = 0; l < SSDK_MAX_LEVELS; ++l) { GR.cls[l] = nullptr; GR.box[l] = nullptr; }
    if (with_grad) {
        // zero-fill segments [n, 2n) for the box gradients, appended to the chunk list
        long long chunks_g = S.chunk0[nl];
        S.nseg = 2 * nl;
        for (int sg = 0; sg < S.nseg; ++sg) {
            const int l = sg < nl ? sg : sg - nl;
            if (sg >= nl) {
                S.chunk0[sg] = chunks_g;
                S.count[sg] = (long long)B * G.per_loc * 4 * G.hw[l];
                chunks_g += (S.count[sg] + 4 * FLAT_CHUNK4 - 1) / (4 * FLAT_CHUNK4);
            }
            if (S.count[sg] > 0) {
                float* gp = sg < nl ? grads->class_predictions[l] : grads->encoded_boxes[l];
                SSDK_REQUIRE(gp != nullptr, SSDK_ERR_ARG, "ssdk_head_ssd_loss_forward_backward: grads of level %d are NULL", l);
                SSDK_REQUIRE(aligned16(gp), SSDK_ERR_SHAPE, "ssdk_head_ssd_loss_forward_backward: grads of level %d must be 16-byte aligned", l);
                S.dst[sg] = gp;
                if (sg < nl) GR.cls[l] = gp; else GR.box[l] = gp;        // box gradients: zero-filled by the flat kernel,
            }                                                            // matched anchors scattered by head_rows_kernel
        }
        for (int sg = S.nseg; sg <= 2 * SSDK_MAX_LEVELS; ++sg) S.chunk0[sg] = chunks_g;
    }
    const long long chunks = S.chunk0[S.nseg];

    // persistent-style grid: exactly the number of co-resident CTAs (a larger grid would add a partial second wave)
    static int occ_cache[4] = {0, 0, 0, 0};
    const int variant = (gamma == 2.0 ? 0 : 2) + (with_grad ? 1 : 0);
    if (occ_cache[variant] == 0) {
        const void* fn = variant == 0 ? (const void*)head_flat_kernel<0, false> : variant == 1 ? (const void*)head_flat_kernel<0, true>
                       : variant == 2 ? (const void*)head_flat_kernel<1, false> : (const void*)head_flat_kernel<1, true>;
        int occ = 0;
        SSDK_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, FLAT_THREADS, 0));
        occ_cache[variant] = occ > 0 ? occ : 1;
    }
    int per_sm = occ_cache[variant];
    if (ctx->tune_head_ctas) per_sm = ctx->tune_head_ctas;                       // SSDK_HEAD_CTAS (validated at context creation)
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long grid_flat = (long long)ctx->num_sms * per_sm;
    if (grid_flat > chunks) grid_flat = chunks;
    if (grid_flat < 1) grid_flat = 1;
    long long grid_rows = (NA + 4 * ROWS_THREADS * ROWS_G - 1) / (4 * ROWS_THREADS * ROWS_G);
    if (grid_rows > (long long)ctx->num_sms * 8) grid_rows = (long long)ctx->num_sms * 8;

    const size_t ws_bytes = 16 + (size_t)ctx->num_sms * 8 * 4 * sizeof(double);
    if (ctx->ws_head.cap < ws_bytes) {
        SSDK_TRY(ssdk_ensure(ctx, &ctx->ws_head, ws_bytes));
        SSDK_CHECK_CUDA(cudaMemsetAsync(ctx->ws_head.p, 0, 16, ctx->stream));
    }
    unsigned* ticket = (unsigned*)ctx->ws_head.p;
    double* flat_partials = (double*)((char*)ctx->ws_head.p + 16);
    double* rows_partials = flat_partials + (size_t)ctx->num_sms * 8;

    const float gf = (float)gamma, af = (float)alpha;
    const bool g2 = (gamma == 2.0);
#define SSDK_LAUNCH_HEAD(GM, WG)                                                                                              \
    do {                                                                                                                      \
        if (phases & 1)                                                                                                       \
            SSDK_KERNEL(ctx, SSDK_K_HEAD_FLAT,                                                                                \
                        head_flat_kernel<GM, WG><<<(int)grid_flat, FLAT_THREADS, 0, ctx->stream>>>(S, gf, af, num_matches,    \
                                                                                                  upstream, flat_partials)); \
        if (phases & 2)                                                                                                       \
            SSDK_KERNEL(ctx, SSDK_K_HEAD_ROWS,                                                                                \
                        head_rows_kernel<GM, WG><<<(int)grid_rows, ROWS_THREADS, 0, ctx->stream>>>(                           \
                            G, GR, (const float4*)reg_targets, cls_targets, matches, (int)A, NA, gf, af, num_matches, upstream, \
                            flat_partials, (int)grid_flat, rows_partials, ticket, out_sums));                                 \
    } while (0)
    if (g2 && !with_grad) SSDK_LAUNCH_HEAD(0, false);
    else if (g2) SSDK_LAUNCH_HEAD(0, true);
    else if (!with_grad) SSDK_LAUNCH_HEAD(1, false);
    else SSDK_LAUNCH_HEAD(1, true);
#undef SSDK_LAUNCH_HEAD
    return SSDK_OK;
}

static int head_loss_impl(ssdk_ctx* ctx, const ssdk_head* head, const float* reg_targets, const int32_t* cls_targets,
                          const int32_t* matches, int B, int64_t A, int C, double gamma, double alpha, const double* num_matches,
                          const float* upstream, double* out_sums, const ssdk_head_grads* grads, bool with_grad) {
    SSDK_ENTER(ctx);
    HeadGeom G;
    SSDK_TRY(ssdk_head_geom(head, B, A, C, true, true, &G));
    return ssdk_head_loss_core(ctx, G, reg_targets, cls_targets, matches, B, A, C, gamma, alpha, num_matches, upstream, out_sums,
                               grads, with_grad, 3);
}

// HeadGeom of the anchor-major tensors [B,A,C] / [B,A,4]: one channels_last "level" with one anchor per location
HeadGeom ssdk_flat_geom(const float* logits, const float* codes, int64_t A, int C) {
    HeadGeom g;
    g.num_levels = 1; g.per_loc = 1; g.channels_first = 0; g.C = C;
    for (int l = 0; l < SSDK_MAX_LEVELS; ++l) { g.anchor_off[l] = l ? (int)A : 0; g.hw[l] = 0; g.cls[l] = nullptr; g.box[l] = nullptr; }
    g.anchor_off[SSDK_MAX_LEVELS] = (int)A;
    g.hw[0] = (int)A; g.cls[0] = logits; g.box[0] = codes;
    return g;
}

int ssdk_train_step_impl(ssdk_ctx* ctx, const HeadGeom& G, const float* anchors, const float* gt_boxes, const int32_t* gt_labels,
                         const int32_t* num_boxes, int B, int64_t A, int C, int Gmax, double pos_thr, double neg_thr, double gamma,
                         double alpha, int flags, double* out_sums, float* out_losses, float* out_reg, int32_t* out_cls,
                         int32_t* out_matches);

extern "C" {

int ssdk_head_ssd_targets_and_loss(ssdk_ctx* ctx, const ssdk_head* head, const float* anchors, const float* gt_boxes,
                                   const int32_t* gt_labels, const int32_t* num_boxes, int B, int64_t A, int C, int Gmax,
                                   double positives_threshold, double negatives_threshold, double gamma, double alpha,
                                   double* out_sums, float* out_reg, int32_t* out_cls, int32_t* out_matches) {
    SSDK_ENTER(ctx);
    HeadGeom G;
    SSDK_TRY(ssdk_head_geom(head, B, A, C, true, true, &G));
    return ssdk_train_step_impl(ctx, G, anchors, gt_boxes, gt_labels, num_boxes, B, A, C, Gmax, positives_threshold, negatives_threshold,
                                gamma, alpha, 0, out_sums, nullptr, out_reg, out_cls, out_matches);
}

int ssdk_head_ssd_loss_step(ssdk_ctx* ctx, const ssdk_head* head, const float* anchors, const float* gt_boxes,
                            const int32_t* gt_labels, const int32_t* num_boxes, int B, int64_t A, int C, int Gmax,
                            double positives_threshold, double negatives_threshold, double gamma, double alpha, int flags,
                            double* out_sums, float* out_losses, float* out_reg, int32_t* out_cls, int32_t* out_matches) {
    SSDK_ENTER(ctx);
    HeadGeom G;
    SSDK_TRY(ssdk_head_geom(head, B, A, C, true, true, &G));
    return ssdk_train_step_impl(ctx, G, anchors, gt_boxes, gt_labels, num_boxes, B, A, C, Gmax, positives_threshold, negatives_threshold,
                                gamma, alpha, flags, out_sums, out_losses, out_reg, out_cls, out_matches);
}

int ssdk_head_ssd_loss(ssdk_ctx* ctx, const ssdk_head* head, const float* reg_targets, const int32_t* cls_targets,
                       const int32_t* matches, int B, int64_t A, int C, double gamma, double alpha, double* out_sums) {
    return head_loss_impl(ctx, head, reg_targets, cls_targets, matches, B, A, C, gamma, alpha, nullptr, nullptr, out_sums, nullptr,
                          false);
}

int ssdk_head_ssd_loss_forward_backward(ssdk_ctx* ctx, const ssdk_head* head, const float* reg_targets,
                                        const int32_t* cls_targets, const int32_t* matches, int B, int64_t A, int C,
                                        double gamma, double alpha, const double* num_matches, const float* upstream,
                                        double* out_sums, const ssdk_head_grads* grads) {
    return head_loss_impl(ctx, head, reg_targets, cls_targets, matches, B, A, C, gamma, alpha, num_matches, upstream, out_sums,
                          grads, true);
}

int ssdk_head_concat(ssdk_ctx* ctx, const ssdk_head* head, int B, int C, float* out_encoded_boxes, float* out_class_predictions) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(head != nullptr, SSDK_ERR_ARG, "ssdk_head_concat: null head descriptor");
    long long A = 0;
    for (int l = 0; l < head->num_levels && l < SSDK_MAX_LEVELS; ++l)
        A += (long long)head->height[l] * head->width[l] * head->anchors_per_location;
    HeadGeom G;
    SSDK_TRY(ssdk_head_geom(head, B, A, C, out_class_predictions != nullptr, out_encoded_boxes != nullptr, &G));
    if (B == 0 || A == 0) return SSDK_OK;
    SSDK_REQUIRE(B <= 65535, SSDK_ERR_SHAPE, "ssdk_head_concat: batch %d > 65535", B);
    for (int which = 0; which < 2; ++which) {
        float* out = which ? out_class_predictions : out_encoded_boxes;
        if (!out) continue;
        const int D = which ? C : 4;                                      // values per anchor
        for (int l = 0; l < G.num_levels; ++l) {
            const int hw = G.hw[l];
            if (hw == 0) continue;
            const float* in = which ? G.cls[l] : G.box[l];
            const int CH = G.per_loc * D;
            const long long level_off = (long long)G.anchor_off[l] * D;
            if (G.channels_first) {
                const dim3 grid(ceil_div_i(hw, 32), ceil_div_i(CH, 32), B);
                SSDK_KERNEL(ctx, SSDK_K_HEAD_CONCAT,
                            head_transpose_kernel<<<grid, 256, 0, ctx->stream>>>(in, out, CH, hw, A * D, level_off));
            } else {
                const long long per_image = (long long)hw * CH;
                long long gx = (per_image + 256 * 8 - 1) / (256 * 8);
                if (gx > 4096) gx = 4096;
                SSDK_KERNEL(ctx, SSDK_K_HEAD_CONCAT,
                            head_copy_kernel<<<dim3((unsigned)gx, B), 256, 0, ctx->stream>>>(in, out, per_image, A * D, level_off));
            }
        }
    }
    return SSDK_OK;
}

}  // extern "C"
