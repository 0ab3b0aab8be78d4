// Shared host/device helpers for libssdk (sm_100a).  See include/ssdk.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/ssdk.h"

// ----------------------------------------------------------------------------------------------
// context + error plumbing
// ----------------------------------------------------------------------------------------------
struct ssdk_buf {
    void* p = nullptr;
    size_t cap = 0;
};

// kernel ids for the optional per-kernel event timing (ssdk_ctx_set_profiling)
enum ssdk_kernel_id {
    SSDK_K_ANCHORS = 0, SSDK_K_MATCH, SSDK_K_FORCE_MATCH, SSDK_K_LOSS, SSDK_K_LOSS_REDUCE, SSDK_K_FILTER,
    SSDK_K_SORT, SSDK_K_NMS, SSDK_K_PACK, SSDK_K_OTHER, SSDK_K_LOSS_BACKWARD, SSDK_K_HEAD_FLAT, SSDK_K_HEAD_ROWS,
    SSDK_K_HEAD_CONCAT, SSDK_K_COMM, SSDK_K_TRAIN_STEP, SSDK_K_NMS_ROUNDS, SSDK_K_COUNT
};
#define SSDK_PROFILE_EVENTS 2048

struct ssdk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    int64_t launches = 0;
    // private workspace (grows on demand)
    ssdk_buf ws_gtbest;      // matcher: per-(image, gt) packed (iou, anchor) keys
    ssdk_buf ws_partials;    // loss: per-CTA partial sums
    ssdk_buf ws_reg, ws_cls, ws_matches;   // targets when the caller does not want them
    ssdk_buf ws_cand;        // postprocess: candidate keys  [B][cap] u64
    ssdk_buf ws_counts;      // postprocess: per-image counters + per-(image,class) segment tables
    ssdk_buf ws_seg;         // postprocess: per-(image,class) kept boxes/scores/anchors
    ssdk_buf ws_stage[8];    // *_host entry points: device staging
    ssdk_buf ws_head;        // head-layout loss: ticket + per-CTA partials of the flat and the rows kernel
    ssdk_buf ws_summ;        // level summaries: per-(image, level) scratch
    // profiling: pairs of events around kernels, drained by ssdk_ctx_profile_read
    int profiling = 0;
    int prof_n = 0;
    cudaEvent_t* prof_ev = nullptr;          // [2 * SSDK_PROFILE_EVENTS]
    int prof_id[SSDK_PROFILE_EVENTS];
    double prof_ms[SSDK_K_COUNT] = {0};
    long long prof_calls[SSDK_K_COUNT] = {0};
    ssdk_buf ws_train;       // fused training step: per-CTA partials | ticket, forced-match deltas, per-image tickets, per-GT keys (zero between launches)
    int fused_train_step = 1;      // SSDK_OPT_FUSED_TRAIN_STEP
    int match_ctas_per_sm = 0;     // SSDK_OPT_MATCH_CTAS_PER_SM (0 = automatic)
    int match_flat_share_pct = -1; // SSDK_OPT_MATCH_FLAT_SHARE_PCT (-1 = automatic)
    int train_dynamic_chunks = 1;  // SSDK_OPT_TRAIN_DYNAMIC_CHUNKS: the flat pass of the fused training step hands out its chunks dynamically
    int train_ctas_per_sm = 0;     // SSDK_OPT_TRAIN_CTAS_PER_SM (0 = as many as fit)
    int use_pdl = 1;               // SSDK_OPT_PROGRAMMATIC_LAUNCH: chain the post-processing kernels with programmatic dependent launch
    // tuning knobs, read from the environment ONCE (ssdk_ctx_create) and validated there; 0 = automatic
    int tune_head_ctas = 0, tune_loss_rpw = 0, tune_loss_stages = 0, tune_loss_ctas = 0;
    unsigned long long* round_times = nullptr;   // inside ws_counts: phase timestamps of the last dense-segment rounds (ssdk_ctx_round_times)
    int* dev_err = nullptr;  // device int: sticky asynchronous error raised by a kernel (ssdk_ctx_async_error)
    void* comm = nullptr;    // peer-memory communicator (comm.cu), NULL until ssdk_comm_local_handle
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // workspace groups (SsdkWsGuard): event recorded after the last use, the stream it was recorded on
    cudaEvent_t ws_event[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t ws_stream[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ws_valid[8] = {false, false, false, false, false, false, false, false};
};

void ssdk_set_error(const char* fmt, ...);
int ssdk_set_max_smem(ssdk_ctx* ctx, const void* func, int bytes);
int ssdk_ensure(ssdk_ctx* ctx, ssdk_buf* b, size_t bytes);

#define SSDK_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ssdk_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,            \
                           cudaGetErrorString(_e));                                        \
            return SSDK_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define SSDK_CHECK_LAUNCH(ctx)                                                             \
    do {                                                                                   \
        (ctx)->launches++;                                                                 \
        SSDK_CHECK_CUDA(cudaGetLastError());                                               \
    } while (0)

// Bracket one kernel launch with events when profiling is on (no cost otherwise).
int ssdk_prof_begin(ssdk_ctx* ctx, int id);
void ssdk_prof_end(ssdk_ctx* ctx, int slot);
#define SSDK_KERNEL(ctx, id, ...)                                                          \
    do {                                                                                   \
        const int _slot = (ctx)->profiling ? ssdk_prof_begin((ctx), (id)) : -1;            \
        __VA_ARGS__;                                                                       \
        if (_slot >= 0) ssdk_prof_end((ctx), _slot);                                       \
        SSDK_CHECK_LAUNCH(ctx);                                                            \
    } while (0)

#define SSDK_REQUIRE(cond, code, ...)                                                      \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            ssdk_set_error(__VA_ARGS__);                                                   \
            return (code);                                                                 \
        }                                                                                  \
    } while (0)

#define SSDK_TRY(expr)                                                                     \
    do {                                                                                   \
        int _s = (expr);                                                                   \
        if (_s != SSDK_OK) return _s;                                                      \
    } while (0)

// Entry guard of every API function: makes the context's device current for the duration of the call and RESTORES the
// caller's device afterwards (a process that drives several GPUs must not find its current device switched by a library call).
struct SsdkDeviceGuard {
    int prev = -1;
    int status = SSDK_OK;
    explicit SsdkDeviceGuard(ssdk_ctx* ctx) {
        if (!ctx) {
            ssdk_set_error("null context");
            status = SSDK_ERR_ARG;
            return;
        }
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != ctx->device) {
            const cudaError_t e = cudaSetDevice(ctx->device);
            if (e != cudaSuccess) {
                ssdk_set_error("cudaSetDevice(%d) failed: %s", ctx->device, cudaGetErrorString(e));
                status = SSDK_ERR_CUDA;
                prev = -1;
            }
        } else {
            prev = -1;                                      // nothing to restore
        }
    }
    ~SsdkDeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    SsdkDeviceGuard(const SsdkDeviceGuard&) = delete;
    SsdkDeviceGuard& operator=(const SsdkDeviceGuard&) = delete;
};
#define SSDK_ENTER(ctx)                  \
    SsdkDeviceGuard _ssdk_dev_guard(ctx); \
    SSDK_TRY(_ssdk_dev_guard.status)

// Workspace groups: buffers inside the context that one entry point writes and a later one may reuse.  Work is enqueued on
// whatever stream the caller set, so two calls on DIFFERENT streams could touch the same buffer concurrently.  Every entry
// point brackets its use of a group with this guard: on entry the call's stream waits for the group's last use if that
// was on another stream; on exit the use is recorded.  While a stream is being captured into a CUDA graph no events are
// exchanged (a captured event cannot be waited on from outside the capture): inside a graph the caller's fork / join
// (e.g. graph.concurrent) defines the order, and calls that share a group must not be put on parallel branches.
enum ssdk_ws_group_id { SSDK_WS_TRAIN = 0, SSDK_WS_MATCH, SSDK_WS_LOSS, SSDK_WS_POST, SSDK_WS_STAGE, SSDK_WS_SUMM, SSDK_WS_GROUPS };
struct SsdkWsGuard {
    ssdk_ctx* ctx;
    int group;
    bool capturing = false;
    SsdkWsGuard(ssdk_ctx* c, int g);
    ~SsdkWsGuard();
    SsdkWsGuard(const SsdkWsGuard&) = delete;
    SsdkWsGuard& operator=(const SsdkWsGuard&) = delete;
};

// ----------------------------------------------------------------------------------------------
// Parity-critical float32 arithmetic.  The reference executes one TF kernel per Python-level op,
// so every operation rounds separately: explicit _rn intrinsics keep nvcc from contracting
// a*b+c into an FMA and force IEEE division / sqrt.
// ----------------------------------------------------------------------------------------------
#define SSDK_EPS 1e-8f  // detector/constants.py:12

__device__ __forceinline__ float f_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float f_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float f_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float f_div(float a, float b) { return __fdiv_rn(a, b); }

// area: box_utils.py:53-61
__device__ __forceinline__ float box_area(const float4 b) {
    return f_mul(f_sub(b.z, b.x), f_sub(b.w, b.y));
}
// intersection: box_utils.py:30-50
__device__ __forceinline__ float box_intersection(const float4 a, const float4 b) {
    const float ih = fmaxf(0.0f, f_sub(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float iw = fmaxf(0.0f, f_sub(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    return f_mul(ih, iw);
}
// iou: box_utils.py:14-27  ((a1 + a2) - inter, + eps, divide, clip)
__device__ __forceinline__ float box_iou_areas(const float4 a, float area_a, const float4 b, float area_b) {
    const float inter = box_intersection(a, b);
    const float uni = f_sub(f_add(area_a, area_b), inter);
    const float q = f_div(inter, f_add(uni, SSDK_EPS));
    return fminf(fmaxf(q, 0.0f), 1.0f);
}

// encode: box_utils.py:80-111 (to_center_coordinates :64-77); scale factors constants.py:15
__device__ __forceinline__ float4 box_encode(const float4 box, const float4 anchor) {
    float ha = f_sub(anchor.z, anchor.x), wa = f_sub(anchor.w, anchor.y);
    const float cya = f_add(anchor.x, f_mul(0.5f, ha)), cxa = f_add(anchor.y, f_mul(0.5f, wa));
    float h = f_sub(box.z, box.x), w = f_sub(box.w, box.y);
    const float cy = f_add(box.x, f_mul(0.5f, h)), cx = f_add(box.y, f_mul(0.5f, w));
    ha = f_add(ha, SSDK_EPS); wa = f_add(wa, SSDK_EPS);
    h = f_add(h, SSDK_EPS); w = f_add(w, SSDK_EPS);
    float4 t;
    t.x = f_mul(f_div(f_sub(cy, cya), ha), 10.0f);
    t.y = f_mul(f_div(f_sub(cx, cxa), wa), 10.0f);
    t.z = f_mul(logf(f_div(h, ha)), 5.0f);
    t.w = f_mul(logf(f_div(w, wa)), 5.0f);
    return t;
}

// decode: box_utils.py:114-142
__device__ __forceinline__ float4 box_decode(const float4 code, const float4 anchor) {
    const float ha = f_sub(anchor.z, anchor.x), wa = f_sub(anchor.w, anchor.y);
    const float cya = f_add(anchor.x, f_mul(0.5f, ha)), cxa = f_add(anchor.y, f_mul(0.5f, wa));
    const float ty = f_div(code.x, 10.0f), tx = f_div(code.y, 10.0f);
    const float th = f_div(code.z, 5.0f), tw = f_div(code.w, 5.0f);
    const float h = f_mul(expf(th), ha), w = f_mul(expf(tw), wa);
    const float cy = f_add(f_mul(ty, ha), cya), cx = f_add(f_mul(tx, wa), cxa);
    const float hh = f_mul(0.5f, h), hw = f_mul(0.5f, w);
    return make_float4(f_sub(cy, hh), f_sub(cx, hw), f_add(cy, hh), f_add(cx, hw));
}

__device__ __forceinline__ float4 box_clip01(float4 b) {
    b.x = fminf(fmaxf(b.x, 0.0f), 1.0f); b.y = fminf(fmaxf(b.y, 0.0f), 1.0f);
    b.z = fminf(fmaxf(b.z, 0.0f), 1.0f); b.w = fminf(fmaxf(b.w, 0.0f), 1.0f);
    return b;
}

// TF 1.12 NonMaxSuppressionV3 suppression test (external to the reference; see oracle/nms.py).
__device__ __forceinline__ bool nms_iou_greater(const float4 bi, const float4 bj, float thr) {
    const float ymin_i = fminf(bi.x, bi.z), xmin_i = fminf(bi.y, bi.w);
    const float ymax_i = fmaxf(bi.x, bi.z), xmax_i = fmaxf(bi.y, bi.w);
    const float ymin_j = fminf(bj.x, bj.z), xmin_j = fminf(bj.y, bj.w);
    const float ymax_j = fmaxf(bj.x, bj.z), xmax_j = fmaxf(bj.y, bj.w);
    const float area_i = f_mul(f_sub(ymax_i, ymin_i), f_sub(xmax_i, xmin_i));
    const float area_j = f_mul(f_sub(ymax_j, ymin_j), f_sub(xmax_j, xmin_j));
    if (area_i <= 0.0f || area_j <= 0.0f) return false;
    const float ih = fmaxf(f_sub(fminf(ymax_i, ymax_j), fmaxf(ymin_i, ymin_j)), 0.0f);
    const float iw = fmaxf(f_sub(fminf(xmax_i, xmax_j), fmaxf(xmin_i, xmin_j)), 0.0f);
    const float inter = f_mul(ih, iw);
    const float iou = f_div(inter, f_sub(f_add(area_i, area_j), inter));
    return iou > thr;
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may start while its
// predecessor in the stream is still running; it must execute pdl_wait() before it touches anything the predecessor wrote (the
// wait returns once the predecessor grid has completed and its memory is visible).  pdl_launch_dependents() in the predecessor
// allows the successor's launch as soon as every one of its own CTAs has got that far.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef __CUDACC__
#include <utility>
// kernel<<<grid, block, smem, ctx->stream>>>(args...) with the programmatic-serialization attribute when `pdl` (and the context allows it)
template <typename... KArgs, typename... Args>
static inline cudaError_t ssdk_launch(ssdk_ctx* ctx, bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    memset(at, 0, sizeof(at));
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && ctx->use_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// streaming (read-once) 128-bit load that does not allocate in L1
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

static inline int ceil_div_i(long long a, long long b) { return (int)((a + b - 1) / b); }

// ----------------------------------------------------------------------------------------------
// Head layout (detector/box_predictor.py:67-104): per-level tower outputs consumed without reshape_and_concatenate.
// Anchor a of image b lives on level l (anchor_off[l] <= a < anchor_off[l+1]) at location loc = (a - anchor_off[l]) / n,
// anchor-in-cell k = (a - anchor_off[l]) % n; its class-c logit is channel k*C + c, its box coordinate j channel k*4 + j.
// ----------------------------------------------------------------------------------------------
struct HeadGeom {
    int num_levels, per_loc, channels_first, C;
    int anchor_off[SSDK_MAX_LEVELS + 1];
    int hw[SSDK_MAX_LEVELS];
    const float* cls[SSDK_MAX_LEVELS];
    const float* box[SSDK_MAX_LEVELS];
};

// Validates an ssdk_head against (B, A, C) and fills the device-side geometry.
int ssdk_head_geom(const ssdk_head* head, int B, int64_t A, int C, bool need_cls, bool need_box, HeadGeom* out);

__device__ __forceinline__ int head_level_of(const HeadGeom& g, int a) {
    int l = 0;
    while (l + 1 < g.num_levels && a >= g.anchor_off[l + 1]) ++l;
    return l;
}
// offset (in floats) of channel ch (of nch per location) at location loc of image b on a level with hw locations
__device__ __forceinline__ long long head_elem(int channels_first, int b, int nch, int hw, int ch, int loc) {
    return channels_first ? ((long long)b * nch + ch) * hw + loc : ((long long)b * hw + loc) * nch + ch;
}
__device__ __forceinline__ float4 head_load_code(const HeadGeom& g, int b, int a) {
    const int l = head_level_of(g, a);
    const int r = a - g.anchor_off[l];
    const int loc = r / g.per_loc, k = r - loc * g.per_loc;
    const int hw = g.hw[l];
    const float* p = g.box[l] + head_elem(g.channels_first, b, g.per_loc * 4, hw, k * 4, loc);
    if (g.channels_first) return make_float4(__ldg(p), __ldg(p + hw), __ldg(p + 2 * (long long)hw), __ldg(p + 3 * (long long)hw));
    return __ldg((const float4*)p);
}
