// Context, error reporting and workspace management of libssdk.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

static thread_local char g_err[512] = "";

void ssdk_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ssdk_ensure(ssdk_ctx* ctx, ssdk_buf* b, size_t bytes) {
    if (bytes <= b->cap) return SSDK_OK;
    if (b->p) {
        // the old block may still be in use by work queued on the stream
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ssdk_set_error("stream sync before workspace growth: %s", cudaGetErrorString(e)); return SSDK_ERR_CUDA; }
        cudaFree(b->p);
        b->p = nullptr; b->cap = 0;
    }
    size_t want = (bytes + (1u << 20) - 1) & ~((size_t)(1u << 20) - 1);
    cudaError_t e = cudaMalloc(&b->p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ssdk_set_error("workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        b->p = nullptr;
        return SSDK_ERR_NOMEM;
    }
    b->cap = want;
    return SSDK_OK;
}

SsdkWsGuard::SsdkWsGuard(ssdk_ctx* c, int g) : ctx(c), group(g) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &st) != cudaSuccess) { cudaGetLastError(); st = cudaStreamCaptureStatusNone; }
    capturing = st != cudaStreamCaptureStatusNone;
    if (capturing) return;
    if (ctx->ws_valid[group] && ctx->ws_stream[group] != ctx->stream) cudaStreamWaitEvent(ctx->stream, ctx->ws_event[group], 0);
}

SsdkWsGuard::~SsdkWsGuard() {
    if (capturing) {
        ctx->ws_valid[group] = false;               // the last use is inside a graph: nothing an eager call could wait on
        return;
    }
    if (!ctx->ws_event[group] && cudaEventCreateWithFlags(&ctx->ws_event[group], cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        ctx->ws_event[group] = nullptr;
        ctx->ws_valid[group] = false;
        return;
    }
    ctx->ws_valid[group] = cudaEventRecord(ctx->ws_event[group], ctx->stream) == cudaSuccess;
    ctx->ws_stream[group] = ctx->stream;
}

// Raise a kernel's dynamic shared-memory limit once per (device, kernel) instead of on every launch.  The attribute is
// a property of the function on the device, not of a context, so the record is process-wide (contexts of other host
// threads -- e.g. the autograd engine's -- launch the same kernels) and the limit is only ever raised.
#include <mutex>
static std::mutex g_smem_mutex;
static struct { int device; const void* func; int bytes; } g_smem[64];
static int g_smem_n = 0;

int ssdk_set_max_smem(ssdk_ctx* ctx, const void* func, int bytes) {
    std::lock_guard<std::mutex> lock(g_smem_mutex);
    int slot = -1;
    for (int i = 0; i < g_smem_n; ++i)
        if (g_smem[i].device == ctx->device && g_smem[i].func == func) { slot = i; break; }
    if (slot >= 0 && g_smem[slot].bytes >= bytes) return SSDK_OK;
    SSDK_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (slot < 0 && g_smem_n < 64) slot = g_smem_n++;
    if (slot >= 0) { g_smem[slot].device = ctx->device; g_smem[slot].func = func; g_smem[slot].bytes = bytes; }
    return SSDK_OK;
}

static int prof_drain(ssdk_ctx* ctx) {
    if (ctx->prof_n == 0) return SSDK_OK;
    SSDK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < ctx->prof_n; ++i) {
        float ms = 0.f;
        SSDK_CHECK_CUDA(cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        ctx->prof_ms[ctx->prof_id[i]] += ms;
        ctx->prof_calls[ctx->prof_id[i]] += 1;
    }
    ctx->prof_n = 0;
    return SSDK_OK;
}

int ssdk_prof_begin(ssdk_ctx* ctx, int id) {
    if (ctx->prof_n >= SSDK_PROFILE_EVENTS && prof_drain(ctx) != SSDK_OK) return -1;
    const int slot = ctx->prof_n++;
    ctx->prof_id[slot] = id;
    cudaEventRecord(ctx->prof_ev[2 * slot], ctx->stream);
    return slot;
}

void ssdk_prof_end(ssdk_ctx* ctx, int slot) { cudaEventRecord(ctx->prof_ev[2 * slot + 1], ctx->stream); }

// integer knob from the environment: unset -> dflt; otherwise clamped to [lo, hi] and rounded down to a multiple of `mult`
static int env_int(const char* name, int dflt, int lo, int hi, int mult) {
    const char* e = getenv(name);
    if (!e || !*e) return dflt;
    long v = strtol(e, nullptr, 10);
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    v -= v % mult;
    if (v < lo) v = lo;
    return (int)v;
}

extern "C" {

int ssdk_ctx_set_profiling(ssdk_ctx* ctx, int enable) {
    SSDK_ENTER(ctx);
    if (enable && !ctx->prof_ev) {
        ctx->prof_ev = new cudaEvent_t[2 * SSDK_PROFILE_EVENTS];
        for (int i = 0; i < 2 * SSDK_PROFILE_EVENTS; ++i) SSDK_CHECK_CUDA(cudaEventCreate(&ctx->prof_ev[i]));
    }
    SSDK_TRY(prof_drain(ctx));
    ctx->profiling = enable ? 1 : 0;
    return SSDK_OK;
}

int ssdk_ctx_profile_read(ssdk_ctx* ctx, double* out_ms, int64_t* out_calls, int n, int reset) {
    SSDK_ENTER(ctx);
    SSDK_TRY(prof_drain(ctx));
    for (int i = 0; i < n && i < SSDK_K_COUNT; ++i) {
        if (out_ms) out_ms[i] = ctx->prof_ms[i];
        if (out_calls) out_calls[i] = ctx->prof_calls[i];
    }
    if (reset) for (int i = 0; i < SSDK_K_COUNT; ++i) { ctx->prof_ms[i] = 0; ctx->prof_calls[i] = 0; }
    return SSDK_OK;
}

int ssdk_version(void) { return SSDK_VERSION; }

const char* ssdk_last_error(void) { return g_err; }

int ssdk_ctx_create(int device, void* stream, ssdk_ctx** out) {
    SSDK_REQUIRE(out != nullptr, SSDK_ERR_ARG, "ssdk_ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        ssdk_set_error("no CUDA device available (%s); libssdk has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return SSDK_ERR_CUDA;
    }
    SSDK_REQUIRE(device >= 0 && device < count, SSDK_ERR_ARG, "device %d out of range [0,%d)", device, count);
    struct Restore {                                    // the caller's current device is left as it was
        int prev = -1;
        Restore() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
        ~Restore() { if (prev >= 0) cudaSetDevice(prev); }
    } restore;
    SSDK_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SSDK_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    SSDK_REQUIRE(prop.major >= 10, SSDK_ERR_CUDA, "device %d is sm_%d%d; libssdk is built for sm_100a only",
                 device, prop.major, prop.minor);
    ssdk_ctx* c = new ssdk_ctx();
    c->device = device;
    c->stream = (cudaStream_t)stream;
    c->num_sms = prop.multiProcessorCount;
    for (int i = 0; i < 4; ++i) SSDK_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming));
    SSDK_CHECK_CUDA(cudaMalloc((void**)&c->dev_err, 16));
    SSDK_CHECK_CUDA(cudaMemset(c->dev_err, 0, 16));
    // tuning knobs: read once, forced into their valid ranges (an out-of-range value cannot reach a kernel)
    c->tune_head_ctas = env_int("SSDK_HEAD_CTAS", 0, 1, 8, 1);            // CTAs per SM of the stand-alone flat pass
    c->tune_loss_rpw = env_int("SSDK_LOSS_RPW", 0, 4, 32, 4);             // rows per warp of ssd_loss_kernel: a multiple of 4 (16-byte TMA tiles)
    c->tune_loss_stages = env_int("SSDK_LOSS_STAGES", 0, 2, 4, 1);        // TMA ring depth
    c->tune_loss_ctas = env_int("SSDK_LOSS_CTAS", 0, 1, 8, 1);            // the partials buffer holds num_sms * 8 CTAs
    c->match_ctas_per_sm = env_int("SSDK_MATCH_CTAS", 0, 0, 7, 1);
    c->match_flat_share_pct = env_int("SSDK_MATCH_FLAT_SHARE", -1, -1, 100, 1);
    c->train_ctas_per_sm = env_int("SSDK_TRAIN_CTAS", 0, 0, 8, 1);
    c->use_pdl = env_int("SSDK_PDL", 1, 0, 1, 1);
    c->train_dynamic_chunks = env_int("SSDK_TRAIN_DYNAMIC", 1, 0, 1, 1);
    *out = c;
    return SSDK_OK;
}

int ssdk_ctx_set_option(ssdk_ctx* ctx, int option, int value) {
    SSDK_REQUIRE(ctx != nullptr, SSDK_ERR_ARG, "null context");
    switch (option) {
        case SSDK_OPT_FUSED_TRAIN_STEP: ctx->fused_train_step = value ? 1 : 0; return SSDK_OK;
        case SSDK_OPT_MATCH_CTAS_PER_SM:
            SSDK_REQUIRE(value >= 0 && value <= 7, SSDK_ERR_ARG, "SSDK_OPT_MATCH_CTAS_PER_SM must be in [0,7] (got %d)", value);
            ctx->match_ctas_per_sm = value;
            return SSDK_OK;
        case SSDK_OPT_PROGRAMMATIC_LAUNCH: ctx->use_pdl = value ? 1 : 0; return SSDK_OK;
        case SSDK_OPT_TRAIN_DYNAMIC_CHUNKS: ctx->train_dynamic_chunks = value ? 1 : 0; return SSDK_OK;
        case SSDK_OPT_TRAIN_CTAS_PER_SM:
            SSDK_REQUIRE(value >= 0 && value <= 8, SSDK_ERR_ARG, "SSDK_OPT_TRAIN_CTAS_PER_SM must be in [0,8] (got %d)", value);
            ctx->train_ctas_per_sm = value;
            return SSDK_OK;
        case SSDK_OPT_MATCH_FLAT_SHARE_PCT:
            SSDK_REQUIRE(value >= -1 && value <= 100, SSDK_ERR_ARG, "SSDK_OPT_MATCH_FLAT_SHARE_PCT must be in [-1,100] (got %d)", value);
            ctx->match_flat_share_pct = value;
            return SSDK_OK;
        default: ssdk_set_error("ssdk_ctx_set_option: unknown option %d", option); return SSDK_ERR_ARG;
    }
}

int ssdk_ctx_set_stream(ssdk_ctx* ctx, void* stream) {
    SSDK_REQUIRE(ctx != nullptr, SSDK_ERR_ARG, "null context");
    ctx->stream = (cudaStream_t)stream;
    return SSDK_OK;
}

int ssdk_ctx_destroy(ssdk_ctx* ctx) {
    if (!ctx) return SSDK_OK;
    ssdk_comm_disconnect(ctx);
    int prev_dev = -1;
    if (cudaGetDevice(&prev_dev) != cudaSuccess) prev_dev = -1;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ssdk_buf* bufs[] = {&ctx->ws_gtbest, &ctx->ws_partials, &ctx->ws_reg, &ctx->ws_cls, &ctx->ws_matches,
                        &ctx->ws_cand, &ctx->ws_counts, &ctx->ws_seg, &ctx->ws_head, &ctx->ws_summ, &ctx->ws_train};
    for (ssdk_buf* b : bufs) if (b->p) cudaFree(b->p);
    for (ssdk_buf& b : ctx->ws_stage) if (b.p) cudaFree(b.p);
    if (ctx->prof_ev) {
        for (int i = 0; i < 2 * SSDK_PROFILE_EVENTS; ++i) cudaEventDestroy(ctx->prof_ev[i]);
        delete[] ctx->prof_ev;
    }
    for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->dev_err) cudaFree(ctx->dev_err);
    for (int i = 0; i < 8; ++i) if (ctx->ws_event[i]) cudaEventDestroy(ctx->ws_event[i]);
    const int own_dev = ctx->device;
    delete ctx;
    if (prev_dev >= 0 && prev_dev != own_dev) cudaSetDevice(prev_dev);
    return SSDK_OK;
}

int64_t ssdk_ctx_workspace_bytes(const ssdk_ctx* ctx) {
    if (!ctx) return 0;
    int64_t t = 0;
    const ssdk_buf* bufs[] = {&ctx->ws_gtbest, &ctx->ws_partials, &ctx->ws_reg, &ctx->ws_cls, &ctx->ws_matches,
                              &ctx->ws_cand, &ctx->ws_counts, &ctx->ws_seg, &ctx->ws_head, &ctx->ws_summ, &ctx->ws_train};
    for (const ssdk_buf* b : bufs) t += (int64_t)b->cap;
    for (const ssdk_buf& b : ctx->ws_stage) t += (int64_t)b.cap;
    return t;
}

int64_t ssdk_ctx_launch_count(const ssdk_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ssdk_ctx_async_error(ssdk_ctx* ctx, int* out_code) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(out_code != nullptr, SSDK_ERR_ARG, "ssdk_ctx_async_error: out_code is NULL");
    SSDK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    SSDK_CHECK_CUDA(cudaMemcpy(out_code, ctx->dev_err, sizeof(int), cudaMemcpyDeviceToHost));
    return SSDK_OK;
}

int ssdk_ctx_synchronize(ssdk_ctx* ctx) {
    SSDK_ENTER(ctx);
    SSDK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SSDK_OK;
}

}  // extern "C"
