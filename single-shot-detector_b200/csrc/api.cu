// Context, error reporting and workspace management of libssdk.
#include "common.cuh"

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

extern "C" {

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
    SSDK_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SSDK_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    SSDK_REQUIRE(prop.major >= 10, SSDK_ERR_CUDA, "device %d is sm_%d%d; libssdk is built for sm_100a only",
                 device, prop.major, prop.minor);
    ssdk_ctx* c = new ssdk_ctx();
    c->device = device;
    c->stream = (cudaStream_t)stream;
    c->num_sms = prop.multiProcessorCount;
    SSDK_CHECK_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) SSDK_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming));
    *out = c;
    return SSDK_OK;
}

int ssdk_ctx_set_stream(ssdk_ctx* ctx, void* stream) {
    SSDK_REQUIRE(ctx != nullptr, SSDK_ERR_ARG, "null context");
    ctx->stream = (cudaStream_t)stream;
    return SSDK_OK;
}

int ssdk_ctx_destroy(ssdk_ctx* ctx) {
    if (!ctx) return SSDK_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ssdk_buf* bufs[] = {&ctx->ws_gtbest, &ctx->ws_partials, &ctx->ws_reg, &ctx->ws_cls, &ctx->ws_matches,
                        &ctx->ws_cand, &ctx->ws_counts, &ctx->ws_seg};
    for (ssdk_buf* b : bufs) if (b->p) cudaFree(b->p);
    for (ssdk_buf& b : ctx->ws_stage) if (b.p) cudaFree(b.p);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    delete ctx;
    return SSDK_OK;
}

int64_t ssdk_ctx_workspace_bytes(const ssdk_ctx* ctx) {
    if (!ctx) return 0;
    int64_t t = 0;
    const ssdk_buf* bufs[] = {&ctx->ws_gtbest, &ctx->ws_partials, &ctx->ws_reg, &ctx->ws_cls, &ctx->ws_matches,
                              &ctx->ws_cand, &ctx->ws_counts, &ctx->ws_seg};
    for (const ssdk_buf* b : bufs) t += (int64_t)b->cap;
    for (const ssdk_buf& b : ctx->ws_stage) t += (int64_t)b.cap;
    return t;
}

int64_t ssdk_ctx_launch_count(const ssdk_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ssdk_ctx_synchronize(ssdk_ctx* ctx) {
    SSDK_TRY(ssdk_ctx_enter(ctx));
    SSDK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SSDK_OK;
}

}  // extern "C"
