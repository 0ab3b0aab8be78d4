// The one exchange step of the path, as a hand-written NVLink / NVSwitch collective.
//
// Images are sharded over the GPUs of a box (one process per GPU); the only coupling between shards is the loss
// normaliser and the two loss sums: num_matches = sum over ALL images (detector/ssd.py:121-123), losses = sums / max(n, 1)
// (ssd.py:131-133).  That is an all-reduce(sum) of three doubles -- 24 bytes, pure latency.  NCCL needs a kernel of its
// own plus its proxy handshake for it (measured here: ~26 us inside a 166 us step at 8 GPUs); this file does it with
// plain peer-memory stores over NVLink inside the SAME kernel that finalizes the losses:
//
//   * every rank owns a small "mailbox" in its device memory, exported with cudaIpcGetMemHandle and mapped by all peers
//     (ssdk_comm_local_handle / ssdk_comm_connect; the 64-byte handles travel through torch.distributed, plumbing only);
//   * all-reduce = each rank stores its values into slot [epoch parity][its rank] of EVERY mailbox (one thread per peer,
//     st.global over NVLink), __threadfence_system(), then stores the epoch number into the matching flag; it then polls
//     its OWN mailbox (local HBM, no traffic on the links) until all `world` flags carry the epoch, and adds the values
//     in rank order -- every rank computes bit-identical sums, deterministically;
//   * the epoch counter lives in device memory and is advanced by the kernel itself, so the kernel can be captured in a
//     CUDA graph and replayed; two parity slots suffice because a rank can only be one exchange ahead of a peer (it
//     needs the peer's flag of exchange e+1, which the peer writes after it has finished reading exchange e).
// A rank whose peers never arrive gives up after ~15 s of polling, writes NaN and raises the mailbox's error word instead
// of hanging the GPU.
#include <string.h>

#include "comm.cuh"

struct ssdk_comm {
    CommMailbox* local = nullptr;
    CommPeers peers;
    bool connected = false;
    void* opened[COMM_MAX_WORLD] = {nullptr};
};

static ssdk_comm* comm_of(ssdk_ctx* ctx) { return (ssdk_comm*)ctx->comm; }

bool ssdk_comm_peers(ssdk_ctx* ctx, CommPeers* out) {
    ssdk_comm* c = comm_of(ctx);
    if (!c || !c->connected) return false;
    *out = c->peers;
    return true;
}

__global__ void __launch_bounds__(32) comm_all_reduce_kernel(const CommPeers P, double* values, int n) { comm_all_reduce(P, values, n); }

// all-reduce of the three loss sums + ssd.py:123,131-133 in one launch
__global__ void __launch_bounds__(32) comm_finalize_kernel(const CommPeers P, double* sums, float* out) {
    comm_all_reduce(P, sums, 3);
    if (threadIdx.x == 0) {
        const double norm = fmax(sums[2], 1.0);
        out[0] = (float)(sums[0] / norm);
        out[1] = (float)(sums[1] / norm);
    }
}

extern "C" {

int ssdk_comm_local_handle(ssdk_ctx* ctx, void* out_handle) {
    SSDK_ENTER(ctx);
    SSDK_REQUIRE(out_handle != nullptr, SSDK_ERR_ARG, "ssdk_comm_local_handle: out_handle is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == SSDK_COMM_HANDLE_BYTES, "handle size");
    ssdk_comm* c = comm_of(ctx);
    if (!c) {
        c = new ssdk_comm();
        ctx->comm = c;
    }
    if (!c->local) {
        SSDK_CHECK_CUDA(cudaMalloc((void**)&c->local, sizeof(CommMailbox)));
        SSDK_CHECK_CUDA(cudaMemset(c->local, 0, sizeof(CommMailbox)));
        SSDK_CHECK_CUDA(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    SSDK_CHECK_CUDA(cudaIpcGetMemHandle(&h, c->local));
    memcpy(out_handle, &h, sizeof(h));
    return SSDK_OK;
}

int ssdk_comm_connect(ssdk_ctx* ctx, int rank, int world, const void* handles) {
    SSDK_ENTER(ctx);
    ssdk_comm* c = comm_of(ctx);
    SSDK_REQUIRE(c && c->local, SSDK_ERR_ARG, "ssdk_comm_connect: call ssdk_comm_local_handle first");
    SSDK_REQUIRE(world >= 1 && world <= COMM_MAX_WORLD && rank >= 0 && rank < world && handles, SSDK_ERR_ARG,
                 "ssdk_comm_connect: bad rank %d / world %d (at most %d ranks)", rank, world, COMM_MAX_WORLD);
    SSDK_REQUIRE(!c->connected, SSDK_ERR_ARG, "ssdk_comm_connect: already connected");
    for (int r = 0; r < COMM_MAX_WORLD; ++r) c->peers.box[r] = nullptr;
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            c->peers.box[r] = c->local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; ++q)
                if (c->opened[q]) { cudaIpcCloseMemHandle(c->opened[q]); c->opened[q] = nullptr; }
            ssdk_set_error("ssdk_comm_connect: cannot map the mailbox of rank %d (%s); peers must be GPUs of the same box with "
                           "P2P access", r, cudaGetErrorString(e));
            return SSDK_ERR_NCCL;
        }
        c->opened[r] = p;
        c->peers.box[r] = (CommMailbox*)p;
    }
    c->peers.rank = rank;
    c->peers.world = world;
    c->connected = true;
    return SSDK_OK;
}

int ssdk_comm_world(const ssdk_ctx* ctx) {
    const ssdk_comm* c = ctx ? (const ssdk_comm*)ctx->comm : nullptr;
    return (c && c->connected) ? c->peers.world : 0;
}

int ssdk_comm_all_reduce_sum(ssdk_ctx* ctx, double* values, int n) {
    SSDK_ENTER(ctx);
    ssdk_comm* c = comm_of(ctx);
    SSDK_REQUIRE(c && c->connected, SSDK_ERR_ARG, "ssdk_comm_all_reduce_sum: not connected");
    SSDK_REQUIRE(values && n >= 1 && n <= COMM_MAX_VALUES, SSDK_ERR_ARG, "ssdk_comm_all_reduce_sum: n must be in [1,%d]", COMM_MAX_VALUES);
    SSDK_KERNEL(ctx, SSDK_K_COMM, comm_all_reduce_kernel<<<1, 32, 0, ctx->stream>>>(c->peers, values, n));
    return SSDK_OK;
}

int ssdk_comm_loss_finalize(ssdk_ctx* ctx, double* sums, float* out_losses) {
    SSDK_ENTER(ctx);
    ssdk_comm* c = comm_of(ctx);
    SSDK_REQUIRE(c && c->connected, SSDK_ERR_ARG, "ssdk_comm_loss_finalize: not connected");
    SSDK_REQUIRE(sums && out_losses, SSDK_ERR_ARG, "ssdk_comm_loss_finalize: null pointer");
    SSDK_KERNEL(ctx, SSDK_K_COMM, comm_finalize_kernel<<<1, 32, 0, ctx->stream>>>(c->peers, sums, out_losses));
    return SSDK_OK;
}

int ssdk_comm_error(ssdk_ctx* ctx, int64_t* out_epoch_of_timeout) {
    SSDK_ENTER(ctx);
    ssdk_comm* c = comm_of(ctx);
    SSDK_REQUIRE(c && c->local && out_epoch_of_timeout, SSDK_ERR_ARG, "ssdk_comm_error: no communicator");
    SSDK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned long long e = 0;
    SSDK_CHECK_CUDA(cudaMemcpy(&e, &c->local->error, sizeof(e), cudaMemcpyDeviceToHost));
    *out_epoch_of_timeout = (int64_t)e;
    return SSDK_OK;
}

int ssdk_comm_disconnect(ssdk_ctx* ctx) {
    if (!ctx || !ctx->comm) return SSDK_OK;
    SsdkDeviceGuard dev_guard(ctx);
    ssdk_comm* c = comm_of(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < COMM_MAX_WORLD; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->opened[r]);
    if (c->local) cudaFree(c->local);
    delete c;
    ctx->comm = nullptr;
    return SSDK_OK;
}

}  // extern "C"
