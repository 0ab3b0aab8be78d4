// Device side of the peer-memory all-reduce (see comm.cu), shared with the fused training-step kernel (train_step.cu), whose
// last CTA performs the exchange itself instead of a separate launch.
#pragma once
#include "common.cuh"

#define COMM_MAX_WORLD 16
#define COMM_MAX_VALUES 8

struct CommMailbox {
    unsigned long long flag[2][COMM_MAX_WORLD];                 // epoch number written by rank r (after its data)
    double data[2][COMM_MAX_WORLD][COMM_MAX_VALUES];
    unsigned long long epoch;                                   // local: number of exchanges done
    unsigned long long error;                                   // local: non-zero after a timeout
};

struct CommPeers {
    CommMailbox* box[COMM_MAX_WORLD];                           // box[rank] is the local one
    int rank, world;
};

// values[0..n) += the same entries of every other rank (in place); n <= COMM_MAX_VALUES.  One CTA of 32 * k threads.
__device__ __forceinline__ void comm_all_reduce(const CommPeers P, double* values, int n) {
    __shared__ unsigned long long s_epoch;
    __shared__ int s_fail;
    CommMailbox* mine = P.box[P.rank];
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_epoch = mine->epoch + 1ull;
        s_fail = 0;
    }
    __syncthreads();
    const unsigned long long e = s_epoch;
    const int slot = (int)(e & 1ull);
    if (tid < P.world) {
        CommMailbox* peer = P.box[tid];
        for (int i = 0; i < n; ++i) peer->data[slot][P.rank][i] = values[i];
        __threadfence_system();                                  // data before flag, visible to the peer device
        *((volatile unsigned long long*)&peer->flag[slot][P.rank]) = e;
    }
    if (tid < P.world) {
        volatile unsigned long long* f = (volatile unsigned long long*)&mine->flag[slot][tid];
        const long long t0 = clock64();
        while (*f != e) {
            if (clock64() - t0 > 30000000000ll) { s_fail = 1; break; }     // ~15 s at 1.9 GHz
            __nanosleep(100);
        }
        __threadfence_system();
    }
    __syncthreads();
    if (tid < n) {
        double t = 0.0;
        for (int r = 0; r < P.world; ++r) t += *((volatile double*)&mine->data[slot][r][tid]);   // rank order: identical everywhere
        values[tid] = s_fail ? __longlong_as_double(0x7ff8000000000000ll) : t;
    }
    if (tid == 0) {
        mine->epoch = e;
        if (s_fail) mine->error = e;
    }
    __syncthreads();
}


// host side (comm.cu): the peer table of a connected context; returns false when the context is not connected
bool ssdk_comm_peers(ssdk_ctx* ctx, CommPeers* out);
