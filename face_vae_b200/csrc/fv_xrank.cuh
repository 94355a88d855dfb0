// Device side of the cross-rank batch-norm statistic exchange (see fv_xrank.cu for the protocol): shared by the stand-alone
// exchange kernel and by the reduction kernels that run the exchange in the tail of their LAST block (fv_glue.cu), which
// removes one single-block launch per batch-norm layer and pass from the critical path of the data-parallel step.
#pragma once
#include <cstdio>
#include <stdint.h>
#include <cuda_runtime.h>

namespace fv {

static constexpr int kXSlots = 8;          // ring of exchange slots
static constexpr int kXRow = 1024;         // elements per (slot, rank) row: 2 * C_max
static constexpr int kXMaxWorld = 16;
// symmetric buffer layout: rows[kXSlots][kXMaxWorld][kXRow] of {float value, uint32 epoch} (8 bytes each), zero-initialised

__device__ __forceinline__ void st_tagged_sys(unsigned long long* p, float v, uint32_t tag) {
    asm volatile("st.relaxed.sys.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ void ld_tagged_sys(const unsigned long long* p, float& v, uint32_t& tag) {
    uint32_t a, b;
    asm volatile("ld.relaxed.sys.global.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
    v = __uint_as_float(a);
    tag = b;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct XrankArgs {
    const float* local;                       // [2C] this rank's partial sums
    unsigned long long* const* peer_bufs;     // [world] symmetric buffers (peer-mapped)
    int rank, world;
    unsigned long long* epoch_ctr;
    int C, mode;                              // mode 0: forward finalize, mode 1: backward finalize
    double count;
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    float momentum, eps;
    float* out;                               // fwd: stat[4][C]; bwd: coef[2][C]
    float* dgamma;
    float* dbeta;
    int accumulate;
    unsigned long long timeout_ns;            // wall-clock bound on the wait for a peer (FACEVAE_XRANK_TIMEOUT_S, default 600 s)
};

// One thread block: push, gather in rank order, finalize.  `tot` (shared, kXRow floats) and `epoch_s` are the caller's.
__device__ __forceinline__ void xrank_exchange_finalize(const XrankArgs& a, float* tot, uint32_t* epoch_s) {
    const int n = 2 * a.C;
    if (threadIdx.x == 0) *epoch_s = (uint32_t)(atomicAdd(a.epoch_ctr, 1ULL) + 1ULL);
    __syncthreads();
    const uint32_t epoch = *epoch_s;
    const size_t slot_base = (size_t)(epoch % kXSlots) * kXMaxWorld * kXRow;
    // push my partial sums, tagged with the epoch, into row (slot, rank) of every peer (and of myself)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = a.local[i];
        for (int p = 0; p < a.world; ++p) st_tagged_sys(a.peer_bufs[p] + slot_base + (size_t)a.rank * kXRow + i, v, epoch);
    }
    // gather: poll every element of every rank's row in MY buffer until it carries this epoch; fixed summation order.
    // The bound on the wait is wall-clock and long (a peer may be writing a checkpoint, evaluating, or paging in a first
    // step): NCCL tolerates minutes, so does this -- round 1 counted 2^24 polls (a few seconds) and killed the job.
    const unsigned long long* mine = a.peer_bufs[a.rank] + slot_base;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float acc = 0.f;
        for (int p = 0; p < a.world; ++p) {
            float v;
            uint32_t tag, spins = 0;
            unsigned long long t0 = 0;
            ld_tagged_sys(mine + (size_t)p * kXRow + i, v, tag);
            while (tag != epoch) {
                if ((++spins & 0xFFFu) == 0) {                 // look at the clock every 4096 polls
                    const unsigned long long now = global_ns();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > a.timeout_ns) {
                        printf("fv: cross-rank BN exchange timed out after %llu s (rank %d waiting for rank %d, epoch %u)\n",
                               a.timeout_ns / 1000000000ULL, a.rank, p, epoch);
                        __trap();
                    }
                    __nanosleep(200);
                }
                ld_tagged_sys(mine + (size_t)p * kXRow + i, v, tag);
            }
            acc += v;
        }
        tot[i] = acc;
    }
    __syncthreads();
    const int C = a.C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        if (a.mode == 0) {
            const double mean = (double)tot[c] / a.count;
            double var = (double)tot[C + c] / a.count - mean * mean;
            if (var < 0) var = 0;
            const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
            const float sc = a.gamma[c] * invstd;
            a.out[c] = (float)mean;
            a.out[C + c] = invstd;
            a.out[2 * C + c] = sc;
            a.out[3 * C + c] = a.beta[c] - (float)mean * sc;
            if (a.running_mean) {
                const double unbiased = a.count > 1 ? var * a.count / (a.count - 1) : var;
                a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * (float)mean;
                a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unbiased;
            }
        } else {
            const float s1 = a.local[c], s2 = a.local[C + c];
            if (a.dbeta) a.dbeta[c] = a.accumulate ? a.dbeta[c] + s1 : s1;
            if (a.dgamma) a.dgamma[c] = a.accumulate ? a.dgamma[c] + s2 : s2;
            a.out[c] = (float)((double)tot[c] / a.count);
            a.out[C + c] = (float)((double)tot[C + c] / a.count);
        }
    }
}

}  // namespace fv
