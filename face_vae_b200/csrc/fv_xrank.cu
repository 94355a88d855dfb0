// Cross-rank batch-norm statistic exchange fused with the finalize step, over NVLink peer memory.
//
// nn.SyncBatchNorm under DDP exchanges [mean, invstd, count] with an all_gather per layer in forward and all-reduces
// [sum_dy, sum_dy_xmu] per layer in backward (torch/nn/modules/_functions.py:39-83,144-159; reference modules.py:19,
// logger.py:55): 26 latency-bound collectives per step on the critical path of the anchor model.  Here each exchange is
// ONE single-block kernel: every rank pushes its 2C partial sums straight into a slot of every peer's symmetric buffer
// (stores on NVLink-mapped pointers), waits until all peers' rows have landed in its own buffer, adds the R rows in rank
// order (bitwise identical on every rank) and -- in the same kernel -- produces what the next kernel needs: the
// [mean, invstd, scale, shift] block + running-stat update (forward) or dgamma/dbeta + the two coupling coefficients
// (backward).
// Protocol: every element travels as ONE 8-byte store {value, epoch tag}; the receiver polls the element itself until the
// tag carries the current epoch (the low-latency scheme of NCCL's LL protocol).  There is no separate flag, hence no
// system-scope fence between data and flag: the first version (data, __threadfence_system, flag) paid an extra NVLink round
// trip per exchange.  The epoch lives in device memory and advances by one per launch on every rank, so the launch
// sequence can be captured in a CUDA graph and replayed; a ring of slots keeps a fast rank's next exchange from
// overwriting rows a slow rank is still reading (a rank can start exchange e + 2 only after every peer has finished e).
#include <cstdio>
#include <cstdlib>

#include "../../include/facevae_b200.h"
#include "fv_host.h"
#include "fv_ptx.cuh"

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

__global__ void __launch_bounds__(256, 1) bn_xrank_kernel(const XrankArgs a) {
    __shared__ uint32_t epoch_s;
    __shared__ float tot[kXRow];
    xrank_exchange_finalize(a, tot, &epoch_s);
}

// All ranks of a (virtual) world as the blocks of ONE cooperative launch on one GPU: block r is rank r, with its own
// inputs, outputs and epoch counter at a fixed stride from the base pointers.  Used by the single-GPU parity test of the
// exchange protocol (waiting blocks must be co-resident, which separate launches would not guarantee).
__global__ void __launch_bounds__(256, 1) bn_xrank_emulate_kernel(const XrankArgs base) {
    __shared__ uint32_t epoch_s;
    __shared__ float tot[kXRow];
    XrankArgs a = base;
    const int r = blockIdx.x, C = base.C;
    a.rank = r;
    a.local = base.local + (size_t)r * 2 * C;
    a.epoch_ctr = base.epoch_ctr + r;
    if (base.gamma) a.gamma = base.gamma + (size_t)r * C;
    if (base.beta) a.beta = base.beta + (size_t)r * C;
    if (base.running_mean) a.running_mean = base.running_mean + (size_t)r * C;
    if (base.running_var) a.running_var = base.running_var + (size_t)r * C;
    a.out = base.out + (size_t)r * (base.mode == 0 ? 4 : 2) * C;
    if (base.dgamma) a.dgamma = base.dgamma + (size_t)r * C;
    if (base.dbeta) a.dbeta = base.dbeta + (size_t)r * C;
    xrank_exchange_finalize(a, tot, &epoch_s);
}

static unsigned long long xrank_timeout_ns() {
    static unsigned long long cached = 0;
    if (!cached) {
        const char* v = getenv("FACEVAE_XRANK_TIMEOUT_S");
        double s = v ? atof(v) : 600.0;
        if (!(s > 0)) s = 600.0;
        cached = (unsigned long long)(s * 1e9);
    }
    return cached;
}

}  // namespace fv

extern "C" __attribute__((visibility("default"))) long long fv_xrank_buffer_floats(void) {
    return (long long)(2 * (size_t)fv::kXSlots * fv::kXMaxWorld * fv::kXRow);      // 8-byte tagged elements
}

static int xrank_check(const char* who, const float* sums_local, void* peer_bufs_dev, int rank, int world, void* epoch_ctr, int mode, double count,
                       const float* gamma, const float* beta, float* out, int C) {
    using namespace fv;
    if (!sums_local || !peer_bufs_dev || !epoch_ctr || !out || count <= 0) return fail(FV_ERR_ARG, "%s: bad arguments", who);
    if (world < 1 || world > kXMaxWorld || rank < 0 || rank >= world) return fail(FV_ERR_ARG, "%s: rank %d / world %d", who, rank, world);
    if (2 * C > kXRow) return fail(FV_ERR_UNSUPPORTED, "%s: C=%d exceeds %d", who, C, kXRow / 2);
    if (mode == 0 && (!gamma || !beta)) return fail(FV_ERR_ARG, "%s: gamma/beta required in forward mode", who);
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_bn_finalize_xrank(const float* sums_local, void* peer_bufs_dev, int rank, int world,
                                                                         void* epoch_ctr, int mode, double count, const float* gamma,
                                                                         const float* beta, float* running_mean, float* running_var,
                                                                         float momentum, float eps, float* out, float* dgamma, float* dbeta,
                                                                         int accumulate, int C, void* stream) {
    using namespace fv;
    if (int e = xrank_check("fv_bn_finalize_xrank", sums_local, peer_bufs_dev, rank, world, epoch_ctr, mode, count, gamma, beta, out, C)) return e;
    XrankArgs a{sums_local, reinterpret_cast<unsigned long long* const*>(peer_bufs_dev), rank, world, reinterpret_cast<unsigned long long*>(epoch_ctr),
                C, mode, count, gamma, beta, running_mean, running_var, momentum, eps, out, dgamma, dbeta, accumulate, xrank_timeout_ns()};
    bn_xrank_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a);
    FV_LAUNCH_CHECK("bn_xrank_kernel");
    return FV_OK;
}

// Test entry: `world` ranks emulated as the blocks of one cooperative launch.  Every per-rank array is the concatenation of
// the ranks' arrays (sums_local [world][2C], gamma / beta / running_* / dgamma / dbeta [world][C], out [world][4C or 2C],
// epoch_ctr [world]); peer_bufs_dev: `world` buffers of fv_xrank_buffer_floats() floats on this device.
extern "C" __attribute__((visibility("default"))) int fv_bn_finalize_xrank_emulate(const float* sums_local, void* peer_bufs_dev, int world, void* epoch_ctr,
                                                                                 int mode, double count, const float* gamma, const float* beta,
                                                                                 float* running_mean, float* running_var, float momentum, float eps,
                                                                                 float* out, float* dgamma, float* dbeta, int accumulate, int C,
                                                                                 void* stream) {
    using namespace fv;
    if (int e = xrank_check("fv_bn_finalize_xrank_emulate", sums_local, peer_bufs_dev, 0, world, epoch_ctr, mode, count, gamma, beta, out, C)) return e;
    XrankArgs a{sums_local, reinterpret_cast<unsigned long long* const*>(peer_bufs_dev), 0, world, reinterpret_cast<unsigned long long*>(epoch_ctr),
                C, mode, count, gamma, beta, running_mean, running_var, momentum, eps, out, dgamma, dbeta, accumulate, xrank_timeout_ns()};
    void* args[] = {&a};
    FV_CUDA(cudaLaunchCooperativeKernel((const void*)bn_xrank_emulate_kernel, dim3(world), dim3(256), args, 0, (cudaStream_t)stream));
    return FV_OK;
}
