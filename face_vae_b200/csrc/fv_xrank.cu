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
#include "fv_xrank.cuh"

namespace fv {

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
