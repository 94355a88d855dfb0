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

// mode 0: forward finalize, mode 1: backward finalize
__global__ void __launch_bounds__(256, 1)
bn_xrank_kernel(const float* __restrict__ local, unsigned long long* const* __restrict__ peer_bufs, int rank, int world,
                unsigned long long* __restrict__ epoch_ctr, int C, int mode, double count,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean, float* running_var,
                float momentum, float eps, float* __restrict__ out /* fwd: stat[4][C]; bwd: coef[2][C] */,
                float* dgamma, float* dbeta, int accumulate) {
    __shared__ uint32_t epoch_s;
    __shared__ float tot[kXRow];
    const int n = 2 * C;
    if (threadIdx.x == 0) epoch_s = (uint32_t)(atomicAdd(epoch_ctr, 1ULL) + 1ULL);
    __syncthreads();
    const uint32_t epoch = epoch_s;
    const size_t slot_base = (size_t)(epoch % kXSlots) * kXMaxWorld * kXRow;
    // push my partial sums, tagged with the epoch, into row (slot, rank) of every peer (and of myself)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = local[i];
        for (int p = 0; p < world; ++p) st_tagged_sys(peer_bufs[p] + slot_base + (size_t)rank * kXRow + i, v, epoch);
    }
    // gather: poll every element of every rank's row in MY buffer until it carries this epoch; fixed summation order
    const unsigned long long* mine = peer_bufs[rank] + slot_base;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float a = 0.f;
        for (int p = 0; p < world; ++p) {
            float v;
            uint32_t tag, spins = 0;
            ld_tagged_sys(mine + (size_t)p * kXRow + i, v, tag);
            while (tag != epoch) {
                if (++spins > (1u << 24)) {
                    printf("fv: cross-rank BN exchange timed out (rank %d waiting for rank %d, epoch %u)\n", rank, p, epoch);
                    __trap();
                }
                ld_tagged_sys(mine + (size_t)p * kXRow + i, v, tag);
            }
            a += v;
        }
        tot[i] = a;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        if (mode == 0) {
            const double mean = (double)tot[c] / count;
            double var = (double)tot[C + c] / count - mean * mean;
            if (var < 0) var = 0;
            const float invstd = (float)(1.0 / sqrt(var + (double)eps));
            const float sc = gamma[c] * invstd;
            out[c] = (float)mean;
            out[C + c] = invstd;
            out[2 * C + c] = sc;
            out[3 * C + c] = beta[c] - (float)mean * sc;
            if (running_mean) {
                const double unbiased = count > 1 ? var * count / (count - 1) : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            }
        } else {
            const float s1 = local[c], s2 = local[C + c];
            if (dbeta) dbeta[c] = accumulate ? dbeta[c] + s1 : s1;
            if (dgamma) dgamma[c] = accumulate ? dgamma[c] + s2 : s2;
            out[c] = (float)((double)tot[c] / count);
            out[C + c] = (float)((double)tot[C + c] / count);
        }
    }
}

}  // namespace fv

extern "C" __attribute__((visibility("default"))) long long fv_xrank_buffer_floats(void) {
    return (long long)(2 * (size_t)fv::kXSlots * fv::kXMaxWorld * fv::kXRow);      // 8-byte tagged elements
}

extern "C" __attribute__((visibility("default"))) int fv_bn_finalize_xrank(const float* sums_local, void* peer_bufs_dev, int rank, int world,
                                                                         void* epoch_ctr, int mode, double count, const float* gamma,
                                                                         const float* beta, float* running_mean, float* running_var,
                                                                         float momentum, float eps, float* out, float* dgamma, float* dbeta,
                                                                         int accumulate, int C, void* stream) {
    using namespace fv;
    if (!sums_local || !peer_bufs_dev || !epoch_ctr || !out || count <= 0) return fail(FV_ERR_ARG, "fv_bn_finalize_xrank: bad arguments");
    if (world < 1 || world > kXMaxWorld || rank < 0 || rank >= world) return fail(FV_ERR_ARG, "fv_bn_finalize_xrank: rank %d / world %d", rank, world);
    if (2 * C > kXRow) return fail(FV_ERR_UNSUPPORTED, "fv_bn_finalize_xrank: C=%d exceeds %d", C, kXRow / 2);
    if (mode == 0 && (!gamma || !beta)) return fail(FV_ERR_ARG, "fv_bn_finalize_xrank: gamma/beta required in forward mode");
    bn_xrank_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sums_local, reinterpret_cast<unsigned long long* const*>(peer_bufs_dev), rank, world,
                                                         reinterpret_cast<unsigned long long*>(epoch_ctr), C, mode, count, gamma, beta,
                                                         running_mean, running_var, momentum, eps, out, dgamma, dbeta, accumulate);
    FV_LAUNCH_CHECK("bn_xrank_kernel");
    return FV_OK;
}
