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

// ---- gradient mean all-reduce over peer memory (reference: DistributedDataParallel's bucketed NCCL all-reduce, logger.py:55) ---
// The flat gradient buffer of distributed.GradientReducer is a symmetric allocation: every rank can address every rank's copy.
// ONE kernel per step, two-shot and in place: rank r owns the r-th slice; it PULLS that slice from all ranks over NVLink, adds
// the R values in rank order (so the result is bitwise identical everywhere and run to run), scales by 1/R and PUSHES the
// result into the same offsets of all R buffers.  Two flag barriers (64-bit epochs in a small symmetric flag array
// [2][kXMaxWorld]): "my gradients are complete" before the pull, "my pushes have landed and I no longer read your buffer" after
// it -- the second one is what allows the next kernel (Adam) to read the buffer and the next backward to overwrite it.
// NCCL needed ~80 us (one bucket) to ~105 us (2 MB buckets; its CTAs displace the persistent one-CTA-per-SM convolution kernels
// they are supposed to overlap with) for these 15 MB at 2 GPUs.
struct GradArArgs {
    float* const* bufs;                   // [world] peer-mapped flat buffers
    unsigned long long* const* flags;     // [world] peer-mapped flag arrays [2][kXMaxWorld]
    int rank, world;
    long long n4;                         // float4 elements of the buffer
    unsigned long long* epoch_ctr;
    unsigned int* ticket;
    float scale;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_flag_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_flag_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_f4_sys(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_flag(const unsigned long long* p, unsigned long long e, unsigned long long timeout_ns, int rank, int peer,
                                          const char* what) {
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while (ld_flag_sys(p) < e) {
        if ((++spins & 0x3FFu) == 0) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > timeout_ns) {
                printf("fv: gradient all-reduce timed out after %llu s (rank %d waiting for rank %d, %s, epoch %llu)\n", timeout_ns / 1000000000ULL,
                       rank, peer, what, e);
                __trap();
            }
            __nanosleep(100);
        }
    }
}

// W: world size known at compile time (0 = run-time loop); U float4 elements per thread and iteration, U * W loads in flight
template <int W, int U>
__global__ void __launch_bounds__(256) grad_allreduce_kernel(const GradArArgs a) {
    const int world = W ? W : a.world;
    const unsigned long long e = *a.epoch_ctr + 1ULL;
    unsigned long long* my_flags = a.flags[a.rank];
    if (blockIdx.x == 0 && threadIdx.x < world && (int)threadIdx.x != a.rank) st_flag_sys(a.flags[threadIdx.x] + a.rank, e);
    if (threadIdx.x < world && (int)threadIdx.x != a.rank) wait_flag(my_flags + threadIdx.x, e, a.timeout_ns, a.rank, threadIdx.x, "gradients ready");
    __syncthreads();
    const long long lo = a.n4 * a.rank / world, hi = a.n4 * (a.rank + 1) / world;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * U) {
        float4 v[U][W ? W : 1];
        if (W) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int p = 0; p < (W ? W : 1); ++p)
                    if (i0 + u * stride < hi) v[u][p] = ld_f4_sys(reinterpret_cast<const float4*>(a.bufs[p]) + i0 + u * stride);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i >= hi) break;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (W) {
#pragma unroll
                for (int p = 0; p < (W ? W : 1); ++p) { acc.x += v[u][p].x; acc.y += v[u][p].y; acc.z += v[u][p].z; acc.w += v[u][p].w; }
            } else {
                for (int p = 0; p < world; ++p) {
                    const float4 t = ld_f4_sys(reinterpret_cast<const float4*>(a.bufs[p]) + i);
                    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                }
            }
            acc.x *= a.scale; acc.y *= a.scale; acc.z *= a.scale; acc.w *= a.scale;
#pragma unroll
            for (int p = 0; p < (W ? W : kXMaxWorld); ++p)
                if (p < world) reinterpret_cast<float4*>(a.bufs[p])[i] = acc;
        }
    }
    // every block: its pushes are ordered before the ticket; the last block tells the peers and waits for theirs
    __shared__ unsigned int last_s;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last_s = atomicAdd(a.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!last_s) return;
    __threadfence_system();
    if (threadIdx.x < world && (int)threadIdx.x != a.rank) {
        st_flag_sys(a.flags[threadIdx.x] + kXMaxWorld + a.rank, e);
        wait_flag(my_flags + kXMaxWorld + threadIdx.x, e, a.timeout_ns, a.rank, threadIdx.x, "pushes landed");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *a.ticket = 0u;
        *a.epoch_ctr = e;
    }
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

extern "C" __attribute__((visibility("default"))) long long fv_grad_allreduce_flag_words(void) { return 2 * fv::kXMaxWorld; }

// In-place mean (scale = 1 / world) all-reduce of the symmetric flat gradient buffers: bufs_dev / flags_dev are device arrays of
// `world` peer-mapped pointers (buffer of n floats, n % 4 == 0, 16-byte aligned; flag array of fv_grad_allreduce_flag_words()
// zero-initialised 64-bit words), epoch_ctr (u64) and ticket (u32) zero-initialised device words of this rank.
extern "C" __attribute__((visibility("default"))) int fv_grad_allreduce(void* bufs_dev, void* flags_dev, int rank, int world, long long n, void* epoch_ctr,
                                                                      void* ticket, float scale, void* stream) {
    using namespace fv;
    if (!bufs_dev || !flags_dev || !epoch_ctr || !ticket || n < 4 || n % 4) return fail(FV_ERR_ARG, "fv_grad_allreduce: bad arguments (n=%lld)", n);
    if (world < 1 || world > kXMaxWorld || rank < 0 || rank >= world) return fail(FV_ERR_ARG, "fv_grad_allreduce: rank %d / world %d", rank, world);
    GradArArgs a{reinterpret_cast<float* const*>(bufs_dev), reinterpret_cast<unsigned long long* const*>(flags_dev), rank, world, n / 4,
                 reinterpret_cast<unsigned long long*>(epoch_ctr), reinterpret_cast<unsigned int*>(ticket), scale, xrank_timeout_ns()};
    const long long slice = (a.n4 + world - 1) / world;
    cudaStream_t st = (cudaStream_t)stream;
    auto blocks = [&](int u) {
        long long b = (slice + 256LL * u - 1) / (256LL * u);
        const int cap = num_sms();
        return (int)(b < 1 ? 1 : (b > cap ? cap : b));
    };
    switch (world) {
        case 2: grad_allreduce_kernel<2, 8><<<blocks(8), 256, 0, st>>>(a); break;
        case 4: grad_allreduce_kernel<4, 4><<<blocks(4), 256, 0, st>>>(a); break;
        case 8: grad_allreduce_kernel<8, 2><<<blocks(2), 256, 0, st>>>(a); break;
        default: grad_allreduce_kernel<0, 2><<<blocks(2), 256, 0, st>>>(a); break;
    }
    FV_LAUNCH_CHECK("grad_allreduce_kernel");
    return FV_OK;
}
