// First layer of the encoder: SameBlock2D(3 -> 32) = 1x1 conv + training-mode SyncBatchNorm + ReLU on the raw NCHW fp32
// frames (reference modules.py:97-108 as used by EFE_conv5.down, models.py:749).
//
// With 3 input channels this layer carries 0.07 % of the FLOPs but, done the generic way (layout pass, GEMM with K
// padded to 16, statistics pass, norm+act pass; reduce + apply + wgrad in backward), it moves a 134 MB tensor seven
// times.  A 1x1 conv followed by batch norm is a per-pixel affine map whose batch statistics follow from the first
// and second moments of the INPUT:  mean_y = W mean_x + b,  var_y = w^T Cov_x w.  So
//   forward  = moments of x (9 numbers, one pass over 25 MB) + one pass  a = relu(A x + c)  writing NHWC bf16;
//   backward = ONE pass over (g, x) accumulating sum(dz) and sum(dz * x) per output channel (128 numbers); dW, dgamma,
//              dbeta follow in closed form (the BN coupling terms need only those sums and the input moments).
// All sums are accumulated in double, combined across blocks in a fixed order (fv_reduce.cuh: bitwise reproducible) and
// all-reduced by the host across data-parallel ranks like every other batch-norm statistic.
#include <cstdlib>

#include "../../include/facevae_b200.h"
#include "fv_host.h"
#include "fv_ptx.cuh"
#include "fv_reduce.cuh"

namespace fv {

static constexpr int kPwThreads = 256;
static constexpr int kPwMaxC = 4;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sums[0..C) = sum_p x_c ; sums[C + c*C + d] = sum_p x_c x_d   (written, not accumulated)
template <int C>
__global__ void pw_moments_kernel(const float* __restrict__ x, double* __restrict__ sums, int N, int HW, void* ws) {
    __shared__ double red[kPwThreads / 32][C + C * C];
    __shared__ double blk[C + C * C], tot[C + C * C];
    __shared__ int red_flag;
    float s[C], q[C][C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        s[c] = 0.f;
#pragma unroll
        for (int d = 0; d < C; ++d) q[c][d] = 0.f;
    }
    const long long P = (long long)N * HW;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(p / HW), hw = (int)(p % HW);
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldg(x + ((long long)n * C + c) * HW + hw);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            s[c] += v[c];
#pragma unroll
            for (int d = c; d < C; ++d) q[c][d] = fmaf(v[c], v[d], q[c][d]);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const double a = warp_sum_d((double)s[c]);
        if (lane == 0) red[warp][c] = a;
#pragma unroll
        for (int d = 0; d < C; ++d) {
            const double b = warp_sum_d((double)(d >= c ? q[c][d] : q[d][c]));
            if (lane == 0) red[warp][C + c * C + d] = b;
        }
    }
    __syncthreads();
    if (threadIdx.x < C + C * C) {
        double a = 0;
        for (int w = 0; w < kPwThreads / 32; ++w) a += red[w][threadIdx.x];
        blk[threadIdx.x] = a;
    }
    __syncthreads();
    if (det_reduce<double>(ws, C + C * C, gridDim.x, blockIdx.x, blk, tot, threadIdx.x, blockDim.x, BlockSync{}, &red_flag))
        if (threadIdx.x < C + C * C) sums[threadIdx.x] = tot[threadIdx.x];
}

// coef[co][0..C) = A = gamma*invstd*W, coef[co][C] = c = beta - gamma*invstd*(W mean_x); stat[0][co] = mean_y, stat[1][co] = invstd
__global__ void pw_prepare_kernel(const double* __restrict__ sums, double count, const float* __restrict__ w, const float* __restrict__ bias,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean, float* running_var,
                                  float momentum, float eps, float* __restrict__ coef, float* __restrict__ stat, int Co, int C) {
    const int co = blockIdx.x * blockDim.x + threadIdx.x;
    if (co >= Co) return;
    double mx[kPwMaxC], wm = 0, var = 0;
    for (int c = 0; c < C; ++c) mx[c] = sums[c] / count;
    for (int c = 0; c < C; ++c) {
        wm += (double)w[co * C + c] * mx[c];
        for (int d = 0; d < C; ++d)
            var += (double)w[co * C + c] * (double)w[co * C + d] * (sums[C + c * C + d] / count - mx[c] * mx[d]);
    }
    if (var < 0) var = 0;
    const double b = bias ? (double)bias[co] : 0.0;
    const double mean_y = wm + b;
    const double invstd = 1.0 / sqrt(var + (double)eps);
    const double gi = (double)gamma[co] * invstd;
    for (int c = 0; c < C; ++c) coef[co * (C + 1) + c] = (float)(gi * (double)w[co * C + c]);
    coef[co * (C + 1) + C] = (float)((double)beta[co] - gi * wm);
    stat[co] = (float)mean_y;
    stat[Co + co] = (float)invstd;
    if (running_mean) {
        const double unbiased = count > 1 ? var * count / (count - 1) : var;
        running_mean[co] = (1.f - momentum) * running_mean[co] + momentum * (float)mean_y;
        running_var[co] = (1.f - momentum) * running_var[co] + momentum * (float)unbiased;
    }
}

// a[p, co] = act(A[co] . x[p] + c[co]) -> NHWC bf16; one thread per pixel, consecutive threads = consecutive pixels
template <int C, int CO>
__global__ void pw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ coef, __nv_bfloat16* __restrict__ out, int N, int HW,
                              int act) {
    __shared__ float cs[CO * (C + 1)];
    for (int i = threadIdx.x; i < CO * (C + 1); i += blockDim.x) cs[i] = coef[i];
    __syncthreads();
    const long long P = (long long)N * HW;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(p / HW), hw = (int)(p % HW);
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldg(x + ((long long)n * C + c) * HW + hw);
        uint4* o = reinterpret_cast<uint4*>(out + p * CO);
#pragma unroll
        for (int g = 0; g < CO / 8; ++g) {
            float r[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float* cc = cs + (g * 8 + k) * (C + 1);
                float z = cc[C];
#pragma unroll
                for (int c = 0; c < C; ++c) z = fmaf(cc[c], v[c], z);
                r[k] = act == FV_ACT_RELU ? fmaxf(z, 0.f) : (act == FV_ACT_LEAKY ? (z > 0.f ? z : 0.2f * z) : z);
            }
            o[g] = make_uint4(pack_bf16(r[0], r[1]), pack_bf16(r[2], r[3]), pack_bf16(r[4], r[5]), pack_bf16(r[6], r[7]));
        }
    }
}

// sums[co] = sum_p dz[p,co];  sums[CO + co*C + c] = sum_p dz[p,co] * x[p,c];   dz = g * act'(A x + c).
// Two threads per pixel (16 output channels each) keep the partial sums at 64 registers so two blocks fit per SM.
template <int C, int CO>
__global__ void __launch_bounds__(kPwThreads, 2)
pw_bwd_reduce_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ g, const float* __restrict__ coef,
                     double* __restrict__ sums, int N, int HW, int act, void* ws) {
    constexpr int CH = CO / 2;                                   // channels per thread
    __shared__ float cs[CO * (C + 1)];
    __shared__ float red[kPwThreads / 32][2][CH * (C + 1)];
    __shared__ double blk[CO * (C + 1)], tot[CO * (C + 1)];
    __shared__ int red_flag;
    for (int i = threadIdx.x; i < CO * (C + 1); i += blockDim.x) cs[i] = coef[i];
    __syncthreads();
    const int half = threadIdx.x & 1;
    float acc[CH][C + 1];
#pragma unroll
    for (int co = 0; co < CH; ++co)
#pragma unroll
        for (int c = 0; c <= C; ++c) acc[co][c] = 0.f;
    const long long P = (long long)N * HW;
    const long long stride = (long long)gridDim.x * (blockDim.x / 2);
    // two pixels per thread and iteration, all loads issued before the arithmetic (one pixel in flight per thread left the
    // kernel at 2 TB/s: 44 bytes per thread cannot cover the HBM latency)
    constexpr int UP = 2;
    for (long long p0 = blockIdx.x * (long long)(blockDim.x / 2) + (threadIdx.x >> 1); p0 < P; p0 += UP * stride) {
        float v[UP][C];
        uint4 raw[UP][CH / 8];
        bool ok[UP];
#pragma unroll
        for (int u = 0; u < UP; ++u) {
            const long long p = p0 + u * stride;
            ok[u] = p < P;
            const long long pc = ok[u] ? p : p0;
            const int n = (int)(pc / HW), hw = (int)(pc % HW);
#pragma unroll
            for (int c = 0; c < C; ++c) v[u][c] = __ldg(x + ((long long)n * C + c) * HW + hw);
            const uint4* gp = reinterpret_cast<const uint4*>(g + pc * CO + half * CH);
#pragma unroll
            for (int grp = 0; grp < CH / 8; ++grp) raw[u][grp] = ok[u] ? __ldg(gp + grp) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < UP; ++u) {
#pragma unroll
            for (int grp = 0; grp < CH / 8; ++grp) {
                const float gg[8] = {bf16_lo(raw[u][grp].x), bf16_hi(raw[u][grp].x), bf16_lo(raw[u][grp].y), bf16_hi(raw[u][grp].y),
                                     bf16_lo(raw[u][grp].z), bf16_hi(raw[u][grp].z), bf16_lo(raw[u][grp].w), bf16_hi(raw[u][grp].w)};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int col = grp * 8 + k;
                    const float* cc = cs + (half * CH + col) * (C + 1);
                    float z = cc[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) z = fmaf(cc[c], v[u][c], z);
                    const float d = act == FV_ACT_RELU ? (z > 0.f ? 1.f : 0.f) : (act == FV_ACT_LEAKY ? (z > 0.f ? 1.f : 0.2f) : 1.f);
                    const float dz = gg[k] * d;               // a pixel past the end carries g = 0
                    acc[col][C] += dz;
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[col][c] = fmaf(dz, v[u][c], acc[col][c]);
                }
            }
        }
    }
    // lanes of equal parity hold the same channel half: reduce over the 16 lanes of each parity
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int co = 0; co < CH; ++co)
#pragma unroll
        for (int c = 0; c <= C; ++c) {
            float a = acc[co][c];
#pragma unroll
            for (int o = 16; o > 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane < 2) red[warp][lane][co * (C + 1) + c] = a;
        }
    __syncthreads();
    for (int i = threadIdx.x; i < CO * (C + 1); i += blockDim.x) {
        const int co = i / (C + 1), c = i % (C + 1);
        const int h = co / CH, col = co % CH;
        double a = 0;
        for (int w = 0; w < kPwThreads / 32; ++w) a += (double)red[w][h][col * (C + 1) + c];
        blk[c == C ? co : CO + co * C + c] = a;
    }
    __syncthreads();
    if (det_reduce<double>(ws, CO * (C + 1), gridDim.x, blockIdx.x, blk, tot, threadIdx.x, blockDim.x, BlockSync{}, &red_flag))
        for (int i = threadIdx.x; i < CO * (C + 1); i += blockDim.x) sums[i] = tot[i];
}

// The same sums with the loads decoupled from the arithmetic: the kernel above issues a tile's loads, waits a full HBM latency
// (~2 us at 16 warps per SM) and only then computes -- 28 such rounds per SM = the whole 78 us it took (ncu: DRAM 25 %, every
// warp parked on its first use of g).  Here one thread streams 256-pixel tiles (the contiguous 16 KB of g and the C 1-KB rows of
// x) into a 4-stage shared-memory ring with cp.async.bulk (TMA, completion on an mbarrier) three tiles ahead of the arithmetic.
// Requires HW % 256 == 0 (a tile stays inside one image plane) and 16-byte aligned tensors.
static constexpr int kPwTile = 256, kPwStages = 4;

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int C, int CO>
__global__ void __launch_bounds__(kPwThreads, 2)
pw_bwd_reduce_pipe_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ g, const float* __restrict__ coef,
                          double* __restrict__ sums, int N, int HW, int act, void* ws) {
    constexpr int CH = CO / 2;
    constexpr int G_BYTES = kPwTile * CO * 2, X_BYTES = kPwTile * 4, STAGE = G_BYTES + C * X_BYTES;
    extern __shared__ __align__(128) uint8_t pw_smem[];
    __shared__ float cs[CO * (C + 1)];
    __shared__ float red[kPwThreads / 32][2][CH * (C + 1)];
    __shared__ double blk[CO * (C + 1)], tot[CO * (C + 1)];
    __shared__ int red_flag;
    __shared__ __align__(8) uint64_t full[kPwStages];
    for (int i = threadIdx.x; i < CO * (C + 1); i += blockDim.x) cs[i] = coef[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kPwStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int tiles_per_img = HW / kPwTile;
    const long long num_tiles = (long long)N * tiles_per_img;
    auto issue = [&](long long tile, int slot) {          // thread 0 only
        const int n = (int)(tile / tiles_per_img), hw0 = (int)(tile % tiles_per_img) * kPwTile;
        uint8_t* dst = pw_smem + (size_t)slot * STAGE;
        mbar_arrive_expect_tx(&full[slot], (uint32_t)STAGE);
        bulk_load(dst, g + ((long long)n * HW + hw0) * CO, G_BYTES, &full[slot]);
#pragma unroll
        for (int c = 0; c < C; ++c) bulk_load(dst + G_BYTES + c * X_BYTES, x + ((long long)n * C + c) * HW + hw0, X_BYTES, &full[slot]);
    };
    if (threadIdx.x == 0)
        for (int k = 0; k < kPwStages - 1; ++k)
            if (blockIdx.x + (long long)k * gridDim.x < num_tiles) issue(blockIdx.x + (long long)k * gridDim.x, k);
    const int half = threadIdx.x & 1, px = threadIdx.x >> 1;
    float acc[CH][C + 1];
#pragma unroll
    for (int co = 0; co < CH; ++co)
#pragma unroll
        for (int c = 0; c <= C; ++c) acc[co][c] = 0.f;
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int slot = it % kPwStages;
        // the slot refilled here was read in the previous iteration: the barrier at its end ordered those reads before this copy
        if (threadIdx.x == 0) {
            const long long nxt = tile + (long long)(kPwStages - 1) * gridDim.x;
            if (nxt < num_tiles) issue(nxt, (it + kPwStages - 1) % kPwStages);
        }
        mbar_wait(&full[slot], (it / kPwStages) & 1);
        const uint8_t* st = pw_smem + (size_t)slot * STAGE;
        const float* xs = reinterpret_cast<const float*>(st + G_BYTES);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int p = px + u * (kPwThreads / 2);
            float v[C];
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = xs[c * kPwTile + p];
            const uint4* gp = reinterpret_cast<const uint4*>(st + (size_t)p * CO * 2 + half * CH * 2);
#pragma unroll
            for (int grp = 0; grp < CH / 8; ++grp) {
                const uint4 raw = gp[grp];
                const float gg[8] = {bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y), bf16_lo(raw.z), bf16_hi(raw.z), bf16_lo(raw.w), bf16_hi(raw.w)};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int col = grp * 8 + k;
                    const float* cc = cs + (half * CH + col) * (C + 1);
                    float z = cc[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) z = fmaf(cc[c], v[c], z);
                    const float d = act == FV_ACT_RELU ? (z > 0.f ? 1.f : 0.f) : (act == FV_ACT_LEAKY ? (z > 0.f ? 1.f : 0.2f) : 1.f);
                    const float dz = gg[k] * d;
                    acc[col][C] += dz;
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[col][c] = fmaf(dz, v[c], acc[col][c]);
                }
            }
        }
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int co = 0; co < CH; ++co)
#pragma unroll
        for (int c = 0; c <= C; ++c) {
            float a = acc[co][c];
#pragma unroll
            for (int o = 16; o > 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane < 2) red[warp][lane][co * (C + 1) + c] = a;
        }
    __syncthreads();
    for (int i = threadIdx.x; i < CO * (C + 1); i += blockDim.x) {
        const int co = i / (C + 1), c = i % (C + 1);
        const int h = co / CH, col = co % CH;
        double a = 0;
        for (int w = 0; w < kPwThreads / 32; ++w) a += (double)red[w][h][col * (C + 1) + c];
        blk[c == C ? co : CO + co * C + c] = a;
    }
    __syncthreads();
    if (det_reduce<double>(ws, CO * (C + 1), gridDim.x, blockIdx.x, blk, tot, threadIdx.x, blockDim.x, BlockSync{}, &red_flag))
        for (int i = threadIdx.x; i < CO * (C + 1); i += blockDim.x) sums[i] = tot[i];
}

// closed-form parameter gradients from the forward moments (fs) and the backward sums (bs); see the header comment
__global__ void pw_bwd_finalize_kernel(const double* __restrict__ fs, const double* __restrict__ bs, double count, const float* __restrict__ w,
                                       const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ stat,
                                       float* __restrict__ dw, float* __restrict__ dgamma, float* __restrict__ dbeta, int Co, int C) {
    const int co = blockIdx.x * blockDim.x + threadIdx.x;
    if (co >= Co) return;
    const double mean_y = stat[co], invstd = stat[Co + co], b = bias ? (double)bias[co] : 0.0;
    const double s_dz = bs[co];
    double s_dzy = b * s_dz;
    for (int c = 0; c < C; ++c) s_dzy += (double)w[co * C + c] * bs[Co + co * C + c];
    const double dgam = invstd * (s_dzy - mean_y * s_dz);
    dgamma[co] = (float)dgam;
    dbeta[co] = (float)s_dz;
    const double c1 = s_dz / count, c2 = dgam / count, gi = (double)gamma[co] * invstd;
    for (int ci = 0; ci < C; ++ci) {
        double s_yx = b * fs[ci];                                   // sum_p y[co] x[ci]
        for (int cj = 0; cj < C; ++cj) s_yx += (double)w[co * C + cj] * fs[C + cj * C + ci];
        const double s_xhat_x = invstd * (s_yx - mean_y * fs[ci]);  // sum_p xhat[co] x[ci]
        dw[co * C + ci] = (float)(gi * (bs[Co + co * C + ci] - c1 * fs[ci] - c2 * s_xhat_x));
    }
}

static inline int pw_grid(long long pixels) {
    long long b = (pixels + kPwThreads - 1) / kPwThreads;
    const long long cap = (long long)num_sms() * 8;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace fv

using namespace fv;
#define STREAM ((cudaStream_t)stream)
#define PW_DISPATCH_C(C_, EXPR3, EXPR4, EXPR1, EXPR2) \
    switch (C_) { case 1: EXPR1; break; case 2: EXPR2; break; case 3: EXPR3; break; default: EXPR4; break; }

extern "C" __attribute__((visibility("default"))) int fv_pw_moments(const float* x, double* sums, int N, int C, int HW, void* ws, void* stream) {
    if (!x || !sums || !ws || C < 1 || C > kPwMaxC) return fail(FV_ERR_ARG, "fv_pw_moments: C=%d must be 1..%d", C, kPwMaxC);
    const int grid = pw_grid((long long)N * HW);
    PW_DISPATCH_C(C, (pw_moments_kernel<3><<<grid, kPwThreads, 0, STREAM>>>(x, sums, N, HW, ws)),
                  (pw_moments_kernel<4><<<grid, kPwThreads, 0, STREAM>>>(x, sums, N, HW, ws)),
                  (pw_moments_kernel<1><<<grid, kPwThreads, 0, STREAM>>>(x, sums, N, HW, ws)),
                  (pw_moments_kernel<2><<<grid, kPwThreads, 0, STREAM>>>(x, sums, N, HW, ws)));
    FV_LAUNCH_CHECK("pw_moments_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_pw_prepare(const double* sums, double count, const float* w, const float* bias,
                                                                  const float* gamma, const float* beta, float* running_mean,
                                                                  float* running_var, float momentum, float eps, float* coef, float* stat,
                                                                  int Co, int C, void* stream) {
    if (!sums || !w || !gamma || !beta || !coef || !stat || count <= 0 || C < 1 || C > kPwMaxC) return fail(FV_ERR_ARG, "fv_pw_prepare: bad arguments");
    pw_prepare_kernel<<<(Co + 63) / 64, 64, 0, STREAM>>>(sums, count, w, bias, gamma, beta, running_mean, running_var, momentum, eps, coef, stat, Co, C);
    FV_LAUNCH_CHECK("pw_prepare_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_pw_fwd(const float* x, const float* coef, void* out, int N, int C, int HW, int Co, int act,
                                                              void* stream) {
    if (!x || !coef || !out) return fail(FV_ERR_ARG, "fv_pw_fwd: null pointer");
    if (Co != 32 || C < 1 || C > kPwMaxC) return fail(FV_ERR_UNSUPPORTED, "fv_pw_fwd: supports C in 1..4 and Co = 32 (got %d -> %d)", C, Co);
    const int grid = pw_grid((long long)N * HW);
    __nv_bfloat16* o = (__nv_bfloat16*)out;
    PW_DISPATCH_C(C, (pw_fwd_kernel<3, 32><<<grid, kPwThreads, 0, STREAM>>>(x, coef, o, N, HW, act)),
                  (pw_fwd_kernel<4, 32><<<grid, kPwThreads, 0, STREAM>>>(x, coef, o, N, HW, act)),
                  (pw_fwd_kernel<1, 32><<<grid, kPwThreads, 0, STREAM>>>(x, coef, o, N, HW, act)),
                  (pw_fwd_kernel<2, 32><<<grid, kPwThreads, 0, STREAM>>>(x, coef, o, N, HW, act)));
    FV_LAUNCH_CHECK("pw_fwd_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_pw_bwd_reduce(const float* x, const void* g, const float* coef, double* sums, int N, int C,
                                                                     int HW, int Co, int act, void* ws, void* stream) {
    if (!x || !g || !coef || !sums || !ws) return fail(FV_ERR_ARG, "fv_pw_bwd_reduce: null pointer");
    if (Co != 32 || C < 1 || C > kPwMaxC) return fail(FV_ERR_UNSUPPORTED, "fv_pw_bwd_reduce: supports C in 1..4 and Co = 32 (got %d -> %d)", C, Co);
    const int grid = pw_grid((long long)N * HW);
    const __nv_bfloat16* gp = (const __nv_bfloat16*)g;
    const char* env = getenv("FV_PW_PIPE");
    if (C == 3 && HW % kPwTile == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g)) & 15) == 0 && !(env && atoi(env) == 0)) {
        constexpr size_t smem = (size_t)kPwStages * (kPwTile * 32 * 2 + 3 * kPwTile * 4);
        static bool attr_set = false;
        if (!attr_set) {
            FV_CUDA(cudaFuncSetAttribute(pw_bwd_reduce_pipe_kernel<3, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const long long tiles = (long long)N * (HW / kPwTile);
        const int pgrid = (int)(tiles < 2LL * num_sms() ? tiles : 2LL * num_sms());
        pw_bwd_reduce_pipe_kernel<3, 32><<<pgrid, kPwThreads, smem, STREAM>>>(x, gp, coef, sums, N, HW, act, ws);
        FV_LAUNCH_CHECK("pw_bwd_reduce_pipe_kernel");
        return FV_OK;
    }
    PW_DISPATCH_C(C, (pw_bwd_reduce_kernel<3, 32><<<grid, kPwThreads, 0, STREAM>>>(x, gp, coef, sums, N, HW, act, ws)),
                  (pw_bwd_reduce_kernel<4, 32><<<grid, kPwThreads, 0, STREAM>>>(x, gp, coef, sums, N, HW, act, ws)),
                  (pw_bwd_reduce_kernel<1, 32><<<grid, kPwThreads, 0, STREAM>>>(x, gp, coef, sums, N, HW, act, ws)),
                  (pw_bwd_reduce_kernel<2, 32><<<grid, kPwThreads, 0, STREAM>>>(x, gp, coef, sums, N, HW, act, ws)));
    FV_LAUNCH_CHECK("pw_bwd_reduce_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_pw_bwd_finalize(const double* fsums, const double* bsums, double count, const float* w,
                                                                       const float* bias, const float* gamma, const float* stat, float* dw,
                                                                       float* dgamma, float* dbeta, int Co, int C, void* stream) {
    if (!fsums || !bsums || !w || !gamma || !stat || !dw || !dgamma || !dbeta || count <= 0) return fail(FV_ERR_ARG, "fv_pw_bwd_finalize: bad arguments");
    pw_bwd_finalize_kernel<<<(Co + 63) / 64, 64, 0, STREAM>>>(fsums, bsums, count, w, bias, gamma, stat, dw, dgamma, dbeta, Co, C);
    FV_LAUNCH_CHECK("pw_bwd_finalize_kernel");
    return FV_OK;
}
