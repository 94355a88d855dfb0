// Memory-bound glue of the face-vae hot path: layout conversion, filter preparation, batch-norm statistics,
// fused norm + activation (+ 2x2 average pool / nearest up-sample) forward and backward, the re-parameterisation
// fused with the KL reduction, and the reconstruction loss fused with its gradient.  Every kernel is a single
// coalesced, 16-byte-vectorised pass (8 channels per thread in NHWC), sized in multiples of the SM count.
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "../../include/facevae_b200.h"
#include "fv_host.h"
#include "fv_ptx.cuh"
#include "fv_reduce.cuh"
#include "fv_xrank.cuh"

namespace fv {

static constexpr int kThreads = 256;

// Grids are ONE resident wave (num_sms x blocks that fit on an SM at the kernel's register count) of grid-stride blocks:
// a block's life is a few dependent latency rounds (per-channel constants, the loads, the block reduction + atomics), so
// several short waves multiply that fixed cost -- 18 us for a 4 MB tensor at 8 blocks per SM in the round-1 profile.
static inline int grid_for(long long work_items, int per_block = kThreads, int waves = 8) {
    long long b = (work_items + per_block - 1) / per_block;
    long long cap = (long long)num_sms() * (b > (long long)num_sms() * 64 ? 8 : waves);
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---------------------------------------------------------------- 8-wide load / store helpers
template <typename T> struct V8;
template <> struct V8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
        f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
        *reinterpret_cast<uint4*>(p) =
            make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
    }
};
template <> struct V8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
};

// Division by an image extent / channel-group count that is almost always a power of two: the generic 32-bit division costs ~20
// instructions, twice per row in the index decomposition of the pooled kernels -- as much as the arithmetic on the row's eight
// channels (ncu: 114 instructions per 8-channel row in the pooled backward reduction, issue-bound at 88 us for 336 MB).
struct FastDiv {
    unsigned d;
    int shift;                                   // log2(d) when d is a power of two, -1 otherwise
    __device__ __forceinline__ explicit FastDiv(unsigned d_) : d(d_), shift(-1) {
        if (d_ && (d_ & (d_ - 1)) == 0) {
            shift = 0;
            while ((1u << shift) < d_) ++shift;
        }
    }
    __device__ __forceinline__ void divmod(unsigned x, unsigned& q, unsigned& r) const {
        if (shift >= 0) { q = x >> shift; r = x & (d - 1); }
        else { q = x / d; r = x - q * d; }
    }
};

template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
    uint4 v;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void zero() { v = make_uint4(0u, 0u, 0u, 0u); }
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
        f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
    }
};
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = *reinterpret_cast<const float4*>(p);
        b = *reinterpret_cast<const float4*>(p + 4);
    }
    __device__ __forceinline__ void zero() { a = make_float4(0.f, 0.f, 0.f, 0.f); b = a; }
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};

__device__ __forceinline__ float act_fwd(float z, int act) {
    return act == FV_ACT_RELU ? fmaxf(z, 0.f) : (act == FV_ACT_LEAKY ? (z > 0.f ? z : 0.2f * z) : z);
}
__device__ __forceinline__ float act_grad(float z, int act) {
    return act == FV_ACT_RELU ? (z > 0.f ? 1.f : 0.f) : (act == FV_ACT_LEAKY ? (z > 0.f ? 1.f : 0.2f) : 1.f);
}

// ---------------------------------------------------------------- layout
// NCHW fp32 -> NHWC (C padded to Cp with zeros).  One thread per (pixel, 8-channel group); consecutive threads walk
// consecutive pixels so the per-plane reads coalesce.
template <typename TO>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, TO* __restrict__ dst, int N, int C, int HW, int Cp) {
    const int groups = Cp / 8;
    const long long total = (long long)N * HW * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i % ((long long)N * HW);
        const int g = (int)(i / ((long long)N * HW));
        const int n = (int)(pix / HW), hw = (int)(pix % HW);
        float f[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = g * 8 + k;
            f[k] = c < C ? __ldg(src + ((long long)n * C + c) * HW + hw) : 0.f;
        }
        V8<TO>::store(dst + pix * Cp + g * 8, f);
    }
}

// NHWC (channel stride Cs) -> NCHW fp32 (first C channels); optionally accumulates into dst.
template <typename TI>
__global__ void nhwc_to_nchw_kernel(const TI* __restrict__ src, float* __restrict__ dst, int N, int C, int HW, int Cs,
                                    int accumulate) {
    const int groups = (C + 7) / 8;
    const long long total = (long long)N * HW * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i % ((long long)N * HW);
        const int g = (int)(i / ((long long)N * HW));
        const int n = (int)(pix / HW), hw = (int)(pix % HW);
        float f[8];
        V8<TI>::load(src + pix * Cs + g * 8, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = g * 8 + k;
            if (c < C) {
                float* o = dst + ((long long)n * C + c) * HW + hw;
                *o = accumulate ? *o + f[k] : f[k];
            }
        }
    }
}

// ---------------------------------------------------------------- filters
// w fp32 [Co][Ci][R][S] (nn.Conv2d layout) -> wf bf16 [Co_pad][taps][Ci_pad] (fprop, K-major) and
// wd bf16 [Ci_pad][taps][Co_pad] with the taps rotated by 180 degrees (data gradient as a forward conv of dY).
__global__ void weight_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd,
                                   int Co, int Ci, int R, int S, int Co_pad, int Ci_pad) {
    const int taps = R * S;
    const long long nf = (long long)Co_pad * taps * Ci_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nf; i += (long long)gridDim.x * blockDim.x) {
        if (wf) {
            const int ci = (int)(i % Ci_pad);
            const int tap = (int)((i / Ci_pad) % taps);
            const int co = (int)(i / ((long long)Ci_pad * taps));
            const float v = (co < Co && ci < Ci) ? w[((long long)co * Ci + ci) * taps + tap] : 0.f;
            wf[i] = __float2bfloat16(v);
        }
        if (wd) {
            const int co = (int)(i % Co_pad);
            const int tap = (int)((i / Co_pad) % taps);
            const int ci = (int)(i / ((long long)Co_pad * taps));
            const float v = (co < Co && ci < Ci) ? w[((long long)co * Ci + ci) * taps + (taps - 1 - tap)] : 0.f;
            wd[i] = __float2bfloat16(v);
        }
    }
}

// the same for a table of layers (blockIdx.y = layer): one launch per step instead of one per convolution.  kind 0: plain
// convolution (wf | wd), 1: up-sampling 3x3 convolution (wx2 | ws2, weight_prep_up_kernel), 2: 4x4 stride-2 convolution
// (wf | wx2, weight_prep_s2_kernel).
__device__ __forceinline__ void up_rows(int a, int u, int& r0, int& r1);
__device__ __forceinline__ float up_tap_sum(const float* __restrict__ w9, int a, int u, int b, int v) {
    int r0, r1, s0, s1;
    up_rows(a, u, r0, r1);
    up_rows(b, v, s0, s1);
    float acc = 0.f;
    for (int r = r0; r <= r1; ++r)
        for (int s_ = s0; s_ <= s1; ++s_) acc += __ldg(w9 + r * 3 + s_);
    return acc;
}
// one item (index i of the padded operand layouts) of one layer of the batched filter preparation
__device__ __forceinline__ void prep_item(const fv_prep_desc& d, unsigned i) {
    const float* __restrict__ w = d.w;
    __nv_bfloat16* o0 = static_cast<__nv_bfloat16*>(d.wf);
    __nv_bfloat16* o1 = static_cast<__nv_bfloat16*>(d.wd);
    const unsigned taps = d.kind == 0 ? d.R * d.S : 16, cip = d.Ci_pad, cop = d.Co_pad;
    {
        if (d.kind == 0) {
            if (o0) {
                const unsigned ci = i % cip, t2 = i / cip;
                const unsigned tap = t2 % taps, co = t2 / taps;
                const float v = (co < (unsigned)d.Co && ci < (unsigned)d.Ci) ? __ldg(w + (co * d.Ci + ci) * taps + tap) : 0.f;
                o0[i] = __float2bfloat16(v);
            }
            if (o1) {
                const unsigned co = i % cop, t2 = i / cop;
                const unsigned tap = t2 % taps, ci = t2 / taps;
                const float v = (co < (unsigned)d.Co && ci < (unsigned)d.Ci) ? __ldg(w + (co * d.Ci + ci) * taps + (taps - 1 - tap)) : 0.f;
                o1[i] = __float2bfloat16(v);
            }
        } else if (d.kind == 1) {
            if (o0) {                                      // wx2 [4 phases][Co_pad][4 taps][Ci_pad]
                const unsigned ci = i % cip, tap = (i / cip) % 4, co = (i / (4 * cip)) % cop, ph = i / (4 * cip * cop);
                const float v = (co < (unsigned)d.Co && ci < (unsigned)d.Ci) ? up_tap_sum(w + (co * d.Ci + ci) * 9, ph >> 1, tap >> 1, ph & 1, tap & 1) : 0.f;
                o0[i] = __float2bfloat16(v);
            }
            if (o1) {                                      // ws2 [Ci_pad][16 taps][Co_pad]
                const unsigned co = i % cop, t16 = (i / cop) % 16, ci = i / (16 * cop);
                const unsigned r4 = t16 >> 2, s4 = t16 & 3;
                const float v = (co < (unsigned)d.Co && ci < (unsigned)d.Ci)
                                    ? up_tap_sum(w + (co * d.Ci + ci) * 9, (r4 & 1) ? 0 : 1, r4 < 2 ? 1 : 0, (s4 & 1) ? 0 : 1, s4 < 2 ? 1 : 0) : 0.f;
                o1[i] = __float2bfloat16(v);
            }
        } else {
            if (o0) {                                      // wf [Co_pad][16][Ci_pad]
                const unsigned ci = i % cip, tap = (i / cip) % 16, co = i / (16 * cip);
                o0[i] = __float2bfloat16((co < (unsigned)d.Co && ci < (unsigned)d.Ci) ? __ldg(w + (co * d.Ci + ci) * 16 + tap) : 0.f);
            }
            if (o1) {                                      // wx2 [4 phases][Ci_pad][4 taps][Co_pad]
                const unsigned co = i % cop, tap = (i / cop) % 4, ci = (i / (4 * cop)) % cip, ph = i / (4 * cop * cip);
                const unsigned r4 = 3 - 2 * (tap >> 1) - (ph >> 1), s4 = 3 - 2 * (tap & 1) - (ph & 1);
                o1[i] = __float2bfloat16((co < (unsigned)d.Co && ci < (unsigned)d.Ci) ? __ldg(w + (co * d.Ci + ci) * 16 + r4 * 4 + s4) : 0.f);
            }
        }
    }
}

__global__ void weight_prep_batched_kernel(const fv_prep_desc* __restrict__ table) {
    const fv_prep_desc d = table[blockIdx.y];
    const unsigned taps = d.kind == 0 ? d.R * d.S : 16;
    const unsigned nf = (unsigned)d.Co_pad * taps * (unsigned)d.Ci_pad;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) prep_item(d, i);
}

// Flat grid: block b works on 1024 consecutive items of the layer whose [reserved, next.reserved) range contains b -- blocks in
// proportion to the layer sizes (they differ 1000x), one item per thread and pass instead of ~30 serial items per thread.
static constexpr int kPrepBlockItems = 1024;
__global__ void __launch_bounds__(256) weight_prep_flat_kernel(const fv_prep_desc* __restrict__ table, int n_layers) {
    __shared__ int layer_s;
    if (threadIdx.x == 0) {
        int l = 0;
        while (l + 1 < n_layers && table[l + 1].reserved <= (int)blockIdx.x) ++l;
        layer_s = l;
    }
    __syncthreads();
    const fv_prep_desc d = table[layer_s];
    const unsigned taps = d.kind == 0 ? d.R * d.S : 16;
    const unsigned nf = (unsigned)d.Co_pad * taps * (unsigned)d.Ci_pad;
    const unsigned base = (blockIdx.x - (unsigned)d.reserved) * kPrepBlockItems;
#pragma unroll
    for (int k = 0; k < kPrepBlockItems / 256; ++k) {
        const unsigned i = base + k * 256 + threadIdx.x;
        if (i < nf) prep_item(d, i);
    }
}

// Tile version: a block stages a [16 output channels][32 input channels][taps] tile of the fp32 filter in shared memory with
// coalesced reads (32 * taps contiguous floats per output channel) and writes both operands from there in THEIR storage order.
// The item-wise kernels read the filter with the stride of whichever operand they are writing: for the data-gradient operand
// (output channel fastest) that is one 32-byte sector per element -- 43 us for the anchor's 3.8 M weights.
static constexpr int kPrepTco = 16, kPrepTci = 32, kPrepMaxTaps = 16;
__device__ __forceinline__ float up_tap_sum_s(const float* w9, int a, int u, int b, int v) {
    int r0, r1, s0, s1;
    up_rows(a, u, r0, r1);
    up_rows(b, v, s0, s1);
    float acc = 0.f;
    for (int r = r0; r <= r1; ++r)
        for (int s_ = s0; s_ <= s1; ++s_) acc += w9[r * 3 + s_];
    return acc;
}
static constexpr int kPrepRow = kPrepTci * (kPrepMaxTaps + 1) + 1;   // + 1: output-channel-fastest reads hit 16 banks, not one
// TAPS > 0: the source filter's tap count as a compile-time constant (the index arithmetic divides by it per element)
template <int TAPS>
__device__ __forceinline__ void prep_tile_body(const fv_prep_desc& d, int local, int taps_rt, float (*ws)[kPrepRow]) {
    const int taps = TAPS ? TAPS : taps_rt;
    const int cip = d.Ci_pad, cop = d.Co_pad;
    const int tiles_ci = (cip + kPrepTci - 1) / kPrepTci;
    const int co0 = (local / tiles_ci) * kPrepTco, ci0 = (local % tiles_ci) * kPrepTci;
    const int ts = taps | 1;                                                     // odd row stride: conflict-free strided reads
    // stage: ws[co_l][ci_l * ts + t] = w[co0 + co_l][ci0 + ci_l][t] (zero outside the real filter)
    const int per_co = kPrepTci * taps;
    // eight loads in flight per thread: one load -> one shared store per iteration left every thread waiting a full memory
    // latency 18 times in a row (ncu: the store's scoreboard stall was the whole kernel)
    constexpr int kBatch = 8;
    for (int e0 = threadIdx.x; e0 < kPrepTco * per_co; e0 += 256 * kBatch) {
        float v[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int e = e0 + k * 256;
            const int co_l = e / per_co, rem = e - co_l * per_co;
            const int ci_l = rem / taps, t = rem - ci_l * taps;
            const int co = co0 + co_l, ci = ci0 + ci_l;
            v[k] = (e < kPrepTco * per_co && co < d.Co && ci < d.Ci) ? __ldg(d.w + ((size_t)co * d.Ci + ci) * taps + t) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int e = e0 + k * 256;
            if (e < kPrepTco * per_co) {
                const int co_l = e / per_co, rem = e - co_l * per_co;
                const int ci_l = rem / taps, t = rem - ci_l * taps;
                ws[co_l][ci_l * ts + t] = v[k];
            }
        }
    }
    __syncthreads();
    __nv_bfloat16* o0 = static_cast<__nv_bfloat16*>(d.wf);
    __nv_bfloat16* o1 = static_cast<__nv_bfloat16*>(d.wd);
    const int nci = min(kPrepTci, cip - ci0);                                    // input channels of this tile inside the padded extent
    if (d.kind == 0) {
        if (o0)                                                                  // wf [Co_pad][taps][Ci_pad]: ci fastest
            for (int e = threadIdx.x; e < kPrepTco * taps * kPrepTci; e += 256) {
                const int ci_l = e % kPrepTci, t = (e / kPrepTci) % taps, co_l = e / (kPrepTci * taps);
                if (ci_l < nci) o0[((size_t)(co0 + co_l) * taps + t) * cip + ci0 + ci_l] = __float2bfloat16(ws[co_l][ci_l * ts + t]);
            }
        if (o1)                                                                  // wd [Ci_pad][taps][Co_pad], taps rotated: co fastest
            for (int e = threadIdx.x; e < kPrepTci * taps * kPrepTco; e += 256) {
                const int co_l = e % kPrepTco, t = (e / kPrepTco) % taps, ci_l = e / (kPrepTco * taps);
                if (ci_l < nci) o1[((size_t)(ci0 + ci_l) * taps + t) * cop + co0 + co_l] = __float2bfloat16(ws[co_l][ci_l * ts + (taps - 1 - t)]);
            }
    } else if (d.kind == 1) {
        if (o0)                                                                  // wx2 [4 phases][Co_pad][4 taps][Ci_pad]
            for (int e = threadIdx.x; e < 4 * kPrepTco * 4 * kPrepTci; e += 256) {
                const int ci_l = e % kPrepTci, tap = (e / kPrepTci) % 4, co_l = (e / (4 * kPrepTci)) % kPrepTco, ph = e / (4 * kPrepTci * kPrepTco);
                if (ci_l < nci)
                    o0[(((size_t)ph * cop + co0 + co_l) * 4 + tap) * cip + ci0 + ci_l] =
                        __float2bfloat16(up_tap_sum_s(&ws[co_l][ci_l * ts], ph >> 1, tap >> 1, ph & 1, tap & 1));
            }
        if (o1)                                                                  // ws2 [Ci_pad][16 taps][Co_pad]
            for (int e = threadIdx.x; e < kPrepTci * 16 * kPrepTco; e += 256) {
                const int co_l = e % kPrepTco, t16 = (e / kPrepTco) % 16, ci_l = e / (16 * kPrepTco);
                const int r4 = t16 >> 2, s4 = t16 & 3;
                if (ci_l < nci)
                    o1[((size_t)(ci0 + ci_l) * 16 + t16) * cop + co0 + co_l] =
                        __float2bfloat16(up_tap_sum_s(&ws[co_l][ci_l * ts], (r4 & 1) ? 0 : 1, r4 < 2 ? 1 : 0, (s4 & 1) ? 0 : 1, s4 < 2 ? 1 : 0));
            }
    } else {
        if (o0)                                                                  // wf [Co_pad][16][Ci_pad]
            for (int e = threadIdx.x; e < kPrepTco * 16 * kPrepTci; e += 256) {
                const int ci_l = e % kPrepTci, t = (e / kPrepTci) % 16, co_l = e / (kPrepTci * 16);
                if (ci_l < nci) o0[((size_t)(co0 + co_l) * 16 + t) * cip + ci0 + ci_l] = __float2bfloat16(ws[co_l][ci_l * ts + t]);
            }
        if (o1)                                                                  // wx2 [4 phases][Ci_pad][4 taps][Co_pad]
            for (int e = threadIdx.x; e < 4 * kPrepTci * 4 * kPrepTco; e += 256) {
                const int co_l = e % kPrepTco, tap = (e / kPrepTco) % 4, ci_l = (e / (4 * kPrepTco)) % kPrepTci, ph = e / (4 * kPrepTco * kPrepTci);
                const int r4 = 3 - 2 * (tap >> 1) - (ph >> 1), s4 = 3 - 2 * (tap & 1) - (ph & 1);
                if (ci_l < nci)
                    o1[(((size_t)ph * cip + ci0 + ci_l) * 4 + tap) * cop + co0 + co_l] = __float2bfloat16(ws[co_l][ci_l * ts + r4 * 4 + s4]);
            }
    }
}

__global__ void __launch_bounds__(256) weight_prep_tile_kernel(const fv_prep_desc* __restrict__ table, int n_layers) {
    __shared__ float ws[kPrepTco][kPrepRow];
    __shared__ int layer_s;
    if (threadIdx.x == 0) {
        int l = 0;
        while (l + 1 < n_layers && table[l + 1].reserved <= (int)blockIdx.x) ++l;
        layer_s = l;
    }
    __syncthreads();
    const fv_prep_desc d = table[layer_s];
    const int taps = d.kind == 0 ? d.R * d.S : (d.kind == 1 ? 9 : 16);          // taps of the SOURCE filter
    const int local = (int)blockIdx.x - d.reserved;
    if (taps > kPrepMaxTaps) {                                                    // 5x5 / 7x7 filters: item-wise
        const unsigned nf = (unsigned)d.Co_pad * (unsigned)taps * (unsigned)d.Ci_pad;
        const unsigned base = (unsigned)local * kPrepBlockItems;
        for (int k = 0; k < kPrepBlockItems / 256; ++k) {
            const unsigned i = base + k * 256 + threadIdx.x;
            if (i < nf) prep_item(d, i);
        }
        return;
    }
    if (taps == 9) prep_tile_body<9>(d, local, taps, ws);
    else if (taps == 16) prep_tile_body<16>(d, local, taps, ws);
    else if (taps == 1) prep_tile_body<1>(d, local, taps, ws);
    else prep_tile_body<0>(d, local, taps, ws);
}

// dW partials fp32 [splits][Co_pad][taps][Ci_pad] (one slab per pixel split of the weight-gradient kernels, plain stores) ->
// grad fp32 [Co][Ci][R][S]: the splits are added in split order (reproducible; replaces red.global.add into one buffer).
__global__ void wgrad_finish_kernel(const float* __restrict__ part, float* __restrict__ grad, int Co, int Ci, int taps,
                                    int Ci_pad, int accumulate, int splits, long long split_stride) {
    // threads walk the SOURCE layout (ci fastest): the `splits` reads per element are coalesced, the single write is strided
    const long long total = (long long)Co * taps * Ci_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Ci_pad);
        if (ci >= Ci) continue;
        const int tap = (int)((i / Ci_pad) % taps);
        const int co = (int)(i / ((long long)Ci_pad * taps));
        const float* src = part + i;
        float v = 0.f;
        int sp = 0;
        for (; sp + 4 <= splits; sp += 4) {                  // four independent loads in flight, added in split order
            const float a0 = __ldg(src + (sp + 0) * split_stride), a1 = __ldg(src + (sp + 1) * split_stride),
                        a2 = __ldg(src + (sp + 2) * split_stride), a3 = __ldg(src + (sp + 3) * split_stride);
            v += a0; v += a1; v += a2; v += a3;
        }
        for (; sp < splits; ++sp) v += __ldg(src + sp * split_stride);
        float* dst = grad + ((long long)co * Ci + ci) * taps + tap;
        *dst = accumulate ? *dst + v : v;
    }
}

// out[i] (+)= sum_s part[s][i] in slab order (per-CTA / per-split partial weight gradients): float4 loads, four slabs in flight
__global__ void slab_sum_kernel(const float* __restrict__ part, int slabs, long long stride, float* __restrict__ out, long long n, int accumulate) {
    const long long n4 = n / 4;
    const bool vec = (stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(part) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (vec) {
        const float4* p4 = reinterpret_cast<const float4*>(part);
        const long long s4 = stride / 4;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int sp = 0;
            for (; sp + 4 <= slabs; sp += 4) {
                const float4 a0 = __ldg(p4 + (sp + 0) * s4 + i), a1 = __ldg(p4 + (sp + 1) * s4 + i), a2 = __ldg(p4 + (sp + 2) * s4 + i),
                             a3 = __ldg(p4 + (sp + 3) * s4 + i);
                v.x += a0.x; v.y += a0.y; v.z += a0.z; v.w += a0.w;
                v.x += a1.x; v.y += a1.y; v.z += a1.z; v.w += a1.w;
                v.x += a2.x; v.y += a2.y; v.z += a2.z; v.w += a2.w;
                v.x += a3.x; v.y += a3.y; v.z += a3.z; v.w += a3.w;
            }
            for (; sp < slabs; ++sp) {
                const float4 a = __ldg(p4 + sp * s4 + i);
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            }
            float4* o = reinterpret_cast<float4*>(out) + i;
            if (accumulate) { const float4 c = *o; v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w; }
            *o = v;
        }
    }
    for (long long i = (vec ? n4 * 4 : 0) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = 0.f;
        for (int sp = 0; sp < slabs; ++sp) v += __ldg(part + sp * stride + i);
        out[i] = accumulate ? out[i] + v : v;
    }
}

// ---- filters of the up-sampling convolution (UpBlock2D: nearest x2 + 3x3, reference modules.py:78-89) ------------------------
// Output pixel (2i + a, 2j + b) reads the coarse pixels (i + u - 1 + a, j + v - 1 + b), u, v in {0, 1}; the 3x3 taps that land
// on coarse tap u of phase a are  a = 0: u = 0 <- {r = 0}, u = 1 <- {1, 2};  a = 1: u = 0 <- {0, 1}, u = 1 <- {2}.
__device__ __forceinline__ void up_rows(int a, int u, int& r0, int& r1) {      // inclusive range of 3x3 rows
    if (a == 0) { r0 = u == 0 ? 0 : 1; r1 = u == 0 ? 0 : 2; }
    else        { r0 = u == 0 ? 0 : 2; r1 = u == 0 ? 1 : 2; }
}
// w fp32 [Co][Ci][3][3] -> wx2 bf16 [4 phases][Co_pad][4 taps (u, v)][Ci_pad] (forward operand of fv_conv2d_x2: sums of taps in
// fp32, rounded once) and ws2 bf16 [Ci_pad][16 taps (r4, s4)][Co_pad] (operand of fv_conv2d_s2 for the data gradient: the
// fine pixel (2i + r4 - 1, 2j + s4 - 1) of dY meets phase a, tap u with r4 = 3 - 2u - a).
__global__ void weight_prep_up_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wx2, __nv_bfloat16* __restrict__ ws2,
                                      int Co, int Ci, int Co_pad, int Ci_pad) {
    const long long n = 16LL * Co_pad * Ci_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (wx2) {
            const int ci = (int)(i % Ci_pad);
            const int tap = (int)((i / Ci_pad) % 4);
            const int co = (int)((i / (4LL * Ci_pad)) % Co_pad);
            const int ph = (int)(i / (4LL * Ci_pad * Co_pad));
            wx2[i] = __float2bfloat16((co < Co && ci < Ci) ? up_tap_sum(w + ((long long)co * Ci + ci) * 9, ph >> 1, tap >> 1, ph & 1, tap & 1) : 0.f);
        }
        if (ws2) {
            const int co = (int)(i % Co_pad);
            const int t16 = (int)((i / Co_pad) % 16);
            const int ci = (int)(i / (16LL * Co_pad));
            // r4 = 3 - 2u - a  <=>  (a, u) = (1,1), (0,1), (1,0), (0,0) for r4 = 0..3
            const int r4 = t16 >> 2, s4 = t16 & 3;
            ws2[i] = __float2bfloat16((co < Co && ci < Ci)
                                          ? up_tap_sum(w + ((long long)co * Ci + ci) * 9, (r4 & 1) ? 0 : 1, r4 < 2 ? 1 : 0, (s4 & 1) ? 0 : 1, s4 < 2 ? 1 : 0)
                                          : 0.f);
        }
    }
}

// partial slabs of fv_conv2d_wgrad_x2 [splits][4 phases][Co_pad][4 taps][Ci_pad] -> grad fp32 [Co][Ci][3][3]:
// dW[r][s] = sum over the (phase, tap) pairs whose coarse tap contains (r, s); splits added in order inside each term.
__global__ void wgrad_finish_up_kernel(const float* __restrict__ part, float* __restrict__ grad, int Co, int Ci, int Co_pad, int Ci_pad,
                                       int accumulate, int splits, long long split_stride) {
    // one thread per (co, ci), ci fastest: the 16 phase-tap values are read coalesced (each summed over the splits in split
    // order), then scattered into the 3x3 positions their coarse tap covers
    const long long total = (long long)Co * Ci_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Ci_pad);
        if (ci >= Ci) continue;
        const int co = (int)(i / Ci_pad);
        float w9[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) w9[k] = 0.f;
#pragma unroll
        for (int ph = 0; ph < 4; ++ph)
#pragma unroll
            for (int tap = 0; tap < 4; ++tap) {
                const float* src = part + (((long long)ph * Co_pad + co) * 4 + tap) * Ci_pad + ci;
                float t = 0.f;
                int sp = 0;
                for (; sp + 4 <= splits; sp += 4) {
                    const float a0 = __ldg(src + (sp + 0) * split_stride), a1 = __ldg(src + (sp + 1) * split_stride),
                                a2 = __ldg(src + (sp + 2) * split_stride), a3 = __ldg(src + (sp + 3) * split_stride);
                    t += a0; t += a1; t += a2; t += a3;
                }
                for (; sp < splits; ++sp) t += __ldg(src + sp * split_stride);
                int r0, r1, s0, s1;
                up_rows(ph >> 1, tap >> 1, r0, r1);
                up_rows(ph & 1, tap & 1, s0, s1);
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int s_ = 0; s_ < 3; ++s_)
                        if (r >= r0 && r <= r1 && s_ >= s0 && s_ <= s1) w9[r * 3 + s_] += t;
            }
        float* dst = grad + ((long long)co * Ci + ci) * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) dst[k] = accumulate ? dst[k] + w9[k] : w9[k];
    }
}

// few outputs, many slabs (the per-CTA slabs of the tap-folded out_conv weight gradient: 4.7 k floats x 148): the kernel above has
// one thread walk all slabs of its output serially (24 us of dependent loads).  Here 8 thread groups of a block each add a contiguous
// eighth of the slabs, in slab order, and the eight partial sums are combined in group order: a fixed grouping, still reproducible.
__global__ void __launch_bounds__(256) slab_sum_wide_kernel(const float4* __restrict__ part, int slabs, long long stride4, float4* __restrict__ out, long long n4,
                                                            int accumulate) {
    __shared__ float4 sm[8][32];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * 32 + lane;
    const int per = (slabs + 7) / 8, s0 = g * per, s1 = min(slabs, s0 + per);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
        int sp = s0;
        for (; sp + 4 <= s1; sp += 4) {
            const float4 a0 = __ldg(part + (sp + 0) * stride4 + i), a1 = __ldg(part + (sp + 1) * stride4 + i), a2 = __ldg(part + (sp + 2) * stride4 + i),
                         a3 = __ldg(part + (sp + 3) * stride4 + i);
            v.x += a0.x; v.y += a0.y; v.z += a0.z; v.w += a0.w;
            v.x += a1.x; v.y += a1.y; v.z += a1.z; v.w += a1.w;
            v.x += a2.x; v.y += a2.y; v.z += a2.z; v.w += a2.w;
            v.x += a3.x; v.y += a3.y; v.z += a3.z; v.w += a3.w;
        }
        for (; sp < s1; ++sp) {
            const float4 a = __ldg(part + sp * stride4 + i);
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
    }
    sm[g][lane] = v;
    __syncthreads();
    if (g == 0 && i < n4) {
        float4 t = sm[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) { const float4 a = sm[k][lane]; t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w; }
        if (accumulate) { const float4 c = out[i]; t.x += c.x; t.y += c.y; t.z += c.z; t.w += c.w; }
        out[i] = t;
    }
}

// ---- filters of the 4x4 stride-2 convolution (Conv2dELR as used by EFE_conv6, reference models_utils.py:632-744) ---------------
// w fp32 [Co][Ci][4][4] -> wf bf16 [Co_pad][16][Ci_pad] (fv_conv2d_s2 operand) and wx2 bf16 [4 phases][Ci_pad][4 taps][Co_pad]
// (fv_conv2d_x2 operand of the data gradient: phase (a, b), tap (u, v) <- filter tap (3 - 2u - a, 3 - 2v - b)).
__global__ void weight_prep_s2_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wx2,
                                      int Co, int Ci, int Co_pad, int Ci_pad) {
    const long long n = 16LL * Co_pad * Ci_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (wf) {
            const int ci = (int)(i % Ci_pad);
            const int tap = (int)((i / Ci_pad) % 16);
            const int co = (int)(i / (16LL * Ci_pad));
            wf[i] = __float2bfloat16((co < Co && ci < Ci) ? w[((long long)co * Ci + ci) * 16 + tap] : 0.f);
        }
        if (wx2) {
            const int co = (int)(i % Co_pad);
            const int tap = (int)((i / Co_pad) % 4);
            const int ci = (int)((i / (4LL * Co_pad)) % Ci_pad);
            const int ph = (int)(i / (4LL * Co_pad * Ci_pad));
            const int r4 = 3 - 2 * (tap >> 1) - (ph >> 1), s4 = 3 - 2 * (tap & 1) - (ph & 1);
            wx2[i] = __float2bfloat16((co < Co && ci < Ci) ? w[((long long)co * Ci + ci) * 16 + r4 * 4 + s4] : 0.f);
        }
    }
}

// ---- input pre-scale of EFE_conv5 / EFE_conv6 (reference models.py:764, 872) ---------------------------------------------------------
// F.interpolate(x, mode="bilinear", scale_factor=s, align_corners=False, recompute_scale_factor=True) on NCHW fp32 frames:
// Ho = floor(H * s); source index = (o + 0.5) * (H / Ho) - 0.5, clamped at 0 (ATen area_pixel_compute_source_index); the four
// neighbours are blended in fp32.  For the reference's s = 0.25 on sizes divisible by 4 this is the mean of the 2x2 block at
// (4i + 1 .. 4i + 2, 4j + 1 .. 4j + 2).  One thread per output pixel, consecutive threads = consecutive output columns.
__global__ void bilinear_resize_kernel(const float* __restrict__ x, float* __restrict__ out, long long NC, int H, int W, int Ho, int Wo,
                                       float rh, float rw) {
    const long long total = NC * Ho * Wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ow = (int)(i % Wo), oh = (int)((i / Wo) % Ho);
        const long long nc = i / ((long long)Wo * Ho);
        float sh = rh * (oh + 0.5f) - 0.5f, sw = rw * (ow + 0.5f) - 0.5f;
        sh = sh < 0.f ? 0.f : sh;
        sw = sw < 0.f ? 0.f : sw;
        const int h0 = (int)sh, w0 = (int)sw;
        const int h1 = h0 + (h0 < H - 1 ? 1 : 0), w1 = w0 + (w0 < W - 1 ? 1 : 0);
        const float lh1 = sh - h0, lw1 = sw - w0, lh0 = 1.f - lh1, lw0 = 1.f - lw1;
        const float* p = x + nc * (long long)H * W;
        out[i] = lh0 * (lw0 * __ldg(p + (long long)h0 * W + w0) + lw1 * __ldg(p + (long long)h0 * W + w1)) +
                 lh1 * (lw0 * __ldg(p + (long long)h1 * W + w0) + lw1 * __ldg(p + (long long)h1 * W + w1));
    }
}

// ---- Conv2dELR pieces (reference models_utils.py:632-744) --------------------------------------------------------------------------
// weff[co][:] = gain * w[co][:] / max(||w[co][:]||, 1e-12)  (norm == "demod": F.normalize over dims 1,2,3, then weightgain), or
// gain * w when demod == 0.  One block per output channel; inv_norm[co] is kept for the backward.
__global__ void demod_fwd_kernel(const float* __restrict__ w, float* __restrict__ weff, float* __restrict__ inv_norm, int K, float gain, int demod) {
    __shared__ float red[32];
    const float* row = w + (size_t)blockIdx.x * K;
    float inv = 1.f;
    if (demod) {
        float a = 0.f;
        for (int i = threadIdx.x; i < K; i += blockDim.x) a = fmaf(row[i], row[i], a);
        a = warp_sum(a);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
        __syncthreads();
        float t = 0.f;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) t += red[wv];
        inv = 1.f / fmaxf(sqrtf(t), 1e-12f);
    }
    if (threadIdx.x == 0 && inv_norm) inv_norm[blockIdx.x] = inv;
    for (int i = threadIdx.x; i < K; i += blockDim.x) weff[(size_t)blockIdx.x * K + i] = row[i] * inv * gain;
}
// dw = gain * inv * (dweff - what * <what, dweff>),  what = w * inv   (gradient of w / ||w||); demod == 0: dw = gain * dweff
__global__ void demod_bwd_kernel(const float* __restrict__ w, const float* __restrict__ inv_norm, const float* __restrict__ dweff,
                                 float* __restrict__ dw, int K, float gain, int demod) {
    __shared__ float red[32];
    const size_t base = (size_t)blockIdx.x * K;
    if (!demod) {
        for (int i = threadIdx.x; i < K; i += blockDim.x) dw[base + i] = gain * dweff[base + i];
        return;
    }
    const float inv = inv_norm[blockIdx.x];
    float a = 0.f;
    for (int i = threadIdx.x; i < K; i += blockDim.x) a = fmaf(w[base + i] * inv, dweff[base + i], a);
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    float dot = 0.f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) dot += red[wv];
    for (int i = threadIdx.x; i < K; i += blockDim.x) dw[base + i] = gain * inv * (dweff[base + i] - w[base + i] * inv * dot);
}
// dy = g * act'(out) for an activation applied in the conv epilogue (ReLU / LeakyReLU(0.2): the derivative follows from the sign of
// the stored output); NHWC bf16, 8 elements per thread
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ dy,
                               long long n8, int act) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float o[8], gg[8], r[8];
        V8<__nv_bfloat16>::load(out + i * 8, o);
        V8<__nv_bfloat16>::load(g + i * 8, gg);
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = gg[k] * act_grad(o[k], act);
        V8<__nv_bfloat16>::store(dy + i * 8, r);
    }
}

// ---------------------------------------------------------------- batch-norm statistics
// sums[0..C) = sum_p y[p,c], sums[C..2C) = sum_p y[p,c]^2 over P rows of an NHWC tensor (channel stride C).
// blockDim = 256; a block covers rows_per_iter = 256 / (C/8) rows per step; cross-row reduction through shared memory,
// then the deterministic cross-block reduction of fv_reduce.cuh (fixed summation order: bitwise reproducible).
template <typename T>
__global__ void bn_stats_kernel(const T* __restrict__ y, float* __restrict__ sums, long long P, int C, void* ws, const XrankArgs xr) {
    extern __shared__ float sh[];   // [2][rows_per_iter][C] | block vector [2C] | totals [2C]
    __shared__ int red_flag;
    __shared__ uint32_t xr_epoch;
    const int tpr = C / 8;                       // threads per row
    const int rpi = blockDim.x / tpr;            // rows per iteration
    const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (tr < rpi) {
        const long long stride = (long long)gridDim.x * rpi;
        long long r = (long long)blockIdx.x * rpi + tr;
        // eight independent 16-byte loads in flight per thread; the main loop is unpredicated (conditional loads cost ~10 %
        // of the bandwidth), the last partial batch is issued in one predicated batch (rows past the end: zeros)
        auto batch = [&](long long rb, auto tag) {
            constexpr bool PRED = decltype(tag)::value;
            Raw8<T> raw[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (!PRED || rb + u * stride < P) raw[u].load(y + (rb + u * stride) * C + tc * 8);
                else raw[u].zero();
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float f[8];
                raw[u].unpack(f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    s[k] += f[k];
                    q[k] = fmaf(f[k], f[k], q[k]);
                }
            }
        };
        for (; r + 7 * stride < P; r += 8 * stride) batch(r, std::false_type{});
        if (r < P) batch(r, std::true_type{});
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            sh[tr * C + tc * 8 + k] = s[k];
            sh[(rpi + tr) * C + tc * 8 + k] = q[k];
        }
    }
    __syncthreads();
    float* blk = sh + 2 * rpi * C;
    float* tot = blk + 2 * C;
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
        const int which = c / C, ch = c % C;
        float a = 0.f;
        for (int r = 0; r < rpi; ++r) a += sh[(which * rpi + r) * C + ch];
        blk[c] = a;
    }
    __syncthreads();
    if (det_reduce<float>(ws, 2 * C, gridDim.x, blockIdx.x, blk, tot, threadIdx.x, blockDim.x, BlockSync{}, &red_flag)) {
        for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sums[c] = tot[c];
        if (xr.peer_bufs) {         // data parallel: the block holding the totals exchanges them with the other ranks and finalizes
            __syncthreads();
            xrank_exchange_finalize(xr, tot, &xr_epoch);
        }
    }
}

// mean / invstd / folded scale+shift from the (possibly cross-rank reduced) sums; running-stat update with the
// unbiased variance (momentum form of nn.SyncBatchNorm, reference modules.py:19).
__global__ void bn_finalize_kernel(const float* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var,
                                   float momentum, float eps, float* __restrict__ stat /* [4][C]: mean invstd scale shift */,
                                   int C) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        const double mean = (double)sums[c] / count;
        double var = (double)sums[C + c] / count - mean * mean;
        if (var < 0) var = 0;
        const float invstd = (float)(1.0 / sqrt(var + (double)eps));
        const float sc = gamma[c] * invstd;
        stat[c] = (float)mean;
        stat[C + c] = invstd;
        stat[2 * C + c] = sc;
        stat[3 * C + c] = beta[c] - (float)mean * sc;
        if (running_mean) {
            const double unbiased = count > 1 ? var * count / (count - 1) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    }
}

// eval mode: scale/shift from the running statistics
__global__ void bn_eval_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      float* __restrict__ stat, int C) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        const float invstd = rsqrtf(rv[c] + eps);
        const float sc = gamma[c] * invstd;
        stat[c] = rm[c];
        stat[C + c] = invstd;
        stat[2 * C + c] = sc;
        stat[3 * C + c] = beta[c] - rm[c] * sc;
    }
}

// ---------------------------------------------------------------- norm + act (+pool / +upsample) forward
// a = act(scale[c] * y + shift[c]); mode POOL averages 2x2 windows on the way out (DownBlock2D, reference
// modules.py:59-70), mode UP replicates each pixel 2x2 (the nn.Upsample in front of UpBlock2D's conv, modules.py:78-89).
// H, W are the INPUT spatial sizes.  Output is NHWC (TO) or, with nchw_out, NCHW fp32.
// With fin.sums the kernel finalizes the statistics itself (what bn_finalize_kernel does: every thread derives scale / shift
// of its 8 channels from the sums; block 0 also writes the [4][C] stat block kept for backward and updates the running
// statistics), which removes one tiny launch per batch-norm layer from the step.
struct BnFin {
    const float* sums;      // [2][C] sum | sum of squares, or null: use the finished stat block
    double count;
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    float momentum, eps;
    float* stat_out;        // [4][C]
};
template <typename TI, typename TO, int MODE>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const TI* __restrict__ y, const float* __restrict__ stat, TO* __restrict__ out, int N, int H, int W, int C, int act,
                  int nchw_out, const BnFin fin) {
    const float* scale = stat + 2 * C;
    const float* shift = stat + 3 * C;
    const unsigned groups = C / 8;
    const unsigned Ho = MODE == FV_MODE_POOL ? H / 2 : H, Wo = MODE == FV_MODE_POOL ? W / 2 : W;   // iteration domain
    const unsigned total = (unsigned)N * Ho * Wo * groups;
    // blockDim (256) is a multiple of groups, so a thread keeps the same channel group for its whole grid-stride walk
    const unsigned g = threadIdx.x % groups;
    float sc[8], sf[8];
    if (fin.sums) {
        const bool writer = blockIdx.x == 0 && threadIdx.x < groups;      // one thread per channel group writes the results
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = g * 8 + k;
            // no fp64 division / square root here (every thread of the grid runs this, and fp64 is a trickle on this part):
            // double only for the cancellation-prone E[y^2] - mean^2, the rest in fp32
            const double inv_count = 1.0 / fin.count;            // uniform: one division per thread, hoisted by the compiler
            const double mean = (double)__ldg(fin.sums + c) * inv_count;
            double var = fma((double)__ldg(fin.sums + C + c), inv_count, -mean * mean);
            if (var < 0) var = 0;
            const float invstd = 1.0f / sqrtf((float)var + fin.eps);
            sc[k] = __ldg(fin.gamma + c) * invstd;
            sf[k] = __ldg(fin.beta + c) - (float)mean * sc[k];
            if (writer) {
                fin.stat_out[c] = (float)mean;
                fin.stat_out[C + c] = invstd;
                fin.stat_out[2 * C + c] = sc[k];
                fin.stat_out[3 * C + c] = sf[k];
                if (fin.running_mean) {
                    const double unbiased = fin.count > 1 ? var * fin.count / (fin.count - 1) : var;
                    fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (float)mean;
                    fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (float)unbiased;
                }
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            sc[k] = __ldg(scale + g * 8 + k);
            sf[k] = __ldg(shift + g * 8 + k);
        }
    }
    constexpr int NL = MODE == FV_MODE_POOL ? 4 : 1;          // loads per element
    constexpr int U = MODE == FV_MODE_POOL ? 2 : 4;           // elements in flight per thread
    const FastDiv dG((unsigned)groups), dWo((unsigned)Wo), dHo((unsigned)Ho);
    auto issue = [&](unsigned i, Raw8<TI> (&raw)[NL], unsigned& n, unsigned& ho, unsigned& wo) {
        unsigned pix, gdummy, t2;
        dG.divmod(i, pix, gdummy);
        dWo.divmod(pix, t2, wo);
        dHo.divmod(t2, n, ho);
        if (MODE == FV_MODE_POOL) {
#pragma unroll
            for (int d = 0; d < 4; ++d)
                raw[MODE == FV_MODE_POOL ? d : 0].load(y + (((size_t)n * H + 2 * ho + (d >> 1)) * W + 2 * wo + (d & 1)) * C + g * 8);
        } else {
            raw[0].load(y + (size_t)pix * C + g * 8);
        }
    };
    auto finish = [&](const Raw8<TI> (&raw)[NL], unsigned n, unsigned ho, unsigned wo) {
        float r[8];
        if (MODE == FV_MODE_POOL) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = 0.f;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                float f[8];
                raw[MODE == FV_MODE_POOL ? d : 0].unpack(f);
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] += act_fwd(fmaf(f[k], sc[k], sf[k]), act);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] *= 0.25f;
        } else {
            float f[8];
            raw[0].unpack(f);
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = act_fwd(fmaf(f[k], sc[k], sf[k]), act);
        }
        if (MODE == FV_MODE_UP) {
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx)
                    V8<TO>::store(out + (((size_t)n * 2 * H + 2 * ho + dy) * (2 * W) + 2 * wo + dx) * C + g * 8, r);
        } else if (nchw_out) {
            float* o = reinterpret_cast<float*>(out);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[(((size_t)n * C + g * 8 + k) * Ho + ho) * Wo + wo] = r[k];
        } else {
            V8<TO>::store(out + (((size_t)n * Ho + ho) * Wo + wo) * C + g * 8, r);
        }
    };
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
    auto batch = [&](unsigned ib, auto tag) {      // U independent elements: all loads first
        constexpr bool PRED = decltype(tag)::value;
        Raw8<TI> raw[U][NL];
        unsigned n[U], ho[U], wo[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!PRED || ib + u * stride < total) issue(ib + u * stride, raw[u], n[u], ho[u], wo[u]);
            else {
#pragma unroll
                for (int d = 0; d < NL; ++d) raw[u][d].zero();
                n[u] = ho[u] = wo[u] = 0;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (!PRED || ib + u * stride < total) finish(raw[u], n[u], ho[u], wo[u]);
    };
    for (; i0 + (U - 1) * stride < total; i0 += U * stride) batch(i0, std::false_type{});
    if (i0 < total) batch(i0, std::true_type{});
}

// ---------------------------------------------------------------- norm + act backward
// Effective upstream gradient at the (pre-pool / pre-upsample) activation:  POOL: g/4 of the pooled pixel,
// UP: sum of the 2x2 replicated pixels, NONE: g.  g is NHWC (TG) or NCHW fp32 when GN.  MODE / GN are compile-time so
// the unrolled main loops below carry no branches and the compiler can keep four rows of loads in flight.
template <typename TG, int MODE, bool GN>
struct GLoad {                                      // phase 1: issue the loads (raw registers); phase 2: convert / combine
    static constexpr int NR = (MODE == FV_MODE_UP) ? 4 : 1;
    Raw8<TG> raw[GN ? 1 : NR];
    float nchw[GN ? 8 : 1];
    __device__ __forceinline__ void issue(const TG* __restrict__ g, unsigned n, unsigned h, unsigned w, int H, int W, int C, unsigned grp) {
        if (GN) {
            const float* gf = reinterpret_cast<const float*>(g);
            const int Hg = MODE == FV_MODE_POOL ? H / 2 : H, Wg = MODE == FV_MODE_POOL ? W / 2 : W;
            const unsigned hg = MODE == FV_MODE_POOL ? h / 2 : h, wg = MODE == FV_MODE_POOL ? w / 2 : w;
#pragma unroll
            for (int k = 0; k < 8; ++k) nchw[k] = __ldg(gf + (((size_t)n * C + grp * 8 + k) * Hg + hg) * Wg + wg);
        } else if (MODE == FV_MODE_POOL) {
            raw[0].load(g + (((size_t)n * (H / 2) + h / 2) * (W / 2) + w / 2) * C + grp * 8);
        } else if (MODE == FV_MODE_UP) {
#pragma unroll
            for (int d = 0; d < 4; ++d)
                raw[GN ? 0 : d].load(g + (((size_t)n * 2 * H + 2 * h + (d >> 1)) * (2 * W) + 2 * w + (d & 1)) * C + grp * 8);
        } else {
            raw[0].load(g + (((size_t)n * H + h) * W + w) * C + grp * 8);
        }
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int d = 0; d < (GN ? 1 : NR); ++d) raw[d].zero();
#pragma unroll
        for (int k = 0; k < (GN ? 8 : 1); ++k) nchw[k] = 0.f;
    }
    __device__ __forceinline__ void finish(float (&r)[8]) const {
        if (GN) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = MODE == FV_MODE_POOL ? 0.25f * nchw[k] : nchw[k];
        } else if (MODE == FV_MODE_UP) {
            float f[4][8];
#pragma unroll
            for (int d = 0; d < 4; ++d) raw[GN ? 0 : d].unpack(f[d]);
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = (f[0][k] + f[1][k]) + (f[2][k] + f[3][k]);
        } else {
            raw[0].unpack(r);
            if (MODE == FV_MODE_POOL) {
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] *= 0.25f;
            }
        }
    }
};

__device__ __forceinline__ void row_to_nhw(unsigned r, const FastDiv& dH, const FastDiv& dW, unsigned& n, unsigned& h, unsigned& w) {
    unsigned t2;
    dW.divmod(r, t2, w);
    dH.divmod(t2, n, h);
}

// pass 1: sums[0..C) += sum dz, sums[C..2C) += sum dz * xhat,  dz = g_eff * act'(scale*y+shift), xhat = (y-mean)*invstd.
// The loop accumulates sum(dz) and sum(dz*y) only (xhat is affine in y), which keeps the per-thread constants down to
// scale/shift and lets eight rows of raw 16-byte loads stay in flight.
template <typename TY, typename TG, int MODE, bool GN>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_reduce_kernel(const TY* __restrict__ y, const TG* __restrict__ g, const float* __restrict__ stat,
                         float* __restrict__ sums, int N, int H, int W, int C, int act, void* ws, const XrankArgs xr) {
    extern __shared__ float sh[];   // [2][rows_per_iter][C] | block vector [2C] | totals [2C]
    __shared__ int red_flag;
    __shared__ uint32_t xr_epoch;
    const int tpr = C / 8, rpi = blockDim.x / tpr;
    const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
    const unsigned P = (unsigned)N * H * W;
    const FastDiv dH((unsigned)H), dW((unsigned)W);
    float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sy[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float sc[8], sf[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sc[k] = __ldg(stat + 2 * C + tc * 8 + k);
        sf[k] = __ldg(stat + 3 * C + tc * 8 + k);
    }
    constexpr bool NEED_NHW = (MODE != FV_MODE_NONE) || GN;
    constexpr int U = MODE == FV_MODE_UP ? 2 : 8;
    const unsigned stride = gridDim.x * rpi;
    unsigned r0 = blockIdx.x * rpi + tr;
    auto batch = [&](unsigned rb, auto tag) {
        constexpr bool PRED = decltype(tag)::value;
        Raw8<TY> yr[U];
        GLoad<TG, MODE, GN> gl[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {                         // U rows of y and g in flight, 4 registers per 16-byte load
            const unsigned row = rb + u * stride;             // (tail batch) rows past the end load nothing: g = 0
            if (!PRED || row < P) {
                unsigned n = 0, h = 0, w = row;               // MODE_NONE, NHWC g: the row index is the g index
                if (NEED_NHW) row_to_nhw(row, dH, dW, n, h, w);
                yr[u].load(y + (size_t)row * C + tc * 8);
                gl[u].issue(g, n, h, w, NEED_NHW ? H : 1, NEED_NHW ? W : (int)P, C, tc);
            } else {
                yr[u].zero();
                gl[u].zero();
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float f[8], ge[8];
            yr[u].unpack(f);
            gl[u].finish(ge);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float dz = ge[k] * act_grad(fmaf(f[k], sc[k], sf[k]), act);
                s1[k] += dz;
                sy[k] = fmaf(dz, f[k], sy[k]);
            }
        }
    };
    for (; r0 + (U - 1) * stride < P; r0 += U * stride) batch(r0, std::false_type{});
    if (r0 < P) batch(r0, std::true_type{});
#pragma unroll
    for (int k = 0; k < 8; ++k) {                             // sum dz*xhat = invstd * (sum dz*y - mean * sum dz)
        const float mean = __ldg(stat + tc * 8 + k), invstd = __ldg(stat + C + tc * 8 + k);
        sh[tr * C + tc * 8 + k] = s1[k];
        sh[(rpi + tr) * C + tc * 8 + k] = invstd * (sy[k] - mean * s1[k]);
    }
    __syncthreads();
    float* blk = sh + 2 * rpi * C;
    float* tot = blk + 2 * C;
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
        const int which = c / C, ch = c % C;
        float a = 0.f;
        for (int r = 0; r < rpi; ++r) a += sh[(which * rpi + r) * C + ch];
        blk[c] = a;
    }
    __syncthreads();
    if (det_reduce<float>(ws, 2 * C, gridDim.x, blockIdx.x, blk, tot, threadIdx.x, blockDim.x, BlockSync{}, &red_flag)) {
        for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sums[c] = tot[c];
        if (xr.peer_bufs) {
            __syncthreads();
            xrank_exchange_finalize(xr, tot, &xr_epoch);
        }
    }
}

// dgamma += s2, dbeta += s1 (local sums: autograd's gradient all-reduce averages them later); coef = (cross-rank) sums / count
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ sums_local, const float* __restrict__ sums_global, double count,
                                       float* dgamma, float* dbeta, float* __restrict__ coef, int C, int accumulate) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        const float s1 = sums_local[c], s2 = sums_local[C + c];
        if (dbeta) dbeta[c] = accumulate ? dbeta[c] + s1 : s1;
        if (dgamma) dgamma[c] = accumulate ? dgamma[c] + s2 : s2;
        coef[c] = (float)((double)sums_global[c] / count);
        coef[C + c] = (float)((double)sums_global[C + c] / count);
    }
}

// pass 2: dy = scale * (dz - c1 - xhat * c2) (+ add) = scale*dz + A + B*y with per-channel A = scale*(c2*invstd*mean - c1),
// B = -scale*c2*invstd; written bf16 NHWC: the conv-output gradient fed to dgrad / wgrad.
// With fin_sums (the backward sums of pass 1) the coupling coefficients are derived here (sums / count) and block 0 writes
// dgamma / dbeta: the work of bn_bwd_finalize_kernel without its launch (single-process training; with several ranks the
// cross-rank exchange kernel produces coef).
template <typename TY, typename TG, int MODE, bool GN, bool ADD>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_apply_kernel(const TY* __restrict__ y, const TG* __restrict__ g, const float* __restrict__ stat,
                        const float* __restrict__ coef, const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ dy,
                        int N, int H, int W, int C, int act, const float* __restrict__ fin_sums, double fin_count,
                        float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int tpr = C / 8, rpi = blockDim.x / tpr;
    const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
    const unsigned P = (unsigned)N * H * W;
    const FastDiv dH((unsigned)H), dW((unsigned)W);
    float sc[8], sf[8], ca[8], cb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = tc * 8 + k;
        const float mean = __ldg(stat + c), invstd = __ldg(stat + C + c);
        sc[k] = __ldg(stat + 2 * C + c);
        sf[k] = __ldg(stat + 3 * C + c);
        float c1, c2;
        if (fin_sums) {
            const float s1 = __ldg(fin_sums + c), s2 = __ldg(fin_sums + C + c);
            const float inv_count = (float)(1.0 / fin_count);
            c1 = s1 * inv_count;
            c2 = s2 * inv_count;
            if (blockIdx.x == 0 && tr == 0) {                    // one thread per channel group
                if (dbeta) dbeta[c] = s1;
                if (dgamma) dgamma[c] = s2;
            }
        } else {
            c1 = __ldg(coef + c);
            c2 = __ldg(coef + C + c);
        }
        ca[k] = sc[k] * (c2 * invstd * mean - c1);
        cb[k] = -sc[k] * c2 * invstd;
    }
    constexpr bool NEED_NHW = (MODE != FV_MODE_NONE) || GN;
    constexpr int U = MODE == FV_MODE_UP ? 2 : (ADD ? 4 : 6);
    const unsigned stride = gridDim.x * rpi;
    unsigned r0 = blockIdx.x * rpi + tr;
    auto batch = [&](unsigned rb, auto tag) {
        constexpr bool PRED = decltype(tag)::value;
        Raw8<TY> yr[U];
        GLoad<TG, MODE, GN> gl[U];
        Raw8<__nv_bfloat16> ar[ADD ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned row = rb + u * stride;
            if (!PRED || row < P) {
                unsigned n = 0, h = 0, w = row;
                if (NEED_NHW) row_to_nhw(row, dH, dW, n, h, w);
                yr[u].load(y + (size_t)row * C + tc * 8);
                gl[u].issue(g, n, h, w, NEED_NHW ? H : 1, NEED_NHW ? W : (int)P, C, tc);
                if (ADD) ar[ADD ? u : 0].load(add + (size_t)row * C + tc * 8);
            } else {
                yr[u].zero();
                gl[u].zero();
                if (ADD) ar[ADD ? u : 0].zero();
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float f[8], ge[8], a[8], o[8];
            yr[u].unpack(f);
            gl[u].finish(ge);
            if (ADD) ar[ADD ? u : 0].unpack(a);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float dz = ge[k] * act_grad(fmaf(f[k], sc[k], sf[k]), act);
                o[k] = fmaf(sc[k], dz, fmaf(cb[k], f[k], ca[k]));
                if (ADD) o[k] += a[k];
            }
            if (!PRED || rb + u * stride < P) V8<__nv_bfloat16>::store(dy + (size_t)(rb + u * stride) * C + tc * 8, o);
        }
    };
    for (; r0 + (U - 1) * stride < P; r0 += U * stride) batch(r0, std::false_type{});
    if (r0 < P) batch(r0, std::true_type{});
}

// per-channel column sums of an NHWC bf16 tensor (bias gradient of a conv that does not feed a batch norm)
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ y, float* __restrict__ sums, long long P, int C, long long ld, void* ws) {
    extern __shared__ float sh[];   // [rows_per_iter][C] | block vector [C] | totals [C]
    __shared__ int red_flag;
    const int tpr = C / 8, rpi = blockDim.x / tpr;
    const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (tr < rpi) {
        for (long long r = (long long)blockIdx.x * rpi + tr; r < P; r += (long long)gridDim.x * rpi) {
            float f[8];
            V8<__nv_bfloat16>::load(y + r * ld + tc * 8, f);
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] += f[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) sh[tr * C + tc * 8 + k] = s[k];
    }
    __syncthreads();
    float* blk = sh + rpi * C;
    float* tot = blk + C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f;
        for (int r = 0; r < rpi; ++r) a += sh[r * C + c];
        blk[c] = a;
    }
    __syncthreads();
    if (det_reduce<float>(ws, C, gridDim.x, blockIdx.x, blk, tot, threadIdx.x, blockDim.x, BlockSync{}, &red_flag))
        for (int c = threadIdx.x; c < C; c += blockDim.x) sums[c] = tot[c];
}

// ---------------------------------------------------------------- instance norm
// nn.InstanceNorm2d(C, affine=True) + LeakyReLU of the Discriminator's blocks (reference modules.py:21, models.py:1120-1127):
// statistics per (image, channel) over H*W (biased variance, eps), no running statistics.  One block per (image, 64-channel
// slice): 32 row lanes x 8 channel groups; the row lanes are combined in lane order -- no cross-block reduction, reproducible.
// in_stats:   stat[n][0][c] = mean, stat[n][1][c] = invstd
// in_act_fwd: out = act(gamma * (y - mean) * invstd + beta)
// in_bwd_sums: sums[n][0][c] = sum dz, sums[n][1][c] = sum dz * xhat   (dz = g * act'(...)); dgamma / dbeta = their sums over n
// in_bwd_apply: dy = gamma * invstd * (dz - mean_p(dz) - xhat * mean_p(dz * xhat))
template <int PASS>     // 0: statistics of y; 1: backward sums
__global__ void __launch_bounds__(256)
in_reduce_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ g, const float* __restrict__ stat,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out, int P, int C, int act, float eps) {
    __shared__ float sh[2][32][64];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const int tc = threadIdx.x & 7, tr = threadIdx.x >> 3;
    const int c = c0 + tc * 8;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float mean[8], inv[8], ga[8], be[8];
    if (PASS == 1 && c < C) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            mean[k] = stat[((size_t)n * 2) * C + c + k];
            inv[k] = stat[((size_t)n * 2 + 1) * C + c + k];
            ga[k] = gamma[c + k];
            be[k] = beta[c + k];
        }
    }
    if (c < C) {
        for (int r = tr; r < P; r += 32) {
            float f[8];
            V8<__nv_bfloat16>::load(y + ((size_t)n * P + r) * C + c, f);
            if (PASS == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    a[k] += f[k];
                    b[k] = fmaf(f[k], f[k], b[k]);
                }
            } else {
                float gg[8];
                V8<__nv_bfloat16>::load(g + ((size_t)n * P + r) * C + c, gg);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float xh = (f[k] - mean[k]) * inv[k];
                    const float dz = gg[k] * act_grad(fmaf(ga[k], xh, be[k]), act);
                    a[k] += dz;
                    b[k] = fmaf(dz, xh, b[k]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sh[0][tr][tc * 8 + k] = a[k];
        sh[1][tr][tc * 8 + k] = b[k];
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int which = threadIdx.x >> 6, ch = threadIdx.x & 63;
        if (c0 + ch < C) {
            float t = 0.f;
            for (int r = 0; r < 32; ++r) t += sh[which][r][ch];
            if (PASS == 0) {
                // mean / invstd need both sums: the `which == 0` thread writes the mean, its partner computes the variance
                sh[which][0][ch] = t;
            } else {
                out[((size_t)n * 2 + which) * C + c0 + ch] = t;
            }
        }
    }
    if (PASS == 0) {
        __syncthreads();
        if (threadIdx.x < 64 && c0 + threadIdx.x < C) {
            const double m = (double)sh[0][0][threadIdx.x] / P;
            double var = (double)sh[1][0][threadIdx.x] / P - m * m;
            if (var < 0) var = 0;
            out[((size_t)n * 2) * C + c0 + threadIdx.x] = (float)m;
            out[((size_t)n * 2 + 1) * C + c0 + threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
        }
    }
}

template <int PASS>     // 0: forward; 1: backward apply (sums = the per-image backward sums)
__global__ void __launch_bounds__(256)
in_apply_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ g, const float* __restrict__ stat,
                const float* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                __nv_bfloat16* __restrict__ out, int N, int P, int C, int act) {
    const int groups = C / 8;
    const long long total = (long long)N * P * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int grp = (int)(i % groups);
        const long long pix = i / groups;
        const int n = (int)(pix / P), c = grp * 8;
        float f[8], r[8], gg[8];
        V8<__nv_bfloat16>::load(y + pix * C + c, f);
        if (PASS == 1) V8<__nv_bfloat16>::load(g + pix * C + c, gg);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float mean = __ldg(stat + ((size_t)n * 2) * C + c + k), inv = __ldg(stat + ((size_t)n * 2 + 1) * C + c + k);
            const float ga = __ldg(gamma + c + k), be = __ldg(beta + c + k);
            const float xh = (f[k] - mean) * inv, z = fmaf(ga, xh, be);
            if (PASS == 0) {
                r[k] = act_fwd(z, act);
            } else {
                const float dz = gg[k] * act_grad(z, act);
                const float c1 = __ldg(sums + ((size_t)n * 2) * C + c + k) / P, c2 = __ldg(sums + ((size_t)n * 2 + 1) * C + c + k) / P;
                r[k] = ga * inv * (dz - c1 - xh * c2);
            }
        }
        V8<__nv_bfloat16>::store(out + pix * C + c, r);
    }
}

// ---------------------------------------------------------------- re-parameterisation + KL
// h = [N][2*Dz] fp32 (mu | logstd, the NCHW flatten of the encoder output; reference models.py:559-560),
// z = mu + exp(logstd) * eps (models.py:561), sum_blocks kl_part[n][.] = sum_d(-0.5 - ls + 0.5 mu^2 + 0.5 exp(2 ls)) (losses.py:392).
// One block row per sample chunk, float4 loads; warp-shuffle reduction, the warps of a block combined in warp order, one
// partial per block written to kl_part[n][blockIdx.x] (the caller adds the few partials of a row: no atomics, reproducible).
__global__ void reparam_kl_fwd_kernel(const float* __restrict__ mu_p, const float* __restrict__ ls_p, long long row_stride,
                                      const float* __restrict__ eps, float* __restrict__ z, float* __restrict__ kl_part,
                                      int Dz) {
    __shared__ float red[32];
    const int n = blockIdx.y;
    const float4* mu4 = reinterpret_cast<const float4*>(mu_p + (long long)n * row_stride);
    const float4* ls4 = reinterpret_cast<const float4*>(ls_p + (long long)n * row_stride);
    const float4* e4 = eps ? reinterpret_cast<const float4*>(eps + (long long)n * Dz) : nullptr;
    float4* z4 = z ? reinterpret_cast<float4*>(z + (long long)n * Dz) : nullptr;
    float acc = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Dz / 4; i += gridDim.x * blockDim.x) {
        const float4 m = __ldg(mu4 + i), l = __ldg(ls4 + i);
        const float sx = expf(l.x), sy = expf(l.y), sz = expf(l.z), sw = expf(l.w);
        if (z4) {
            float4 e = e4 ? __ldg(e4 + i) : make_float4(0, 0, 0, 0);
            z4[i] = make_float4(fmaf(sx, e.x, m.x), fmaf(sy, e.y, m.y), fmaf(sz, e.z, m.z), fmaf(sw, e.w, m.w));
        }
        acc += (-0.5f - l.x + 0.5f * m.x * m.x + 0.5f * sx * sx) + (-0.5f - l.y + 0.5f * m.y * m.y + 0.5f * sy * sy) +
               (-0.5f - l.z + 0.5f * m.z * m.z + 0.5f * sz * sz) + (-0.5f - l.w + 0.5f * m.w * m.w + 0.5f * sw * sw);
    }
    if (kl_part) {
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float v = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
            kl_part[(long long)n * gridDim.x + blockIdx.x] = v;
        }
    }
}

// dmu = dz + kscale * mu (+ dmu_ext);  dls = dz * eps * exp(ls) + kscale * (exp(2 ls) - 1) (+ dls_ext)
// kscale = (upstream dK) / (N * Dz), read from device memory when kscale_ptr is given (times kscale).
__global__ void reparam_kl_bwd_kernel(const float* __restrict__ mu_p, const float* __restrict__ ls_p, long long row_stride,
                                      const float* __restrict__ eps, const float* __restrict__ dz,
                                      const float* __restrict__ dmu_ext, const float* __restrict__ dls_ext,
                                      float kscale, const float* __restrict__ kscale_ptr, float* __restrict__ dmu,
                                      float* __restrict__ dls, long long out_stride, int Dz) {
    const int n = blockIdx.y;
    const float ks = kscale_ptr ? kscale * __ldg(kscale_ptr) : kscale;
    // float4 everywhere: Dz % 4 == 0 and 16-byte aligned rows are checked by the host entry point
    const float4* mu4 = reinterpret_cast<const float4*>(mu_p + (long long)n * row_stride);
    const float4* ls4 = reinterpret_cast<const float4*>(ls_p + (long long)n * row_stride);
    const float4* e4 = eps ? reinterpret_cast<const float4*>(eps + (long long)n * Dz) : nullptr;
    const float4* g4 = dz ? reinterpret_cast<const float4*>(dz + (long long)n * Dz) : nullptr;
    const float4* xm4 = dmu_ext ? reinterpret_cast<const float4*>(dmu_ext + (long long)n * Dz) : nullptr;
    const float4* xl4 = dls_ext ? reinterpret_cast<const float4*>(dls_ext + (long long)n * Dz) : nullptr;
    float4* om4 = reinterpret_cast<float4*>(dmu + (long long)n * out_stride);
    float4* ol4 = reinterpret_cast<float4*>(dls + (long long)n * out_stride);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Dz / 4; i += gridDim.x * blockDim.x) {
        const float4 m = __ldg(mu4 + i), l = __ldg(ls4 + i);
        const float4 g = g4 ? __ldg(g4 + i) : zero4, e = e4 ? __ldg(e4 + i) : zero4;
        const float4 xm = xm4 ? __ldg(xm4 + i) : zero4, xl = xl4 ? __ldg(xl4 + i) : zero4;
        const float sx = expf(l.x), sy = expf(l.y), sz = expf(l.z), sw = expf(l.w);
        om4[i] = make_float4(g.x + ks * m.x + xm.x, g.y + ks * m.y + xm.y, g.z + ks * m.z + xm.z, g.w + ks * m.w + xm.w);
        ol4[i] = make_float4(g.x * e.x * sx + ks * (sx * sx - 1.f) + xl.x, g.y * e.y * sy + ks * (sy * sy - 1.f) + xl.y,
                             g.z * e.z * sz + ks * (sz * sz - 1.f) + xl.z, g.w * e.w * sw + ks * (sw * sw - 1.f) + xl.w);
    }
}

// ---------------------------------------------------------------- reconstruction loss (+ sigmoid) forward + gradient
// pred = sigmoid(logits) if use_sigmoid else logits (reference models.py:1110); loss_sum += sum l(pred - target) with
// l = squared (ReconLoss / nn.MSELoss, losses.py:396-403) or absolute (nn.L1Loss, losses.py:128) error;
// grad = gscale * dl/dlogits, written fp32 (same layout as the inputs) and/or bf16 NHWC padded to Cp channels.
// Block reduce: warp shuffle -> shared -> one atomicAdd per block.
// VEC = 4: a thread owns four consecutive pixels of one image (HW % 4 == 0): every plane access is a 16-byte vector.
template <int VEC>
__global__ void recon_loss_kernel(const float* __restrict__ logits, const float* __restrict__ target, float* __restrict__ pred_out,
                                  float* __restrict__ grad_f32, __nv_bfloat16* __restrict__ grad_nhwc, float* __restrict__ loss_sum,
                                  int N, int C, int HW, int Cp, int l1, int use_sigmoid, float gscale, void* ws) {
    __shared__ float red[34];
    __shared__ int red_flag;
    float acc = 0.f;
    const long long P = (long long)N * HW / VEC;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < P; q += (long long)gridDim.x * blockDim.x) {
        const long long pix = q * VEC;
        const int n = (int)(pix / HW), hw = (int)(pix % HW);
        float gr[VEC][16];
        if (grad_nhwc) {
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int k = 0; k < 16; ++k) gr[v][k] = 0.f;
        }
        for (int c = 0; c < C; ++c) {
            const long long idx = ((long long)n * C + c) * HW + hw;
            float o[VEC], t[VEC], sg[VEC], gd[VEC];
            if (VEC == 4) {
                const float4 o4 = __ldg(reinterpret_cast<const float4*>(logits + idx)), t4 = __ldg(reinterpret_cast<const float4*>(target + idx));
                o[0] = o4.x; o[1 % VEC] = o4.y; o[2 % VEC] = o4.z; o[3 % VEC] = o4.w;
                t[0] = t4.x; t[1 % VEC] = t4.y; t[2 % VEC] = t4.z; t[3 % VEC] = t4.w;
            } else {
                o[0] = __ldg(logits + idx);
                t[0] = __ldg(target + idx);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                sg[v] = use_sigmoid ? 1.f / (1.f + expf(-o[v])) : o[v];
                const float d = sg[v] - t[v];
                acc += l1 ? fabsf(d) : d * d;
                float g = l1 ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : 2.f * d;
                if (use_sigmoid) g *= sg[v] * (1.f - sg[v]);
                gd[v] = g * gscale;
                if (grad_nhwc && c < 16) gr[v][c] = gd[v];
            }
            if (VEC == 4) {
                if (pred_out) *reinterpret_cast<float4*>(pred_out + idx) = make_float4(sg[0], sg[1 % VEC], sg[2 % VEC], sg[3 % VEC]);
                if (grad_f32) *reinterpret_cast<float4*>(grad_f32 + idx) = make_float4(gd[0], gd[1 % VEC], gd[2 % VEC], gd[3 % VEC]);
            } else {
                if (pred_out) pred_out[idx] = sg[0];
                if (grad_f32) grad_f32[idx] = gd[0];
            }
        }
        if (grad_nhwc) {
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                for (int c0 = 0; c0 < Cp; c0 += 8) {
                    float f[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] = (c0 + k < 16) ? gr[v][c0 + k] : 0.f;
                    V8<__nv_bfloat16>::store(grad_nhwc + (pix + v) * Cp + c0, f);
                }
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {                      // warps in order, then blocks in order (fv_reduce.cuh): reproducible
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
        red[32] = v;
    }
    __syncthreads();
    if (det_reduce<float>(ws, 1, gridDim.x, blockIdx.x, red + 32, red + 33, threadIdx.x, blockDim.x, BlockSync{}, &red_flag))
        if (threadIdx.x == 0) loss_sum[0] = red[33];
}

// flat variant for arbitrary same-shape fp32 tensors (the drop-in ReconLoss()((a, b))): float4 vectorised
__global__ void recon_loss_flat_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ grad,
                                       float* __restrict__ loss_sum, long long E, int l1, float gscale, void* ws) {
    __shared__ float red[34];
    __shared__ int red_flag;
    float acc = 0.f;
    const long long E4 = E / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < E4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i), y = __ldg(reinterpret_cast<const float4*>(b) + i);
        const float d[4] = {x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w};
        float g[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc += l1 ? fabsf(d[k]) : d[k] * d[k];
            g[k] = gscale * (l1 ? (d[k] > 0.f ? 1.f : (d[k] < 0.f ? -1.f : 0.f)) : 2.f * d[k]);
        }
        if (grad) reinterpret_cast<float4*>(grad)[i] = make_float4(g[0], g[1], g[2], g[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(E - E4 * 4)) {
        const long long i = E4 * 4 + threadIdx.x;
        const float d = a[i] - b[i];
        acc += l1 ? fabsf(d) : d * d;
        if (grad) grad[i] = gscale * (l1 ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : 2.f * d);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {                      // warps in order, then blocks in order (fv_reduce.cuh): reproducible
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
        red[32] = v;
    }
    __syncthreads();
    if (det_reduce<float>(ws, 1, gridDim.x, blockIdx.x, red + 32, red + 33, threadIdx.x, blockDim.x, BlockSync{}, &red_flag))
        if (threadIdx.x == 0) loss_sum[0] = red[33];
}

// Adam (torch.optim.Adam semantics without amsgrad / weight decay, reference logger.py:60) over a table of tensors in ONE
// launch: blockIdx.y = tensor, grid-stride over its elements.  `step` is the step count AFTER this update, read from
// device memory so that the launch can be captured in a CUDA graph.
__global__ void adam_multi_kernel(const fv_adam_desc* __restrict__ table, float lr, double beta1_d, double beta2_d, float eps,
                                  const float* __restrict__ step) {
    const fv_adam_desc d = table[blockIdx.y];
    // coefficients as torch.optim.Adam forms them: 1 - beta and the bias corrections 1 - beta^t in double (0.999f is 4.7e-5
    // away from 0.999 relative to 1 - beta), once per block
    __shared__ float coef[2];
    if (threadIdx.x == 0) {
        const double t = (double)__ldg(step);
        coef[0] = (float)((double)lr / (1.0 - pow(beta1_d, t)));
        coef[1] = (float)(1.0 / sqrt(1.0 - pow(beta2_d, t)));
    }
    __syncthreads();
    const float beta1 = (float)beta1_d, beta2 = (float)beta2_d, omb1 = (float)(1.0 - beta1_d), omb2 = (float)(1.0 - beta2_d);
    const float step_size = coef[0], inv_sqrt_bc2 = coef[1];
    const long long n4 = d.n / 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(d.p) | reinterpret_cast<uintptr_t>(d.g) | reinterpret_cast<uintptr_t>(d.m) |
                       reinterpret_cast<uintptr_t>(d.v)) & 15) == 0;
    auto upd = [&](float& p, float g, float& m, float& v) {
        m = beta1 * m + omb1 * g;
        v = beta2 * v + omb2 * g * g;
        p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
    };
    if (vec) {
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
            float4 p = reinterpret_cast<float4*>(d.p)[i], m = reinterpret_cast<float4*>(d.m)[i], v = reinterpret_cast<float4*>(d.v)[i];
            const float4 g = __ldg(reinterpret_cast<const float4*>(d.g) + i);
            upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
            reinterpret_cast<float4*>(d.p)[i] = p;
            reinterpret_cast<float4*>(d.m)[i] = m;
            reinterpret_cast<float4*>(d.v)[i] = v;
        }
    }
    for (long long i = (vec ? n4 * 4 : 0) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < d.n; i += (long long)gridDim.x * blockDim.x)
        upd(d.p[i], d.g[i], d.m[i], d.v[i]);
}

// out[i] = in[i] * scale_ptr[0] * scale  (chain rule for a scalar upstream gradient living on the device)
template <typename T>
__global__ void scale_kernel(const T* __restrict__ in, T* __restrict__ out, long long n8, const float* __restrict__ scale_ptr,
                             float scale) {
    const float s = scale_ptr ? scale * __ldg(scale_ptr) : scale;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float f[8];
        V8<T>::load(in + i * 8, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] *= s;
        V8<T>::store(out + i * 8, f);
    }
}

}  // namespace fv

// =================================================================== C ABI
using namespace fv;
#define STREAM ((cudaStream_t)stream)

static int check_c8(const char* who, int C) {
    if (C < 8 || C % 8 || C > 2048) return fail(FV_ERR_UNSUPPORTED, "%s: channel count %d must be a multiple of 8 (<= 2048)", who, C);
    if (kThreads % (C / 8) && (C / 8) <= kThreads) return fail(FV_ERR_UNSUPPORTED, "%s: C/8 = %d must divide %d", who, C / 8, kThreads);
    return 0;
}

extern "C" __attribute__((visibility("default"))) int fv_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int N, int C, int H, int W, int Cp, void* stream) {
    if (!src || !dst || Cp % 8 || Cp < C) return fail(FV_ERR_ARG, "fv_nchw_to_nhwc: bad arguments (C=%d Cp=%d)", C, Cp);
    const long long items = (long long)N * H * W * (Cp / 8);
    if (dst_dtype == FV_DT_BF16)
        nchw_to_nhwc_kernel<__nv_bfloat16><<<grid_for(items), kThreads, 0, STREAM>>>(src, (__nv_bfloat16*)dst, N, C, H * W, Cp);
    else
        nchw_to_nhwc_kernel<float><<<grid_for(items), kThreads, 0, STREAM>>>(src, (float*)dst, N, C, H * W, Cp);
    FV_LAUNCH_CHECK("nchw_to_nhwc_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_nhwc_to_nchw(const void* src, int src_dtype, float* dst, int N, int C, int H, int W, int Cs, int accumulate,
                               void* stream) {
    if (!src || !dst || Cs % 8 || Cs < C) return fail(FV_ERR_ARG, "fv_nhwc_to_nchw: bad arguments (C=%d Cs=%d)", C, Cs);
    const long long items = (long long)N * H * W * ((C + 7) / 8);
    if (src_dtype == FV_DT_BF16)
        nhwc_to_nchw_kernel<__nv_bfloat16><<<grid_for(items), kThreads, 0, STREAM>>>((const __nv_bfloat16*)src, dst, N, C, H * W, Cs, accumulate);
    else
        nhwc_to_nchw_kernel<float><<<grid_for(items), kThreads, 0, STREAM>>>((const float*)src, dst, N, C, H * W, Cs, accumulate);
    FV_LAUNCH_CHECK("nhwc_to_nchw_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_weight_prep(const float* w, void* wf, void* wd, int Co, int Ci, int R, int S, int Co_pad, int Ci_pad, void* stream) {
    if (!w || (!wf && !wd) || Co_pad < Co || Ci_pad < Ci) return fail(FV_ERR_ARG, "fv_weight_prep: bad arguments");
    const long long items = (long long)Co_pad * Ci_pad * R * S;
    weight_prep_kernel<<<grid_for(items), kThreads, 0, STREAM>>>(w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd, Co, Ci, R, S, Co_pad, Ci_pad);
    FV_LAUNCH_CHECK("weight_prep_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_weight_prep_batched(const fv_prep_desc* table_dev, int n_layers, long long max_items, void* stream) {
    if (!table_dev || n_layers < 1 || max_items < 1) return fail(FV_ERR_ARG, "fv_weight_prep_batched: bad arguments");
    if (max_items >= (1LL << 31)) return fail(FV_ERR_UNSUPPORTED, "fv_weight_prep_batched: filter too large for 32-bit indexing");
    long long bx = (max_items + kThreads * 2 - 1) / (kThreads * 2);
    const long long cap = (long long)num_sms() * 24 / n_layers + 1;     // the whole table: about three resident waves (layers differ 1000x in size)
    if (bx > cap) bx = cap;
    weight_prep_batched_kernel<<<dim3((unsigned)bx, (unsigned)n_layers), kThreads, 0, STREAM>>>(table_dev);
    FV_LAUNCH_CHECK("weight_prep_batched_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_weight_prep_flat(const fv_prep_desc* table_dev, int n_layers, int total_blocks, void* stream) {
    if (!table_dev || n_layers < 1 || total_blocks < 1) return fail(FV_ERR_ARG, "fv_weight_prep_flat: bad arguments");
    weight_prep_flat_kernel<<<(unsigned)total_blocks, 256, 0, STREAM>>>(table_dev, n_layers);
    FV_LAUNCH_CHECK("weight_prep_flat_kernel");
    return FV_OK;
}
extern "C" __attribute__((visibility("default"))) int fv_weight_prep_block_items(void) { return kPrepBlockItems; }

// blocks of one layer in the grid of fv_weight_prep_tiled (desc.reserved = running sum of these)
extern "C" __attribute__((visibility("default"))) int fv_weight_prep_tiled_blocks(int kind, int Co_pad, int Ci_pad, int R, int S) {
    const int taps = kind == 0 ? R * S : (kind == 1 ? 9 : 16);
    if (taps > kPrepMaxTaps) return (int)(((long long)Co_pad * taps * Ci_pad + kPrepBlockItems - 1) / kPrepBlockItems);
    return ((Co_pad + kPrepTco - 1) / kPrepTco) * ((Ci_pad + kPrepTci - 1) / kPrepTci);
}
extern "C" __attribute__((visibility("default"))) int fv_weight_prep_tiled(const fv_prep_desc* table_dev, int n_layers, int total_blocks, void* stream) {
    if (!table_dev || n_layers < 1 || total_blocks < 1) return fail(FV_ERR_ARG, "fv_weight_prep_tiled: bad arguments");
    weight_prep_tile_kernel<<<(unsigned)total_blocks, 256, 0, STREAM>>>(table_dev, n_layers);
    FV_LAUNCH_CHECK("weight_prep_tile_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_wgrad_finish(const float* part, int splits, float* grad, int Co, int Ci, int R, int S, int Co_pad,
                                                                    int Ci_pad, int accumulate, void* stream) {
    if (!part || !grad || splits < 1) return fail(FV_ERR_ARG, "fv_wgrad_finish: bad arguments");
    wgrad_finish_kernel<<<grid_for((long long)Co * Ci_pad * R * S), kThreads, 0, STREAM>>>(part, grad, Co, Ci, R * S, Ci_pad, accumulate, splits,
                                                                                        (long long)Co_pad * R * S * Ci_pad);
    FV_LAUNCH_CHECK("wgrad_finish_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_slab_sum(const float* part, int slabs, long long slab_stride, float* out, long long n, int accumulate,
                                                                void* stream) {
    if (!part || !out || slabs < 1 || n < 1) return fail(FV_ERR_ARG, "fv_slab_sum: bad arguments");
    const bool aligned = n % 4 == 0 && slab_stride % 4 == 0 && ((reinterpret_cast<uintptr_t>(part) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (aligned && slabs >= 32 && n / 4 <= 32LL * 4 * fv::num_sms()) {         // at most ~4 blocks per SM: the serial walk would leave the GPU idle
        slab_sum_wide_kernel<<<(unsigned)((n / 4 + 31) / 32), 256, 0, STREAM>>>(reinterpret_cast<const float4*>(part), slabs, slab_stride / 4,
                                                                               reinterpret_cast<float4*>(out), n / 4, accumulate);
        FV_LAUNCH_CHECK("slab_sum_wide_kernel");
        return FV_OK;
    }
    slab_sum_kernel<<<grid_for(n / 4 + 1), kThreads, 0, STREAM>>>(part, slabs, slab_stride, out, n, accumulate);
    FV_LAUNCH_CHECK("slab_sum_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_weight_prep_up(const float* w, void* wx2, void* ws2, int Co, int Ci, int Co_pad, int Ci_pad, void* stream) {
    if (!w || (!wx2 && !ws2) || Co_pad < Co || Ci_pad < Ci) return fail(FV_ERR_ARG, "fv_weight_prep_up: bad arguments");
    weight_prep_up_kernel<<<grid_for(16LL * Co_pad * Ci_pad), kThreads, 0, STREAM>>>(w, (__nv_bfloat16*)wx2, (__nv_bfloat16*)ws2, Co, Ci, Co_pad, Ci_pad);
    FV_LAUNCH_CHECK("weight_prep_up_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_weight_prep_s2(const float* w, void* wf, void* wx2, int Co, int Ci, int Co_pad, int Ci_pad, void* stream) {
    if (!w || (!wf && !wx2) || Co_pad < Co || Ci_pad < Ci) return fail(FV_ERR_ARG, "fv_weight_prep_s2: bad arguments");
    weight_prep_s2_kernel<<<grid_for(16LL * Co_pad * Ci_pad), kThreads, 0, STREAM>>>(w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wx2, Co, Ci, Co_pad, Ci_pad);
    FV_LAUNCH_CHECK("weight_prep_s2_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_wgrad_finish_up(const float* part, int splits, float* grad, int Co, int Ci, int Co_pad, int Ci_pad,
                                                                       int accumulate, void* stream) {
    if (!part || !grad || splits < 1) return fail(FV_ERR_ARG, "fv_wgrad_finish_up: bad arguments");
    wgrad_finish_up_kernel<<<grid_for((long long)Co * Ci_pad), kThreads, 0, STREAM>>>(part, grad, Co, Ci, Co_pad, Ci_pad, accumulate, splits,
                                                                                       16LL * Co_pad * Ci_pad);
    FV_LAUNCH_CHECK("wgrad_finish_up_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_bilinear_resize(const float* x, float* out, int N, int C, int H, int W, int Ho, int Wo, void* stream) {
    if (!x || !out || N < 1 || C < 1 || H < 1 || W < 1 || Ho < 1 || Wo < 1) return fail(FV_ERR_ARG, "fv_bilinear_resize: bad arguments");
    bilinear_resize_kernel<<<grid_for((long long)N * C * Ho * Wo), kThreads, 0, STREAM>>>(x, out, (long long)N * C, H, W, Ho, Wo, (float)H / (float)Ho,
                                                                                         (float)W / (float)Wo);
    FV_LAUNCH_CHECK("bilinear_resize_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_demod_fwd(const float* w, float* weff, float* inv_norm, int Co, int K, float gain, int demod, void* stream) {
    if (!w || !weff || Co < 1 || K < 1) return fail(FV_ERR_ARG, "fv_demod_fwd: bad arguments");
    demod_fwd_kernel<<<Co, kThreads, 0, STREAM>>>(w, weff, inv_norm, K, gain, demod);
    FV_LAUNCH_CHECK("demod_fwd_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_demod_bwd(const float* w, const float* inv_norm, const float* dweff, float* dw, int Co, int K, float gain,
                                                                 int demod, void* stream) {
    if (!w || !dweff || !dw || (demod && !inv_norm) || Co < 1 || K < 1) return fail(FV_ERR_ARG, "fv_demod_bwd: bad arguments");
    demod_bwd_kernel<<<Co, kThreads, 0, STREAM>>>(w, inv_norm, dweff, dw, K, gain, demod);
    FV_LAUNCH_CHECK("demod_bwd_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_act_bwd(const void* out, const void* g, void* dy, long long n, int act, void* stream) {
    if (!out || !g || !dy || n < 8 || n % 8) return fail(FV_ERR_ARG, "fv_act_bwd: element count %lld must be a positive multiple of 8", n);
    act_bwd_kernel<<<grid_for(n / 8), kThreads, 0, STREAM>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)g, (__nv_bfloat16*)dy, n / 8, act);
    FV_LAUNCH_CHECK("act_bwd_kernel");
    return FV_OK;
}

static int in_check(const char* who, int N, int H, int W, int C) {
    if (N < 1 || H < 1 || W < 1 || C < 8 || C % 8) return fail(FV_ERR_UNSUPPORTED, "%s: N=%d H=%d W=%d C=%d (C must be a multiple of 8)", who, N, H, W, C);
    if ((long long)N * H * W * (C / 8) >= (1LL << 40)) return fail(FV_ERR_UNSUPPORTED, "%s: tensor too large", who);
    return 0;
}

extern "C" __attribute__((visibility("default"))) int fv_in_stats(const void* y, float* stat, int N, int H, int W, int C, float eps, void* stream) {
    if (!y || !stat) return fail(FV_ERR_ARG, "fv_in_stats: null pointer");
    if (int e = in_check("fv_in_stats", N, H, W, C)) return e;
    in_reduce_kernel<0><<<dim3((C + 63) / 64, N), 256, 0, STREAM>>>((const __nv_bfloat16*)y, nullptr, nullptr, nullptr, nullptr, stat, H * W, C, 0, eps);
    FV_LAUNCH_CHECK("in_reduce_kernel<stats>");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_in_act_fwd(const void* y, const float* stat, const float* gamma, const float* beta, void* out, int N, int H,
                                                                  int W, int C, int act, void* stream) {
    if (!y || !stat || !gamma || !beta || !out) return fail(FV_ERR_ARG, "fv_in_act_fwd: null pointer");
    if (int e = in_check("fv_in_act_fwd", N, H, W, C)) return e;
    in_apply_kernel<0><<<grid_for((long long)N * H * W * (C / 8)), 256, 0, STREAM>>>((const __nv_bfloat16*)y, nullptr, stat, nullptr, gamma, beta,
                                                                                     (__nv_bfloat16*)out, N, H * W, C, act);
    FV_LAUNCH_CHECK("in_apply_kernel<fwd>");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_in_bwd_sums(const void* y, const void* g, const float* stat, const float* gamma, const float* beta,
                                                                   float* sums, int N, int H, int W, int C, int act, void* stream) {
    if (!y || !g || !stat || !gamma || !beta || !sums) return fail(FV_ERR_ARG, "fv_in_bwd_sums: null pointer");
    if (int e = in_check("fv_in_bwd_sums", N, H, W, C)) return e;
    in_reduce_kernel<1><<<dim3((C + 63) / 64, N), 256, 0, STREAM>>>((const __nv_bfloat16*)y, (const __nv_bfloat16*)g, stat, gamma, beta, sums, H * W, C,
                                                                     act, 0.f);
    FV_LAUNCH_CHECK("in_reduce_kernel<bwd>");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_in_bwd_apply(const void* y, const void* g, const float* stat, const float* sums, const float* gamma,
                                                                    const float* beta, void* dy, int N, int H, int W, int C, int act, void* stream) {
    if (!y || !g || !stat || !sums || !gamma || !beta || !dy) return fail(FV_ERR_ARG, "fv_in_bwd_apply: null pointer");
    if (int e = in_check("fv_in_bwd_apply", N, H, W, C)) return e;
    in_apply_kernel<1><<<grid_for((long long)N * H * W * (C / 8)), 256, 0, STREAM>>>((const __nv_bfloat16*)y, (const __nv_bfloat16*)g, stat, sums, gamma, beta,
                                                                                     (__nv_bfloat16*)dy, N, H * W, C, act);
    FV_LAUNCH_CHECK("in_apply_kernel<bwd>");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) long long fv_reduce_ws_bytes(void) {
    return (long long)det_reduce_ws_bytes(kRedGroup * kRedMaxGroups, 1024, sizeof(float));
}
static int check_ws(const char* who, const void* ws, int grid, int n_elems, int elem_size) {
    if (!ws) return fail(FV_ERR_ARG, "%s: null reduction workspace (fv_reduce_ws_bytes() bytes, zero-initialised once)", who);
    if (grid > kRedGroup * kRedMaxGroups || (long long)n_elems * elem_size > 4096)
        return fail(FV_ERR_INTERNAL, "%s: reduction of %d x %d-byte elements over %d blocks exceeds the workspace layout", who, n_elems, elem_size, grid);
    return 0;
}

static int reduce_geometry(int C, long long P, int& grid, size_t& shmem, int resident = 2, int rows_per_thread = 16, int big_waves = 8) {
    const int rpi = kThreads / (C / 8) > 0 ? kThreads / (C / 8) : 1;
    // Reductions end every block with a shared-memory pass and 2C global atomics: at least 16 rows per thread, so that a
    // small tensor does not pay for a thousand blocks' worth of atomics on the same 2C addresses (14 us for 4 MB in round
    // 1).  Small tensors: at most one resident wave (the fixed per-block cost dominates); large tensors: 8 blocks per SM,
    // measured ~5 % faster there than a single wave.
    const long long row_blocks = (P + rpi - 1) / rpi;
    long long blocks = (row_blocks + rows_per_thread - 1) / rows_per_thread;
    static const int big_rows = getenv("FV_REDUCE_BIG") ? atoi(getenv("FV_REDUCE_BIG")) : 64;
    const long long cap = (long long)num_sms() * (row_blocks > (long long)num_sms() * big_rows ? big_waves : resident);
    grid = (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
    shmem = ((size_t)2 * rpi * C + 4 * (size_t)C) * sizeof(float);   // row partials | block vector [2C] | totals [2C]
    return rpi;
}

static unsigned long long xrank_timeout_ns_glue() {
    static unsigned long long cached = 0;
    if (!cached) {
        const char* v = getenv("FACEVAE_XRANK_TIMEOUT_S");
        double s_ = v ? atof(v) : 600.0;
        if (!(s_ > 0)) s_ = 600.0;
        cached = (unsigned long long)(s_ * 1e9);
    }
    return cached;
}

static int bn_stats_impl(const void* y, int dtype, float* sums, long long P, int C, void* ws, const XrankArgs& xr, void* stream) {
    if (!y || !sums) return fail(FV_ERR_ARG, "fv_bn_stats: null pointer");
    if (int e = check_c8("fv_bn_stats", C)) return e;
    int grid; size_t sh;
    // 51 registers: four 256-thread blocks per SM; one resident wave for every size (measured with the ordered cross-block
    // reduction: 63 / 38 / 25 us -> 59 / 33 / 21 us on the 268 / 134 / 67 MB tensors against two waves of smaller blocks)
    reduce_geometry(C, P, grid, sh, 4, 16, 4);
    if (int e = check_ws("fv_bn_stats", ws, grid, 2 * C, 4)) return e;
    if (dtype == FV_DT_BF16)
        bn_stats_kernel<__nv_bfloat16><<<grid, kThreads, sh, STREAM>>>((const __nv_bfloat16*)y, sums, P, C, ws, xr);
    else
        bn_stats_kernel<float><<<grid, kThreads, sh, STREAM>>>((const float*)y, sums, P, C, ws, xr);
    FV_LAUNCH_CHECK("bn_stats_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_bn_stats(const void* y, int dtype, float* sums, long long P, int C, void* ws, void* stream) {
    XrankArgs xr{};
    return bn_stats_impl(y, dtype, sums, P, C, ws, xr, stream);
}

// fv_bn_stats + fv_bn_finalize_xrank(mode 0) in ONE launch: the block that ends up with this rank's sums pushes them to the
// peers, gathers theirs and writes stat[4][C] (+ running statistics).  `sums` still receives the LOCAL sums.
extern "C" __attribute__((visibility("default"))) int fv_bn_stats_xrank(const void* y, int dtype, float* sums, long long P, int C, void* ws, void* peer_bufs_dev,
                                                                      int rank, int world, void* epoch_ctr, double count, const float* gamma,
                                                                      const float* beta, float* running_mean, float* running_var, float momentum,
                                                                      float eps, float* stat, void* stream) {
    if (!peer_bufs_dev || !epoch_ctr || !gamma || !beta || !stat || count <= 0 || world < 1 || world > kXMaxWorld || rank < 0 || rank >= world || 2 * C > kXRow)
        return fail(FV_ERR_ARG, "fv_bn_stats_xrank: bad arguments (rank %d / world %d, C=%d)", rank, world, C);
    XrankArgs xr{sums, reinterpret_cast<unsigned long long* const*>(peer_bufs_dev), rank, world, reinterpret_cast<unsigned long long*>(epoch_ctr), C, 0,
                 count, gamma, beta, running_mean, running_var, momentum, eps, stat, nullptr, nullptr, 0, xrank_timeout_ns_glue()};
    return bn_stats_impl(y, dtype, sums, P, C, ws, xr, stream);
}

extern "C" __attribute__((visibility("default"))) int fv_bn_finalize(const float* sums, double count, const float* gamma, const float* beta, float* running_mean,
                              float* running_var, float momentum, float eps, float* stat, int C, void* stream) {
    if (!sums || !gamma || !beta || !stat || count <= 0) return fail(FV_ERR_ARG, "fv_bn_finalize: bad arguments");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(sums, count, gamma, beta, running_mean, running_var, momentum, eps, stat, C);
    FV_LAUNCH_CHECK("bn_finalize_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                                 float eps, float* stat, int C, void* stream) {
    if (!gamma || !beta || !running_mean || !running_var || !stat) return fail(FV_ERR_ARG, "fv_bn_eval_affine: null pointer");
    bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(gamma, beta, running_mean, running_var, eps, stat, C);
    FV_LAUNCH_CHECK("bn_eval_affine_kernel");
    return FV_OK;
}

static int bn_act_fwd_impl(const void* y, int in_dtype, const float* stat, void* out, int out_dtype, int nchw_out, int N, int H, int W, int C,
                           int mode, int act, const BnFin& fin, void* stream);

extern "C" __attribute__((visibility("default"))) int fv_bn_act_fwd(const void* y, int in_dtype, const float* stat, void* out, int out_dtype, int nchw_out, int N, int H,
                             int W, int C, int mode, int act, void* stream) {
    if (!stat) return fail(FV_ERR_ARG, "fv_bn_act_fwd: null pointer");
    BnFin fin{};
    return bn_act_fwd_impl(y, in_dtype, stat, out, out_dtype, nchw_out, N, H, W, C, mode, act, fin, stream);
}

extern "C" __attribute__((visibility("default"))) int fv_bn_act_fwd_fin(const void* y, int in_dtype, const float* sums, double count, const float* gamma,
                                 const float* beta, float* running_mean, float* running_var, float momentum, float eps, float* stat_out,
                                 void* out, int out_dtype, int nchw_out, int N, int H, int W, int C, int mode, int act, void* stream) {
    if (!sums || !gamma || !beta || !stat_out || count <= 0) return fail(FV_ERR_ARG, "fv_bn_act_fwd_fin: bad arguments");
    BnFin fin{sums, count, gamma, beta, running_mean, running_var, momentum, eps, stat_out};
    return bn_act_fwd_impl(y, in_dtype, stat_out, out, out_dtype, nchw_out, N, H, W, C, mode, act, fin, stream);
}

static int bn_act_fwd_impl(const void* y, int in_dtype, const float* stat, void* out, int out_dtype, int nchw_out, int N, int H, int W, int C,
                           int mode, int act, const BnFin& fin, void* stream) {
    if (!y || !stat || !out) return fail(FV_ERR_ARG, "fv_bn_act_fwd: null pointer");
    if (C % 8 || 256 % (C / 8)) return fail(FV_ERR_UNSUPPORTED, "fv_bn_act_fwd: C=%d must be 8 * (a divisor of 256)", C);
    if ((long long)N * H * W * (C / 8) >= (1LL << 31)) return fail(FV_ERR_UNSUPPORTED, "fv_bn_act_fwd: tensor too large for 32-bit indexing");
    if (mode == FV_MODE_POOL && ((H | W) & 1)) return fail(FV_ERR_UNSUPPORTED, "fv_bn_act_fwd: pooling needs even H, W");
    if (nchw_out && (out_dtype != FV_DT_F32 || mode == FV_MODE_UP)) return fail(FV_ERR_UNSUPPORTED, "fv_bn_act_fwd: NCHW output is fp32, no upsample");
    const int Ho = mode == FV_MODE_POOL ? H / 2 : H, Wo = mode == FV_MODE_POOL ? W / 2 : W;
    const int grid = grid_for((long long)N * Ho * Wo * (C / 8), kThreads, 3);   // 79 registers: three blocks per SM
#define LAUNCH(TI, TO) do { \
        if (mode == FV_MODE_POOL) bn_act_fwd_kernel<TI, TO, FV_MODE_POOL><<<grid, kThreads, 0, STREAM>>>((const TI*)y, stat, (TO*)out, N, H, W, C, act, nchw_out, fin); \
        else if (mode == FV_MODE_UP) bn_act_fwd_kernel<TI, TO, FV_MODE_UP><<<grid, kThreads, 0, STREAM>>>((const TI*)y, stat, (TO*)out, N, H, W, C, act, nchw_out, fin); \
        else bn_act_fwd_kernel<TI, TO, FV_MODE_NONE><<<grid, kThreads, 0, STREAM>>>((const TI*)y, stat, (TO*)out, N, H, W, C, act, nchw_out, fin); } while (0)
    if (in_dtype == FV_DT_BF16 && out_dtype == FV_DT_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
    else if (in_dtype == FV_DT_BF16) LAUNCH(__nv_bfloat16, float);
    else if (out_dtype == FV_DT_BF16) LAUNCH(float, __nv_bfloat16);
    else LAUNCH(float, float);
#undef LAUNCH
    FV_LAUNCH_CHECK("bn_act_fwd_kernel");
    return FV_OK;
}

static int bn_act_bwd_reduce_impl(const void* y, int y_dtype, const void* g, int g_dtype, int g_nchw, const float* stat,
                                    float* sums, int N, int H, int W, int C, int mode, int act, void* ws, const XrankArgs& xr, void* stream) {
    if (!y || !g || !stat || !sums) return fail(FV_ERR_ARG, "fv_bn_act_bwd_reduce: null pointer");
    if (int e = check_c8("fv_bn_act_bwd_reduce", C)) return e;
    if ((long long)N * H * W >= (1LL << 31)) return fail(FV_ERR_UNSUPPORTED, "fv_bn_act_bwd_reduce: tensor too large for 32-bit indexing");
    if (g_nchw && (g_dtype != FV_DT_F32 || mode == FV_MODE_UP)) return fail(FV_ERR_UNSUPPORTED, "fv_bn_act_bwd_reduce: NCHW g is fp32, no upsample");
    int grid; size_t sh;
    // one resident wave also for large tensors: every block ends with 2C atomics on the same addresses (1184 blocks x 128
    // channels = 150 k serialised atomics cost more than the second wave's tail gains)
    reduce_geometry(C, (long long)N * H * W, grid, sh, 2, 16, 2);
    if (int e = check_ws("fv_bn_act_bwd_reduce", ws, grid, 2 * C, 4)) return e;
#define LAUNCH3(TY, TG, M, GNF) bn_act_bwd_reduce_kernel<TY, TG, M, GNF><<<grid, kThreads, sh, STREAM>>>((const TY*)y, (const TG*)g, stat, sums, N, H, W, C, act, ws, xr)
#define LAUNCH2(TY, TG) do { \
        if (mode == FV_MODE_POOL) { if (g_nchw) LAUNCH3(TY, TG, FV_MODE_POOL, true); else LAUNCH3(TY, TG, FV_MODE_POOL, false); } \
        else if (mode == FV_MODE_UP) LAUNCH3(TY, TG, FV_MODE_UP, false); \
        else { if (g_nchw) LAUNCH3(TY, TG, FV_MODE_NONE, true); else LAUNCH3(TY, TG, FV_MODE_NONE, false); } } while (0)
    if (y_dtype == FV_DT_BF16 && g_dtype == FV_DT_BF16) LAUNCH2(__nv_bfloat16, __nv_bfloat16);
    else if (y_dtype == FV_DT_BF16) LAUNCH2(__nv_bfloat16, float);
    else if (g_dtype == FV_DT_BF16) LAUNCH2(float, __nv_bfloat16);
    else LAUNCH2(float, float);
#undef LAUNCH2
#undef LAUNCH3
    FV_LAUNCH_CHECK("bn_act_bwd_reduce_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_bn_act_bwd_reduce(const void* y, int y_dtype, const void* g, int g_dtype, int g_nchw, const float* stat,
                                    float* sums, int N, int H, int W, int C, int mode, int act, void* ws, void* stream) {
    XrankArgs xr{};
    return bn_act_bwd_reduce_impl(y, y_dtype, g, g_dtype, g_nchw, stat, sums, N, H, W, C, mode, act, ws, xr, stream);
}

// fv_bn_act_bwd_reduce + fv_bn_finalize_xrank(mode 1) in ONE launch: coef[2][C] from the cross-rank sums, dgamma / dbeta from the
// local ones (`sums` still receives the local sums).
extern "C" __attribute__((visibility("default"))) int fv_bn_act_bwd_reduce_xrank(const void* y, int y_dtype, const void* g, int g_dtype, int g_nchw,
                                                                               const float* stat, float* sums, int N, int H, int W, int C, int mode,
                                                                               int act, void* ws, void* peer_bufs_dev, int rank, int world,
                                                                               void* epoch_ctr, double count, float* coef, float* dgamma, float* dbeta,
                                                                               void* stream) {
    if (!peer_bufs_dev || !epoch_ctr || !coef || count <= 0 || world < 1 || world > kXMaxWorld || rank < 0 || rank >= world || 2 * C > kXRow)
        return fail(FV_ERR_ARG, "fv_bn_act_bwd_reduce_xrank: bad arguments (rank %d / world %d, C=%d)", rank, world, C);
    XrankArgs xr{sums, reinterpret_cast<unsigned long long* const*>(peer_bufs_dev), rank, world, reinterpret_cast<unsigned long long*>(epoch_ctr), C, 1,
                 count, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, coef, dgamma, dbeta, 0, xrank_timeout_ns_glue()};
    return bn_act_bwd_reduce_impl(y, y_dtype, g, g_dtype, g_nchw, stat, sums, N, H, W, C, mode, act, ws, xr, stream);
}

extern "C" __attribute__((visibility("default"))) int fv_bn_bwd_finalize(const float* sums_local, const float* sums_global, double count, float* dgamma, float* dbeta,
                                  float* coef, int C, int accumulate, void* stream) {
    if (!sums_local || !sums_global || !coef || count <= 0) return fail(FV_ERR_ARG, "fv_bn_bwd_finalize: bad arguments");
    bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(sums_local, sums_global, count, dgamma, dbeta, coef, C, accumulate);
    FV_LAUNCH_CHECK("bn_bwd_finalize_kernel");
    return FV_OK;
}

static int bn_act_bwd_apply_impl(const void* y, int y_dtype, const void* g, int g_dtype, int g_nchw, const float* stat, const float* coef,
                                 const void* add, void* dy, int N, int H, int W, int C, int mode, int act, const float* fin_sums, double fin_count,
                                 float* dgamma, float* dbeta, void* stream);

extern "C" __attribute__((visibility("default"))) int fv_bn_act_bwd_apply(const void* y, int y_dtype, const void* g, int g_dtype, int g_nchw, const float* stat,
                                   const float* coef, const void* add, void* dy, int N, int H, int W, int C, int mode, int act,
                                   void* stream) {
    if (!coef) return fail(FV_ERR_ARG, "fv_bn_act_bwd_apply: null pointer");
    return bn_act_bwd_apply_impl(y, y_dtype, g, g_dtype, g_nchw, stat, coef, add, dy, N, H, W, C, mode, act, nullptr, 1.0, nullptr, nullptr, stream);
}

extern "C" __attribute__((visibility("default"))) int fv_bn_act_bwd_apply_fin(const void* y, int y_dtype, const void* g, int g_dtype, int g_nchw, const float* stat,
                                       const float* sums, double count, float* dgamma, float* dbeta, const void* add, void* dy, int N,
                                       int H, int W, int C, int mode, int act, void* stream) {
    if (!sums || count <= 0) return fail(FV_ERR_ARG, "fv_bn_act_bwd_apply_fin: bad arguments");
    return bn_act_bwd_apply_impl(y, y_dtype, g, g_dtype, g_nchw, stat, sums, add, dy, N, H, W, C, mode, act, sums, count, dgamma, dbeta, stream);
}

static int bn_act_bwd_apply_impl(const void* y, int y_dtype, const void* g, int g_dtype, int g_nchw, const float* stat, const float* coef,
                                 const void* add, void* dy, int N, int H, int W, int C, int mode, int act, const float* fin_sums, double fin_count,
                                 float* dgamma, float* dbeta, void* stream) {
    if (!y || !g || !stat || !coef || !dy) return fail(FV_ERR_ARG, "fv_bn_act_bwd_apply: null pointer");
    if (C % 8 || 256 % (C / 8)) return fail(FV_ERR_UNSUPPORTED, "fv_bn_act_bwd_apply: C=%d must be 8 * (a divisor of 256)", C);
    if ((long long)N * H * W * (C / 8) >= (1LL << 31)) return fail(FV_ERR_UNSUPPORTED, "fv_bn_act_bwd_apply: tensor too large for 32-bit indexing");
    if (g_nchw && (g_dtype != FV_DT_F32 || mode == FV_MODE_UP)) return fail(FV_ERR_UNSUPPORTED, "fv_bn_act_bwd_apply: NCHW g is fp32, no upsample");
    if (int e = check_c8("fv_bn_act_bwd_apply", C)) return e;
    int grid; size_t sh_unused;
    reduce_geometry(C, (long long)N * H * W, grid, sh_unused, 2, 4);   // no block tail here: one batch of rows per thread
#define LAUNCH4(TY, TG, M, GNF, AD) bn_act_bwd_apply_kernel<TY, TG, M, GNF, AD><<<grid, kThreads, 0, STREAM>>>((const TY*)y, (const TG*)g, stat, coef, (const __nv_bfloat16*)add, (__nv_bfloat16*)dy, N, H, W, C, act, fin_sums, fin_count, dgamma, dbeta)
#define LAUNCH3(TY, TG, M, GNF) do { if (add) LAUNCH4(TY, TG, M, GNF, true); else LAUNCH4(TY, TG, M, GNF, false); } while (0)
#define LAUNCH2(TY, TG) do { \
        if (mode == FV_MODE_POOL) { if (g_nchw) LAUNCH3(TY, TG, FV_MODE_POOL, true); else LAUNCH3(TY, TG, FV_MODE_POOL, false); } \
        else if (mode == FV_MODE_UP) LAUNCH3(TY, TG, FV_MODE_UP, false); \
        else { if (g_nchw) LAUNCH3(TY, TG, FV_MODE_NONE, true); else LAUNCH3(TY, TG, FV_MODE_NONE, false); } } while (0)
    if (y_dtype == FV_DT_BF16 && g_dtype == FV_DT_BF16) LAUNCH2(__nv_bfloat16, __nv_bfloat16);
    else if (y_dtype == FV_DT_BF16) LAUNCH2(__nv_bfloat16, float);
    else if (g_dtype == FV_DT_BF16) LAUNCH2(float, __nv_bfloat16);
    else LAUNCH2(float, float);
#undef LAUNCH2
#undef LAUNCH3
#undef LAUNCH4
    FV_LAUNCH_CHECK("bn_act_bwd_apply_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_colsum(const void* y, float* sums, long long P, int C, void* ws, void* stream) {
    if (!y || !sums) return fail(FV_ERR_ARG, "fv_colsum: null pointer");
    if (C % 8 || C < 8) return fail(FV_ERR_UNSUPPORTED, "fv_colsum: channel count %d must be a multiple of 8", C);
    // rows wider than 1024 channels (the 16 -> 256*16 mid_conv of EFE_conv5) are walked in column chunks of <= 1024 (the
    // reduction workspace holds vectors of up to 1024 floats)
    for (int c0 = 0; c0 < C; c0 += 1024) {
        const int cn = C - c0 < 1024 ? C - c0 : 1024;
        if (int e = check_c8("fv_colsum", cn)) return e;
        int grid; size_t sh;
        reduce_geometry(cn, P, grid, sh);
        if (int e = check_ws("fv_colsum", ws, grid, cn, 4)) return e;
        colsum_kernel<<<grid, kThreads, sh, STREAM>>>((const __nv_bfloat16*)y + c0, sums + c0, P, cn, (long long)C, ws);
        FV_LAUNCH_CHECK("colsum_kernel");
    }
    return FV_OK;
}

static int reparam_blocks(int N, int Dz) {
    int bx = (Dz / 4 + kThreads - 1) / kThreads;
    const int cap = (num_sms() * 8 + N - 1) / N;
    if (bx > cap) bx = cap;
    return bx < 1 ? 1 : bx;
}
// partial KL sums per row that fv_reparam_kl_fwd writes (kl_part has N * this many floats)
extern "C" __attribute__((visibility("default"))) int fv_reparam_kl_parts(int N, int Dz) { return (N < 1 || Dz < 4) ? 1 : reparam_blocks(N, Dz); }

static bool misaligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }

extern "C" __attribute__((visibility("default"))) int fv_reparam_kl_fwd(const float* mu, const float* logstd, long long row_stride, const float* eps, float* z,
                                 float* kl_part, int N, int Dz, void* stream) {
    if (!mu || !logstd || Dz % 4 || N < 1) return fail(FV_ERR_ARG, "fv_reparam_kl_fwd: bad arguments (Dz=%d must be a multiple of 4)", Dz);
    if (row_stride % 4 || misaligned16(mu) || misaligned16(logstd) || misaligned16(eps) || misaligned16(z))
        return fail(FV_ERR_ARG, "fv_reparam_kl_fwd: rows must be 16-byte aligned");
    reparam_kl_fwd_kernel<<<dim3(reparam_blocks(N, Dz), N), kThreads, 0, STREAM>>>(mu, logstd, row_stride, eps, z, kl_part, Dz);
    FV_LAUNCH_CHECK("reparam_kl_fwd_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_reparam_kl_bwd(const float* mu, const float* logstd, long long row_stride, const float* eps, const float* dz,
                                 const float* dmu_ext, const float* dls_ext, float kscale, const float* kscale_ptr, float* dmu,
                                 float* dls, long long out_stride, int N, int Dz, void* stream) {
    if (!mu || !logstd || !dmu || !dls || N < 1 || Dz % 4) return fail(FV_ERR_ARG, "fv_reparam_kl_bwd: bad arguments (Dz=%d must be a multiple of 4)", Dz);
    if (row_stride % 4 || out_stride % 4 || misaligned16(mu) || misaligned16(logstd) || misaligned16(eps) || misaligned16(dz) ||
        misaligned16(dmu_ext) || misaligned16(dls_ext) || misaligned16(dmu) || misaligned16(dls))
        return fail(FV_ERR_ARG, "fv_reparam_kl_bwd: rows must be 16-byte aligned");
    reparam_kl_bwd_kernel<<<dim3(reparam_blocks(N, Dz), N), kThreads, 0, STREAM>>>(mu, logstd, row_stride, eps, dz, dmu_ext, dls_ext, kscale, kscale_ptr,
                                                                                 dmu, dls, out_stride, Dz);
    FV_LAUNCH_CHECK("reparam_kl_bwd_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_recon_loss(const float* logits, const float* target, float* pred_out, float* grad_f32, void* grad_nhwc,
                             float* loss_sum, int N, int C, int H, int W, int Cp, int l1, int use_sigmoid, float gscale,
                             void* ws, void* stream) {
    if (!logits || !target || !loss_sum) return fail(FV_ERR_ARG, "fv_recon_loss: null pointer");
    if (grad_nhwc && (Cp % 8 || Cp < C || C > 16)) return fail(FV_ERR_UNSUPPORTED, "fv_recon_loss: NHWC gradient needs C <= 16 <= Cp, Cp %% 8 == 0");
    const bool vec = (H * W) % 4 == 0 && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(pred_out) |
                                           reinterpret_cast<uintptr_t>(grad_f32)) & 15) == 0;
    const int grid = grid_for((long long)N * H * W / (vec ? 4 : 1));
    if (int e = check_ws("fv_recon_loss", ws, grid, 1, 4)) return e;
    if (vec)
        recon_loss_kernel<4><<<grid, kThreads, 0, STREAM>>>(logits, target, pred_out, grad_f32, (__nv_bfloat16*)grad_nhwc, loss_sum, N, C, H * W, Cp, l1,
                                                            use_sigmoid, gscale, ws);
    else
        recon_loss_kernel<1><<<grid, kThreads, 0, STREAM>>>(logits, target, pred_out, grad_f32, (__nv_bfloat16*)grad_nhwc, loss_sum, N, C, H * W, Cp, l1,
                                                            use_sigmoid, gscale, ws);
    FV_LAUNCH_CHECK("recon_loss_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_recon_loss_flat(const float* a, const float* b, float* grad, float* loss_sum, long long E, int l1, float gscale,
                                  void* ws, void* stream) {
    if (!a || !b || !loss_sum || E < 1) return fail(FV_ERR_ARG, "fv_recon_loss_flat: bad arguments");
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(grad)) & 15)
        return fail(FV_ERR_ARG, "fv_recon_loss_flat: pointers must be 16-byte aligned");
    const int grid = grid_for(E / 4 + 1);
    if (int e = check_ws("fv_recon_loss_flat", ws, grid, 1, 4)) return e;
    recon_loss_flat_kernel<<<grid, kThreads, 0, STREAM>>>(a, b, grad, loss_sum, E, l1, gscale, ws);
    FV_LAUNCH_CHECK("recon_loss_flat_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_adam_multi(const fv_adam_desc* table_dev, int n_tensors, long long max_n, float lr, double beta1,
                                                                  double beta2, float eps, const float* step_dev, void* stream) {
    if (!table_dev || !step_dev || n_tensors < 1 || max_n < 1) return fail(FV_ERR_ARG, "fv_adam_multi: bad arguments");
    long long bx = (max_n / 4 + kThreads * 2 - 1) / (kThreads * 2);
    const long long cap = (long long)num_sms() * 16 / n_tensors + 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    adam_multi_kernel<<<dim3((unsigned)bx, (unsigned)n_tensors), kThreads, 0, STREAM>>>(table_dev, lr, beta1, beta2, eps, step_dev);
    FV_LAUNCH_CHECK("adam_multi_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_scale(const void* in, void* out, int dtype, long long n, const float* scale_ptr, float scale, void* stream) {
    if (!in || !out || n % 8) return fail(FV_ERR_ARG, "fv_scale: element count %lld must be a multiple of 8", n);
    if (dtype == FV_DT_BF16)
        scale_kernel<__nv_bfloat16><<<grid_for(n / 8), kThreads, 0, STREAM>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n / 8, scale_ptr, scale);
    else
        scale_kernel<float><<<grid_for(n / 8), kThreads, 0, STREAM>>>((const float*)in, (float*)out, n / 8, scale_ptr, scale);
    FV_LAUNCH_CHECK("scale_kernel");
    return FV_OK;
}
