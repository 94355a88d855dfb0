// Implicit-GEMM 2-D convolution (stride 1, "same" padding) for sm_100a: NHWC bf16 in, fp32 accumulate in TMEM.
//
//   Y[pixel, co] = sum_{r,s,ci} X[pixel + (r - pad, s - pad), ci] * Wt[co, (r*S + s)*Ci + ci]     (+ bias, + residual)
//
// GEMM view: M = N*H*W pixels (128 per tile), N = Co_pad (one UMMA N, <= 256), K = R*S*Ci.  Replaces the
// nn.Conv2d inside the reference's _ConvBlock (reference modules.py:15,32) -- forward, and (called with the
// rotated/transposed filter produced by fv_weight_prep) its data gradient.
//
// Structure (one persistent CTA per SM, 6 warps):
//   warp 0  : TMA producer.  Per K block it loads one filter tap of the activation tile with a 4-D tensor map over
//             (C, W, H, N): box = (KB channels, tw, th, tn) at (kc, w0 + s - pad, h0 + r - pad, n0).  Halo pixels
//             outside the image are zero-filled by the TMA unit, so padding costs no memory and no branches.
//             The matching [Co_pad x KB] slice of the K-major filter matrix comes through a 2-D map.
//   warp 1  : one thread issues tcgen05.mma (UMMA 128 x Co_pad x 16, bf16 -> fp32) into one of two TMEM
//             accumulator buffers and commits to the stage / accumulator mbarriers.
//   warps 2-5: epilogue.  tcgen05.ld their TMEM lane quarter (row = pixel), add bias / residual, convert and store
//             (NHWC bf16, NHWC fp32 or NCHW fp32); overlapped with the next tile's MMAs via the second buffer.
#include <cstdio>

#include "../../include/facevae_b200.h"
#include "fv_host.h"
#include "fv_ptx.cuh"

namespace fv {

struct ConvParams {
    int N, H, W, Ci, Co, Co_pad, R, S, pad;
    int tw, th, tn, tiles_w, tiles_h, tiles_n, num_tiles;
    int KB, kc_blocks, stages, row_bytes;
    int a_bytes, b_bytes, stage_stride;
    int out_mode, tmem_cols;
    const float* bias;
    const __nv_bfloat16* residual;
    void* out;
};

static constexpr int kConvThreads = 192;

__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_stride);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int k_iters = p.R * p.S * p.kc_blocks;
    const int tiles_per_img_group = p.tiles_w * p.tiles_h;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int tn_i = tile / tiles_per_img_group;
                const int rem = tile - tn_i * tiles_per_img_group;
                const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
                const int w0 = tw_i * p.tw, h0 = th_i * p.th, n0 = tn_i * p.tn;
                for (int r = 0; r < p.R; ++r)
                    for (int s = 0; s < p.S; ++s)
                        for (int kc = 0; kc < p.kc_blocks; ++kc, ++it) {
                            const uint32_t st = it % p.stages, ph = (it / p.stages) & 1;
                            mbar_wait(&empty[st], ph ^ 1);
                            uint8_t* a_dst = smem + (size_t)st * p.stage_stride;
                            uint8_t* b_dst = a_dst + p.a_bytes;
                            mbar_arrive_expect_tx(&full[st], (uint32_t)(p.a_bytes + p.b_bytes));
                            tma_load_4d(a_dst, &tmX, &full[st], kc * p.KB, w0 + s - p.pad, h0 + r - p.pad, n0);
                            tma_load_2d(b_dst, &tmW, &full[st], (r * p.S + s) * p.Ci + kc * p.KB, 0);
                        }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, p.Co_pad, 0, 0);
            const uint32_t layout = umma_layout_code(p.row_bytes);
            const uint32_t sbo = 8u * p.row_bytes;
            const int k_sub = p.KB / 16;
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
                mbar_wait(&tempty[acc], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.Co_pad;
                for (int ki = 0; ki < k_iters; ++ki, ++it) {
                    const uint32_t st = it % p.stages, ph = (it / p.stages) & 1;
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)st * p.stage_stride);
                    const uint32_t b_addr = a_addr + p.a_bytes;
                    for (int j = 0; j < k_sub; ++j) {
                        const uint64_t adesc = umma_smem_desc(a_addr + j * 32, 16, sbo, layout);
                        const uint64_t bdesc = umma_smem_desc(b_addr + j * 32, 16, sbo, layout);
                        tc_mma_f16(d_tmem, adesc, bdesc, idesc, (ki > 0 || j > 0) ? 1u : 0u);
                    }
                    tc_commit(&empty[st]);   // frees the smem stage once these MMAs have read it
                }
                tc_commit(&tfull[acc]);      // accumulator complete -> epilogue
            }
        }
    } else {
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;       // tile row == pixel within the tile
        const int w_l = row % p.tw, h_l = (row / p.tw) % p.th, n_l = row / (p.tw * p.th);
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
            const int tn_i = tile / tiles_per_img_group;
            const int rem = tile - tn_i * tiles_per_img_group;
            const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
            const int w = tw_i * p.tw + w_l, h = th_i * p.th + h_l, n = tn_i * p.tn + n_l;
            const bool valid = n < p.N;
            const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * (uint32_t)p.Co_pad;
            const size_t pix = ((size_t)n * p.H + h) * p.W + w;
            for (int c0 = 0; c0 < p.Co_pad; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);
                tmem_ld_wait();
                if (valid) {
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
                if (p.bias) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < p.Co) f[i] += __ldg(p.bias + c0 + i);
                }
                if (p.out_mode == FV_OUT_NCHW_F32) {
                    float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < p.Co) o[(((size_t)n * p.Co + c0 + i) * p.H + h) * p.W + w] = f[i];
                } else {
                    if (p.residual) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.Co_pad + c0);
                        const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            f[2 * i] += bf16_lo(rr[i]);
                            f[2 * i + 1] += bf16_hi(rr[i]);
                        }
                    }
                    if (p.out_mode == FV_OUT_NHWC_BF16) {
                        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.Co_pad + c0);
                        o[0] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                          pack_bf16(f[6], f[7]));
                        o[1] = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]),
                                          pack_bf16(f[14], f[15]));
                    } else {
                        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.Co_pad + c0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) o[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                    }
                }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

static int pick_tile(int N, int H, int W, int& tw, int& th, int& tn) {
    if (W >= 128) {
        if (W % 128) return 1;
        tw = 128; th = 1; tn = 1;
        return 0;
    }
    if (W < 1 || (W & (W - 1))) return 1;       // W < 128 must be a power of two
    tw = W;
    int rest = 128 / tw;
    if (H >= rest) {
        if (H % rest) return 1;
        th = rest; tn = 1;
        return 0;
    }
    if (H & (H - 1)) return 1;
    th = H;
    tn = rest / H;                               // tile spans several images; the tail is zero-filled / masked
    (void)N;
    return 0;
}

}  // namespace fv

extern "C" __attribute__((visibility("default"))) int fv_conv2d(const void* x, const void* w, const float* bias, const void* residual, void* y, int out_mode,
                         int N, int H, int W, int Ci, int Co, int Co_pad, int R, int S, int pad, void* stream) {
    using namespace fv;
    if (!x || !w || !y) return fail(FV_ERR_ARG, "fv_conv2d: null pointer");
    if (N < 1 || H < 1 || W < 1) return fail(FV_ERR_ARG, "fv_conv2d: bad shape N=%d H=%d W=%d", N, H, W);
    if (Ci % 16 || Ci < 16 || (Ci > 64 && Ci % 64) || (Ci < 64 && Ci != 16 && Ci != 32))
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d: Ci=%d must be 16, 32 or a multiple of 64 (pad the channels)", Ci);
    if (Co_pad % 16 || Co_pad < 16 || Co_pad > 256 || Co < 1 || Co > Co_pad)
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d: Co=%d Co_pad=%d (Co_pad must be a multiple of 16 in [16,256])", Co, Co_pad);
    if (R != S || (R != 1 && R != 3 && R != 5 && R != 7) || pad != (R - 1) / 2)
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d: only odd square filters with same padding (R=%d S=%d pad=%d)", R, S, pad);
    if (out_mode < 0 || out_mode > 2) return fail(FV_ERR_ARG, "fv_conv2d: out_mode %d", out_mode);
    if (out_mode == FV_OUT_NCHW_F32 && residual) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d: residual needs an NHWC output");
    ConvParams p{};
    p.N = N; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.Co_pad = Co_pad; p.R = R; p.S = S; p.pad = pad;
    if (pick_tile(N, H, W, p.tw, p.th, p.tn))
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d: H=%d W=%d not tileable (W multiple of 128, or W,H powers of two)", H, W);
    p.tiles_w = W / p.tw; p.tiles_h = H / p.th; p.tiles_n = (N + p.tn - 1) / p.tn;
    p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    p.KB = Ci >= 64 ? 64 : Ci;
    p.kc_blocks = Ci / p.KB;
    p.row_bytes = p.KB * 2;
    p.a_bytes = 128 * p.row_bytes;
    p.b_bytes = Co_pad * p.row_bytes;
    p.stage_stride = p.a_bytes + ((p.b_bytes + 1023) & ~1023);
    const int k_iters = R * S * p.kc_blocks;
    int stages = (200 * 1024) / p.stage_stride;
    if (stages > 8) stages = 8;
    if (stages > k_iters) stages = k_iters;
    if (stages < 2) stages = 2;
    p.stages = stages;
    p.out_mode = out_mode;
    int cols = 32;
    while (cols < 2 * Co_pad) cols <<= 1;
    p.tmem_cols = cols;
    p.bias = bias;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.out = y;

    CUtensorMap tmX, tmW;
    {
        uint64_t dims[4] = {(uint64_t)Ci, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)Ci * 2, (uint64_t)W * Ci * 2, (uint64_t)H * W * Ci * 2};
        uint32_t box[4] = {(uint32_t)p.KB, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
        if (int e = encode_tmap_bf16(&tmX, x, 4, dims, str, box, p.row_bytes)) return e;
    }
    {
        uint64_t dims[2] = {(uint64_t)R * S * Ci, (uint64_t)Co_pad};
        uint64_t str[1] = {(uint64_t)R * S * Ci * 2};
        uint32_t box[2] = {(uint32_t)p.KB, (uint32_t)Co_pad};
        if (int e = encode_tmap_bf16(&tmW, w, 2, dims, str, box, p.row_bytes)) return e;
    }
    const size_t smem = (size_t)p.stages * p.stage_stride + 1024 + 512;
    static bool attr_set = false;
    if (!attr_set) {
        FV_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    conv_igemm_kernel<<<grid, kConvThreads, smem, (cudaStream_t)stream>>>(tmX, tmW, p);
    FV_LAUNCH_CHECK("conv_igemm_kernel");
    return FV_OK;
}
