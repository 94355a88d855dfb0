// Weight gradient of the stride-1 "same" convolution as a tcgen05 GEMM with the pixels as the K dimension:
//
//   dW[co, tap, ci] = sum_pixels  X[pixel + tap_offset, ci] * dY[pixel, co]
//
// Both operands are NHWC, i.e. the contraction index (pixel) is the slow index in memory, so they are fed to the
// tensor core as MN-major operands straight from TMA boxes (no transpose pass):
//   A (M side) = shifted input windows.  One 128-row M tile stacks 128/cw chunks of cw = min(Ci,64) channels, each
//                chunk being one (filter tap, channel chunk) pair, so thin layers (Ci = 16/32/64) still fill M = 128.
//   B (N side) = dY, N = Co_pad.
//   D[(tap,ci), co] accumulates in TMEM over this CTA's share of the pixels (split-K across CTAs); the epilogue adds
//   it into the fp32 buffer dWacc[Co_pad][taps][Ci] with red.global.add.f32.
// Replaces the wgrad half of aten::convolution_backward behind nn.Conv2d (reference modules.py:15,32).
#include "../../include/facevae_b200.h"
#include "fv_host.h"
#include "fv_ptx.cuh"

namespace fv {

struct WgradParams {
    int N, H, W, Ci, Co_pad, R, S, pad, taps;
    int pw, ph, pn, tiles_w, tiles_h, tiles_n, num_pb;   // 64-pixel K blocks
    int cw, cpt, chunks_per_tap, total_chunks;           // A chunking
    int bw, b_chunks;                                    // B chunking
    int mt_total, mt_per_group, groups, splits, pb_per_split;
    int a_chunk_bytes, b_chunk_bytes, a_bytes, b_bytes, stage_stride, stages, tmem_cols;
    float* dw;
};

static constexpr int kWgradThreads = 192;

__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // align up with arithmetic on the array itself so the compiler keeps the shared address space (LDS/STS, not generic)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_stride);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x % p.groups, split = blockIdx.x / p.groups;
    const int mt0 = group * p.mt_per_group;
    const int mt_n = min(p.mt_per_group, p.mt_total - mt0);
    const int chunk0 = mt0 * p.cpt;
    const int n_chunks = min(mt_n * p.cpt, p.total_chunks - chunk0);
    const int pb0 = split * p.pb_per_split;
    const int pb1 = min(pb0 + p.pb_per_split, p.num_pb);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDY);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tfull, 1);
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

    if (warp == 0) {
        {
            const bool leader = elect_one_sync();      // role loops stay warp-uniform; only the issue is predicated
            const uint32_t tx = (uint32_t)(n_chunks * p.a_chunk_bytes + p.b_bytes);
            uint32_t st = 0, phs = 0;
            // pixel-block coordinates advance incrementally (no divisions in the steady state)
            int tn_i = pb0 / (p.tiles_w * p.tiles_h);
            int rem = pb0 - tn_i * (p.tiles_w * p.tiles_h);
            int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
            for (int pb = pb0; pb < pb1; ++pb) {
                const int w0 = tw_i * p.pw, h0 = th_i * p.ph, n0 = tn_i * p.pn;
                mbar_wait(&empty[st], phs ^ 1);
                uint8_t* a_dst = smem + (size_t)st * p.stage_stride;
                uint8_t* b_dst = a_dst + p.a_bytes;
                if (leader) mbar_arrive_expect_tx(&full[st], tx);
                for (int bc = 0; bc < p.b_chunks; ++bc)
                    if (leader) tma_load_4d(b_dst + (size_t)bc * p.b_chunk_bytes, &tmDY, &full[st], bc * p.bw, w0, h0, n0);
                int tap = chunk0 / p.chunks_per_tap, cc = chunk0 - tap * p.chunks_per_tap;
                int r = tap / p.S, sx = tap - r * p.S;
                for (int j = 0; j < n_chunks; ++j) {
                    if (leader)
                        tma_load_4d(a_dst + (size_t)j * p.a_chunk_bytes, &tmX, &full[st], cc * p.cw, w0 + sx - p.pad,
                                    h0 + r - p.pad, n0);
                    if (++cc == p.chunks_per_tap) {
                        cc = 0;
                        if (++sx == p.S) { sx = 0; ++r; }
                    }
                }
                if (++st == (uint32_t)p.stages) { st = 0; phs ^= 1; }
                if (++tw_i == p.tiles_w) {
                    tw_i = 0;
                    if (++th_i == p.tiles_h) { th_i = 0; ++tn_i; }
                }
            }
        }
    } else if (warp == 1) {
        {
            const bool leader = elect_one_sync();
            const uint32_t idesc = umma_idesc_bf16(128, p.Co_pad, 1, 1);
            const uint32_t a_layout = umma_layout_code(p.cw * 2), b_layout = umma_layout_code(p.bw * 2);
            const uint32_t a_sbo = 8u * p.cw * 2, b_sbo = 8u * p.bw * 2;
            const uint32_t a_kstep = 16u * p.cw * 2, b_kstep = 16u * p.bw * 2;   // 16 pixel rows per MMA
            const uint64_t a_hi = umma_smem_desc(0, (uint32_t)p.a_chunk_bytes, a_sbo, a_layout);
            const uint64_t b_hi = umma_smem_desc(0, (uint32_t)p.b_chunk_bytes, b_sbo, b_layout);
            const uint32_t smem_base = smem_u32(smem);
            uint32_t st = 0, phs = 0, accumulate = 0, probe = 0;   // probe: next stage's barrier, tested before this stage's MMAs
            for (int pb = pb0; pb < pb1; ++pb) {
                if (!probe) mbar_wait(&full[st], phs);
                tc_fence_after();
                {
                    const uint32_t nst = st + 1 == (uint32_t)p.stages ? 0u : st + 1, nph = st + 1 == (uint32_t)p.stages ? phs ^ 1u : phs;
                    probe = mbar_test_wait(&full[nst], nph);
                }
                const uint32_t a_addr = smem_base + st * (uint32_t)p.stage_stride;
                const uint32_t b_lo = (a_addr + (uint32_t)p.a_bytes) >> 4;
                for (int t = 0; t < mt_n; ++t) {
                    const uint32_t a_lo = (a_addr + (uint32_t)t * p.cpt * p.a_chunk_bytes) >> 4;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        if (leader) tc_mma_f16(tmem_base + (uint32_t)t * p.Co_pad, a_hi | (uint64_t)((a_lo + k4 * (a_kstep >> 4)) & 0x3FFFu),
                                   b_hi | (uint64_t)((b_lo + k4 * (b_kstep >> 4)) & 0x3FFFu), idesc, accumulate | (uint32_t)(k4 > 0));
                }
                accumulate = 1;
                if (leader) tc_commit(&empty[st]);
                if (++st == (uint32_t)p.stages) { st = 0; phs ^= 1; }
            }
            if (leader) tc_commit(tfull);
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        if (pb1 > pb0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
            for (int t = 0; t < mt_n; ++t) {
                const int cj = chunk0 + t * p.cpt + row / p.cw;
                const bool valid = cj < p.total_chunks;
                const int tap = cj / p.chunks_per_tap, cc = cj - tap * p.chunks_per_tap;
                const int ci = cc * p.cw + row % p.cw;
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + (uint32_t)t * p.Co_pad;
                float* dst = p.dw + (size_t)tap * p.Ci + ci;
                const size_t co_stride = (size_t)p.taps * p.Ci;
                for (int c0 = 0; c0 < p.Co_pad; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c0, v);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) atomicAdd(dst + (size_t)(c0 + i) * co_stride, __uint_as_float(v[i]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// fv_wgrad_ring.cu: sliding-window schedule for thin full-resolution layers; -1 when not eligible
int conv2d_wgrad_ring_try(const void* x, const void* dy, float* dw_acc, int N, int H, int W, int Ci, int Co_pad, int R, int S, int pad,
                          cudaStream_t stream);

}  // namespace fv

static int wgrad_chunk(const void* x, const void* dy, float* dw_acc, int N, int H, int W, int Ci, int Co_pad, int dy_cs, int R, int S, int pad,
                       void* stream);

// Output-channel counts beyond one UMMA N (256) are handled in chunks of <= 256 channels of dY (channel stride Co_pad).
extern "C" __attribute__((visibility("default"))) int fv_conv2d_wgrad(const void* x, const void* dy, float* dw_acc, int N, int H, int W, int Ci, int Co_pad,
                               int R, int S, int pad, void* stream) {
    using namespace fv;
    if (Co_pad <= 256) return wgrad_chunk(x, dy, dw_acc, N, H, W, Ci, Co_pad, Co_pad, R, S, pad, stream);
    if (Co_pad % 64) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_wgrad: Co_pad=%d > 256 must be a multiple of 64", Co_pad);
    if (!x || !dy || !dw_acc) return fail(FV_ERR_ARG, "fv_conv2d_wgrad: null pointer");
    for (int c0 = 0; c0 < Co_pad; c0 += 256) {
        const int cn = Co_pad - c0 < 256 ? Co_pad - c0 : 256;
        const int e = wgrad_chunk(x, static_cast<const char*>(dy) + (size_t)c0 * 2, dw_acc + (size_t)c0 * R * S * Ci, N, H, W, Ci, cn, Co_pad, R, S,
                                  pad, stream);
        if (e) return e;
    }
    return FV_OK;
}

static int wgrad_chunk(const void* x, const void* dy, float* dw_acc, int N, int H, int W, int Ci, int Co_pad, int dy_cs, int R, int S, int pad,
                       void* stream) {
    using namespace fv;
    if (!x || !dy || !dw_acc) return fail(FV_ERR_ARG, "fv_conv2d_wgrad: null pointer");
    if (Ci % 16 || Ci < 16 || (Ci > 64 && Ci % 64) || (Ci < 64 && Ci != 16 && Ci != 32))
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_wgrad: Ci=%d must be 16, 32 or a multiple of 64", Ci);
    if (Co_pad % 16 || Co_pad < 16 || Co_pad > 256 || (Co_pad > 64 && Co_pad % 64) || (Co_pad < 64 && Co_pad != 16 && Co_pad != 32))
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_wgrad: Co_pad=%d must be 16, 32, 64, 128, 192 or 256", Co_pad);
    if (R != S || (R != 1 && R != 3 && R != 5 && R != 7) || pad != (R - 1) / 2)
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_wgrad: only odd square filters with same padding");
    if (dy_cs == Co_pad) {
        const int rr = conv2d_wgrad_ring_try(x, dy, dw_acc, N, H, W, Ci, Co_pad, R, S, pad, (cudaStream_t)stream);
        if (rr >= 0) return rr;
    }
    WgradParams p{};
    p.N = N; p.H = H; p.W = W; p.Ci = Ci; p.Co_pad = Co_pad; p.R = R; p.S = S; p.pad = pad; p.taps = R * S;
    if (W >= 64) {
        if (W % 64) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_wgrad: W=%d must be a multiple of 64 or a power of two", W);
        p.pw = 64; p.ph = 1; p.pn = 1;
    } else {
        if (W & (W - 1)) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_wgrad: W=%d must be a power of two", W);
        p.pw = W;
        int rest = 64 / W;
        if (H >= rest) {
            if (H % rest) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_wgrad: H=%d not divisible by %d", H, rest);
            p.ph = rest; p.pn = 1;
        } else {
            if (H & (H - 1)) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_wgrad: H=%d must be a power of two", H);
            p.ph = H; p.pn = rest / H;
        }
    }
    p.tiles_w = W / p.pw; p.tiles_h = H / p.ph; p.tiles_n = (N + p.pn - 1) / p.pn;
    p.num_pb = p.tiles_w * p.tiles_h * p.tiles_n;
    p.cw = Ci >= 64 ? 64 : Ci;
    p.cpt = 128 / p.cw;
    p.chunks_per_tap = Ci / p.cw;
    p.total_chunks = p.taps * p.chunks_per_tap;
    p.bw = Co_pad >= 64 ? 64 : Co_pad;
    p.b_chunks = Co_pad / p.bw;
    p.mt_total = (p.total_chunks + p.cpt - 1) / p.cpt;
    int mt_max = 512 / Co_pad;
    if (mt_max > 5) mt_max = 5;                        // keeps a stage <= 16 KB * 5 + B
    p.groups = (p.mt_total + mt_max - 1) / mt_max;
    p.mt_per_group = (p.mt_total + p.groups - 1) / p.groups;
    p.groups = (p.mt_total + p.mt_per_group - 1) / p.mt_per_group;
    int splits = num_sms() / p.groups;
    if (splits < 1) splits = 1;
    if (splits > p.num_pb) splits = p.num_pb;
    p.pb_per_split = (p.num_pb + splits - 1) / splits;
    p.splits = (p.num_pb + p.pb_per_split - 1) / p.pb_per_split;
    p.a_chunk_bytes = 64 * p.cw * 2;
    p.b_chunk_bytes = 64 * p.bw * 2;
    p.a_bytes = p.mt_per_group * p.cpt * p.a_chunk_bytes;
    p.b_bytes = p.b_chunks * p.b_chunk_bytes;
    p.stage_stride = p.a_bytes + ((p.b_bytes + 1023) & ~1023);
    int stages = (200 * 1024) / p.stage_stride;
    if (stages > 6) stages = 6;
    if (stages < 2) return fail(FV_ERR_INTERNAL, "fv_conv2d_wgrad: stage of %d bytes does not fit twice", p.stage_stride);
    p.stages = stages;
    int cols = 32;
    while (cols < p.mt_per_group * Co_pad) cols <<= 1;
    p.tmem_cols = cols;
    p.dw = dw_acc;

    CUtensorMap tmX, tmDY;
    {
        uint64_t dims[4] = {(uint64_t)Ci, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)Ci * 2, (uint64_t)W * Ci * 2, (uint64_t)H * W * Ci * 2};
        uint32_t box[4] = {(uint32_t)p.cw, (uint32_t)p.pw, (uint32_t)p.ph, (uint32_t)p.pn};
        if (int e = encode_tmap_bf16(&tmX, x, 4, dims, str, box, p.cw * 2)) return e;
    }
    {
        uint64_t dims[4] = {(uint64_t)Co_pad, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)dy_cs * 2, (uint64_t)W * dy_cs * 2, (uint64_t)H * W * dy_cs * 2};
        uint32_t box[4] = {(uint32_t)p.bw, (uint32_t)p.pw, (uint32_t)p.ph, (uint32_t)p.pn};
        if (int e = encode_tmap_bf16(&tmDY, dy, 4, dims, str, box, p.bw * 2)) return e;
    }
    const size_t smem = (size_t)p.stages * p.stage_stride + 1024 + 512;
    static bool attr_set = false;
    if (!attr_set) {
        FV_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    conv_wgrad_kernel<<<p.groups * p.splits, kWgradThreads, smem, (cudaStream_t)stream>>>(tmX, tmDY, p);
    FV_LAUNCH_CHECK("conv_wgrad_kernel");
    return FV_OK;
}
