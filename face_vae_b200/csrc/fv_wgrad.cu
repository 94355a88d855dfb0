// Weight gradient of the convolutions of fv_conv.cu as a tcgen05 GEMM with the pixels as the K dimension:
//
//   SAME : dW[co, tap, ci] = sum_pixels X[pixel + tap_offset, ci] * dY[pixel, co]
//   X2   : dWp[phase][co, tap, ci] = sum_{coarse pixels} X[(i, j) + tap_offset(phase, tap), ci] * dY[(2i + a, 2j + b), co]
//          (the four 2x2 phase filters of the up-sampling convolution; fv_wgrad_finish_up folds them back into the 3x3 filter)
//   S2   : dW[co, (r4, s4), ci] = sum_{coarse pixels} X[(2i + r4 - 1, 2j + s4 - 1), ci] * dY[(i, j), co]
//
// Both operands are NHWC, i.e. the contraction index (pixel) is the slow index in memory, so they are fed to the
// tensor core as MN-major operands straight from TMA boxes (no transpose pass):
//   A (M side) = shifted input windows.  One 128-row M tile stacks 128/cw chunks of cw = min(Ci,64) channels, each
//                chunk being one (filter tap, channel chunk) pair, so thin layers (Ci = 16/32/64) still fill M = 128.
//   B (N side) = dY, N = Co_pad.
// The tensor on the FINE grid (dY for X2, X for S2) is read through the 5-D view (2C, W, 2, H, N) of fv_conv.cu: a box at row
// parity a and channel offset b*C is the stride-2 sub-lattice (2i + a, 2j + b).
//   D[(tap,ci), co] accumulates in TMEM over this CTA's share of the pixels (split-K across CTAs); every split STORES its
//   partial into its own slab of the workspace [splits][...]; fv_wgrad_finish adds the slabs in split order (reproducible --
//   round 1 used red.global.add.f32 into one buffer, whose arrival order changed the last bits from run to run).
// Replaces the wgrad half of aten::convolution_backward behind nn.Conv2d (reference modules.py:15,32).
#include "../../include/facevae_b200.h"
#include "fv_host.h"
#include "fv_ptx.cuh"

namespace fv {

enum : int { CONV_SAME = 0, CONV_X2 = 1, CONV_S2 = 2 };

struct WgradParams {
    int N, H, W, Ci, Co_pad, taps, nph;
    int pw, ph, pn, tiles_w, tiles_h, tiles_n, num_pb;   // 64-pixel K blocks
    int cw, cpt, chunks_per_tap, total_chunks;           // A chunking
    int bw, b_chunks;                                    // B chunking
    int mt_total, mt_per_group, groups, splits, pb_per_split;
    int a_chunk_bytes, b_chunk_bytes, a_bytes, b_bytes, stage_stride, stages, tmem_cols;
    float* dw;                                           // partial slabs
    long long split_stride, phase_stride, co_stride;     // elements; D[(tap, ci), co] -> dw[split][phase][co*co_stride + tap*Ci_total + ci]
    int ci_total;
    short4 tapA[64];                                     // per (phase, tap): x = channel offset, y = dw, z = row parity, w = dh of the A box
    short4 tapB[4];                                      // per phase: x = channel offset, z = row parity of the dY box
};

static constexpr int kWgradThreads = 192;

__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // align up with arithmetic on the array itself so the compiler keeps the shared address space (LDS/STS, not generic)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_stride);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_phase = p.groups * p.splits;
    const int phase = blockIdx.x / per_phase, brem = blockIdx.x - phase * per_phase;
    const int group = brem % p.groups, split = brem / p.groups;
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
            const short4 tb = p.tapB[phase];
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
                    if (leader) tma_load_5d(b_dst + (size_t)bc * p.b_chunk_bytes, &tmDY, &full[st], tb.x + bc * p.bw, w0, tb.z, h0, n0);
                int tap = chunk0 / p.chunks_per_tap, cc = chunk0 - tap * p.chunks_per_tap;
                for (int j = 0; j < n_chunks; ++j) {
                    const short4 ta = p.tapA[phase * p.taps + tap];
                    if (leader)
                        tma_load_5d(a_dst + (size_t)j * p.a_chunk_bytes, &tmX, &full[st], ta.x + cc * p.cw, w0 + ta.y, ta.z, h0 + ta.w, n0);
                    if (++cc == p.chunks_per_tap) { cc = 0; ++tap; }
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
            float* slab = p.dw + (size_t)split * p.split_stride + (size_t)phase * p.phase_stride;
            for (int t = 0; t < mt_n; ++t) {
                const int cj = chunk0 + t * p.cpt + row / p.cw;
                const bool valid = cj < p.total_chunks;
                const int tap = cj / p.chunks_per_tap, cc = cj - tap * p.chunks_per_tap;
                const int ci = cc * p.cw + row % p.cw;
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + (uint32_t)t * p.Co_pad;
                float* dst = slab + (size_t)tap * p.ci_total + ci;
                for (int c0 = 0; c0 < p.Co_pad; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c0, v);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) dst[(size_t)(c0 + i) * p.co_stride] = __uint_as_float(v[i]);   // lanes = consecutive ci: coalesced
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

// fv_wgrad_ring.cu: sliding-window schedule for thin full-resolution layers
int conv2d_wgrad_ring_splits(int N, int H, int W, int Ci, int Co_pad, int R, int S);     // 0 when not eligible
int conv2d_wgrad_ring_try(const void* x, const void* dy, float* part, long long split_stride, int N, int H, int W, int Ci, int Co_pad, int R,
                          int S, int pad, cudaStream_t stream);
int conv2d_wgrad_ring_x2_splits(int N, int H, int W, int Ci, int Co_pad);                 // 0 when not eligible
int conv2d_wgrad_ring_x2_try(const void* x, const void* dy, float* part, long long split_stride, int N, int H, int W, int Ci, int Co_pad,
                             cudaStream_t stream);

static int pick_pixel_block(int H, int W, int& pw, int& ph, int& pn) {
    if (W >= 64) {
        if (W % 64) return 1;
        pw = 64; ph = 1; pn = 1;
        return 0;
    }
    if (W < 1 || (W & (W - 1))) return 1;
    pw = W;
    const int rest = 64 / W;
    if (H >= rest) {
        if (H % rest) return 1;
        ph = rest; pn = 1;
    } else {
        if (H & (H - 1)) return 1;
        ph = H; pn = rest / H;
    }
    return 0;
}

// geometry of the generic kernel for one chunk of <= 256 output channels: M-tile groups and pixel splits
static int plan_generic(int kind, int N, int H, int W, int Ci, int Co_pad, int taps, WgradParams& p) {
    p.N = N; p.H = H; p.W = W; p.Ci = Ci; p.Co_pad = Co_pad; p.taps = taps;
    p.nph = kind == CONV_X2 ? 4 : 1;
    if (pick_pixel_block(H, W, p.pw, p.ph, p.pn)) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_wgrad: H=%d W=%d not tileable (W multiple of 64 or powers of two)", H, W);
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
    int splits = num_sms() / (p.groups * p.nph);
    if (splits < 1) splits = 1;
    if (splits > p.num_pb) splits = p.num_pb;
    p.pb_per_split = (p.num_pb + splits - 1) / splits;
    p.splits = (p.num_pb + p.pb_per_split - 1) / p.pb_per_split;
    return FV_OK;
}

static int check_channels(int Ci, int Co_pad) {
    if (Ci % 16 || Ci < 16 || (Ci > 64 && Ci % 64) || (Ci < 64 && Ci != 16 && Ci != 32))
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_wgrad: Ci=%d must be 16, 32 or a multiple of 64", Ci);
    if (Co_pad % 16 || Co_pad < 16 || (Co_pad > 64 && Co_pad % 64) || (Co_pad < 64 && Co_pad != 16 && Co_pad != 32))
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_wgrad: Co_pad=%d must be 16, 32 or a multiple of 64", Co_pad);
    return FV_OK;
}

// number of partial slabs the launch for this shape writes (the caller allocates splits * slab elements and hands the count to
// fv_wgrad_finish); kind: 0 same, 1 x2, 2 s2.  0 on unsupported shapes (the launch itself reports the reason).
static int wgrad_splits(int kind, int N, int H, int W, int Ci, int Co_pad, int R, int S) {
    if (check_channels(Ci, Co_pad)) return 0;
    if (kind == CONV_SAME) {
        const int rs = conv2d_wgrad_ring_splits(N, H, W, Ci, Co_pad, R, S);
        if (rs > 0) return rs;
    }
    if (kind == CONV_X2) {
        const int rs = conv2d_wgrad_ring_x2_splits(N, H, W, Ci, Co_pad);
        if (rs > 0) return rs;
    }
    int splits = 0;
    for (int c0 = 0; c0 < Co_pad; c0 += 256) {      // all chunks of a layer share the pixel split (same grid geometry)
        WgradParams p{};
        const int cn = Co_pad - c0 < 256 ? Co_pad - c0 : 256;
        if (plan_generic(kind, N, H, W, Ci, cn, kind == CONV_X2 ? 4 : (kind == CONV_S2 ? 16 : R * S), p)) return 0;
        if (p.splits > splits) splits = p.splits;
    }
    return splits;
}

static int wgrad_chunk(int kind, const void* x, const void* dy, float* part, long long split_stride, int N, int H, int W, int Ci, int Co_pad,
                       int dy_cs, int dy_c0, int R, int S, int pad, int splits_expected, void* stream) {
    WgradParams p{};
    const int taps = kind == CONV_X2 ? 4 : (kind == CONV_S2 ? 16 : R * S);
    if (int e = plan_generic(kind, N, H, W, Ci, Co_pad, taps, p)) return e;
    if (p.splits != splits_expected) return fail(FV_ERR_INTERNAL, "fv_conv2d_wgrad: the channel chunks of this layer disagree on the pixel split");
    if (p.nph * taps > 64) return fail(FV_ERR_INTERNAL, "fv_conv2d_wgrad: tap table overflow");
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
    p.dw = part + (size_t)dy_c0 * taps * Ci;            // this chunk's output channels inside every slab / phase block
    p.split_stride = split_stride;
    p.phase_stride = (long long)dy_cs * taps * Ci;
    p.co_stride = (long long)taps * Ci;
    p.ci_total = Ci;
    for (int ph = 0; ph < p.nph; ++ph) {
        const int a = ph >> 1, b = ph & 1;
        p.tapB[ph] = make_short4((short)(kind == CONV_X2 ? b * dy_cs + dy_c0 : dy_c0), 0, (short)(kind == CONV_X2 ? a : 0), 0);
        for (int t = 0; t < taps; ++t) {
            short4 ta = make_short4(0, 0, 0, 0);
            if (kind == CONV_SAME) {
                ta.y = (short)(t % S - pad); ta.w = (short)(t / S - pad);
            } else if (kind == CONV_X2) {
                const int u = t >> 1, v = t & 1;
                ta.y = (short)(v - 1 + b); ta.w = (short)(u - 1 + a);
            } else {
                const int fr = (t >> 2) - 1, fc = (t & 3) - 1;
                const int dh = fr < 0 ? -1 : fr / 2, dwc = fc < 0 ? -1 : fc / 2;
                ta.w = (short)dh; ta.z = (short)(fr - 2 * dh);
                ta.y = (short)dwc; ta.x = (short)((fc - 2 * dwc) * Ci);
            }
            p.tapA[ph * taps + t] = ta;
        }
    }

    CUtensorMap tmX, tmDY;
    {
        const uint64_t fx = kind == CONV_S2 ? 2 : 1;       // X on the fine grid (S2): the 5-D parity view
        uint64_t dims[5] = {(uint64_t)Ci * fx, (uint64_t)W, fx, (uint64_t)H, (uint64_t)N};
        uint64_t str[4] = {(uint64_t)Ci * fx * 2, (uint64_t)W * fx * Ci * 2, (uint64_t)W * fx * Ci * 2 * fx, (uint64_t)H * fx * W * fx * Ci * 2};
        uint32_t box[5] = {(uint32_t)p.cw, (uint32_t)p.pw, 1, (uint32_t)p.ph, (uint32_t)p.pn};
        if (int e = encode_tmap_bf16(&tmX, x, 5, dims, str, box, p.cw * 2)) return e;
    }
    {
        const uint64_t fy = kind == CONV_X2 ? 2 : 1;       // dY on the fine grid (X2)
        uint64_t dims[5] = {(uint64_t)dy_cs * fy, (uint64_t)W, fy, (uint64_t)H, (uint64_t)N};
        uint64_t str[4] = {(uint64_t)dy_cs * fy * 2, (uint64_t)W * fy * dy_cs * 2, (uint64_t)W * fy * dy_cs * 2 * fy, (uint64_t)H * fy * W * fy * dy_cs * 2};
        uint32_t box[5] = {(uint32_t)p.bw, (uint32_t)p.pw, 1, (uint32_t)p.ph, (uint32_t)p.pn};
        if (int e = encode_tmap_bf16(&tmDY, dy, 5, dims, str, box, p.bw * 2)) return e;
    }
    const size_t smem = (size_t)p.stages * p.stage_stride + 1024 + 512;
    static bool attr_set = false;
    if (!attr_set) {
        FV_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    conv_wgrad_kernel<<<p.nph * p.groups * p.splits, kWgradThreads, smem, (cudaStream_t)stream>>>(tmX, tmDY, p);
    FV_LAUNCH_CHECK("conv_wgrad_kernel");
    return FV_OK;
}

static int wgrad_any(int kind, const void* x, const void* dy, float* part, int splits, int N, int H, int W, int Ci, int Co_pad, int R, int S, int pad,
                     void* stream) {
    if (!x || !dy || !part) return fail(FV_ERR_ARG, "fv_conv2d_wgrad: null pointer");
    if (int e = check_channels(Ci, Co_pad)) return e;
    if (kind == CONV_SAME && (R != S || (R != 1 && R != 3 && R != 5 && R != 7) || pad != (R - 1) / 2))
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_wgrad: only odd square filters with same padding");
    const int taps = kind == CONV_X2 ? 4 : (kind == CONV_S2 ? 16 : R * S);
    const int nph = kind == CONV_X2 ? 4 : 1;
    if (splits != wgrad_splits(kind, N, H, W, Ci, Co_pad, R, S))
        return fail(FV_ERR_ARG, "fv_conv2d_wgrad: splits=%d does not match fv_conv2d_wgrad_splits() for this shape", splits);
    const long long split_stride = (long long)nph * Co_pad * taps * Ci;
    if (kind == CONV_SAME) {
        const int rr = conv2d_wgrad_ring_try(x, dy, part, split_stride, N, H, W, Ci, Co_pad, R, S, pad, (cudaStream_t)stream);
        if (rr >= 0) return rr;
    }
    if (kind == CONV_X2) {
        const int rr = conv2d_wgrad_ring_x2_try(x, dy, part, split_stride, N, H, W, Ci, Co_pad, (cudaStream_t)stream);
        if (rr >= 0) return rr;
    }
    // output-channel counts beyond one UMMA N (256) are handled in chunks of <= 256 channels of dY
    for (int c0 = 0; c0 < Co_pad; c0 += 256) {
        const int cn = Co_pad - c0 < 256 ? Co_pad - c0 : 256;
        if (int e = wgrad_chunk(kind, x, dy, part, split_stride, N, H, W, Ci, cn, Co_pad, c0, R, S, pad, splits, stream)) return e;
    }
    return FV_OK;
}

}  // namespace fv

using namespace fv;

// kind: 0 = stride-1 "same" (R x S), 1 = x2 (up-sampling conv, four 2x2 phases), 2 = s2 (4x4 stride 2).  H, W: the tiling grid
// (x2 / s2: coarse resolution).
extern "C" __attribute__((visibility("default"))) int fv_conv2d_wgrad_splits(int kind, int N, int H, int W, int Ci, int Co_pad, int R, int S) {
    return wgrad_splits(kind, N, H, W, Ci, Co_pad, R, S);
}

extern "C" __attribute__((visibility("default"))) int fv_conv2d_wgrad(const void* x, const void* dy, float* part, int splits, int N, int H, int W, int Ci,
                                                                    int Co_pad, int R, int S, int pad, void* stream) {
    return wgrad_any(CONV_SAME, x, dy, part, splits, N, H, W, Ci, Co_pad, R, S, pad, stream);
}

// x [N,H,W,Ci] coarse, dy [N,2H,2W,Co_pad] fine -> part [splits][4 phases][Co_pad][4 taps][Ci]
extern "C" __attribute__((visibility("default"))) int fv_conv2d_wgrad_x2(const void* x, const void* dy, float* part, int splits, int N, int H, int W, int Ci,
                                                                       int Co_pad, void* stream) {
    return wgrad_any(CONV_X2, x, dy, part, splits, N, H, W, Ci, Co_pad, 2, 2, 0, stream);
}

// x [N,2H,2W,Ci] fine, dy [N,H,W,Co_pad] coarse -> part [splits][Co_pad][16 taps][Ci]
extern "C" __attribute__((visibility("default"))) int fv_conv2d_wgrad_s2(const void* x, const void* dy, float* part, int splits, int N, int H, int W, int Ci,
                                                                       int Co_pad, void* stream) {
    return wgrad_any(CONV_S2, x, dy, part, splits, N, H, W, Ci, Co_pad, 4, 4, 1, stream);
}
