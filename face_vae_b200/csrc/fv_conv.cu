// Implicit-GEMM 2-D convolution for sm_100a: NHWC bf16 in, fp32 accumulate in TMEM.  Three geometries share one kernel:
//
//   SAME : Y[p, co] = sum_{r,s,ci} X[p + (r - pad, s - pad), ci] * Wt[co, (r*S + s)*Ci + ci]     stride 1, odd filter, "same" padding
//          -- the nn.Conv2d inside the reference's _ConvBlock (reference modules.py:15,32), forward, and (called with the
//          rotated / transposed filter of fv_weight_prep) its data gradient;
//   X2   : nearest 2x up-sampling followed by a 3x3 "same" convolution (UpBlock2D, reference modules.py:78-89) WITHOUT the
//          up-sampled tensor: output pixel (2i + a, 2j + b) only ever sees a 2x2 neighbourhood of the coarse input, so the
//          layer is four 2x2 convolutions ("phases" (a, b)) on the coarse grid whose filters are sums of the 3x3 taps
//          (fv_weight_prep_up) -- 16 tap-GEMMs per coarse pixel instead of 36 and a 4x smaller input to read.  The same
//          schedule is the data gradient of a 4x4 stride-2 convolution;
//   S2   : 4x4 stride-2 pad-1 convolution (Conv2dELR as used by EFE_conv6, reference models_utils.py:632-744,
//          models.py:845-852) and -- with summed, mirrored taps -- the data gradient of X2 (it folds the 2x2 sum of the
//          up-sampling backward into the GEMM).  The fine NHWC tensor [N,2H,2W,C] is addressed through a 5-D tensor map
//          (2C, W, 2, H, N): tap (r, s) is a box at row parity a, channel offset b*C, coarse offset (dh, dw).
//
// GEMM view: M = pixels of the tiling grid (128 per tile), N = Co_pad (one UMMA N, <= 256), K = taps * Ci.
//
// Structure (one persistent CTA per SM, 6 warps):
//   warp 0  : TMA producer.  Activations come through a 5-D tensor map; pixels outside the image are zero-filled by the
//             TMA unit, so padding costs no memory and no branches.  A "group" is one activation box plus the filter
//             slices it feeds: one filter tap (tap schedule: box (KB ch, tw, th, tn)), or -- SAME geometry with row-segment
//             tiles -- one filter ROW (slab schedule: box (KB ch, tw + S - 1); the S taps of that row are the same
//             shared-memory slab read through UMMA descriptors whose start address is shifted by s pixel rows).  Box
//             coordinates come from a per-(phase, group) table in the kernel parameters.
//   warp 1  : one thread issues tcgen05.mma (UMMA 128 x Nc x 16, bf16 -> fp32) into one of two TMEM accumulator
//             buffers and commits to the stage / accumulator mbarriers.
//   warps 2-5: epilogue.  tcgen05.ld their TMEM lane quarter (row = pixel), add bias / residual, convert and store
//             (NHWC bf16, NHWC fp32 or NCHW fp32); overlapped with the next tile's MMAs via the second buffer.  Optional
//             fused batch-norm statistics of the stored values, reduced across CTAs in a fixed order (fv_reduce.cuh).
#include <cstdio>
#include <type_traits>
#include <cstdlib>

#include "../../include/facevae_b200.h"
#include "fv_host.h"
#include "fv_ptx.cuh"
#include "fv_reduce.cuh"

namespace fv {

enum : int { CONV_SAME = 0, CONV_X2 = 1, CONV_S2 = 2 };

struct ConvParams {
    int N, H, W;            // tiling grid (X2 / S2: the COARSE resolution)
    int Ci, Co, Co_pad;
    int Ho, Wo, osy, osx;   // output image size and pixel stride: tile pixel (h, w) of phase ph -> (h*osy + oy[ph], w*osx + ox[ph])
    int tw, th, tn, tiles_w, tiles_h, tiles_n, num_tiles;
    int kc_blocks, stages;
    int kcps;               // K blocks per shared-memory stage (1, 2 or 4): fewer barrier round trips per MMA for the small stages
    int groups;             // TMA groups per (phase, tile)
    int nsub;               // filter taps per group (slab schedule: S, tap schedule: 1)
    int nph;                // output phases: 1, or 4 (X2)
    int a_off_b;            // byte offset of the filter slices inside a stage (activation region, rounded to 1 KB)
    int b_slice_stride;     // placement stride of one [Nc x KB] filter slice (rounded to 1 KB)
    int stage_stride;
    int tx_bytes;           // bytes the TMA unit delivers per stage (what the full barrier is armed with)
    int out_mode, tmem_cols;
    int out_cs;             // channel stride (elements) of the NHWC output / residual rows
    int co_base;            // first output channel of this launch inside the out_cs-wide rows (Co_pad > 256 is walked in chunks)
    int w_rows_per_phase;   // rows of the filter matrix per phase (= total Co_pad of the layer)
    int co_parts, Nc;       // output channels split over co_parts CTAs per pixel tile (few tiles): Nc = Co_pad / co_parts
    int num_vtiles;         // nph * co_parts * num_tiles
    // filter-resident mode (b_res): a CTA works on a contiguous run of pixel tiles of ONE (phase, channel part), loads that filter
    // slice once (b_res_off) and streams only activation stages -- otherwise every 128-pixel tile re-fetches the whole slice
    int b_res, b_res_off, b_slice_bytes, bar_off;
    int ctas_per_slice, tiles_per_cta;
    const float* bias;
    const __nv_bfloat16* residual;
    void* out;
    float* stats;           // optional [2][stats_c]: per-channel sum / sum of squares of the stored output (fused fv_bn_stats)
    int stats_c;
    void* red_ws;           // fv_reduce.cuh workspace (with stats)
    int act;                // FV_ACT_*: applied after bias (+ residual), before the store (Conv2dELR: conv -> bias -> LeakyReLU)
    short4 tap[64];         // per (phase, group): x = channel offset, y = dw, z = row parity / 0, w = dh of the activation box
    signed char oy[4], ox[4];
};

static constexpr int kConvThreads = 192;

struct TileCoord {
    int w0, h0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile) {
    const int per_n = p.tiles_w * p.tiles_h;
    const int tn_i = tile / per_n;
    const int rem = tile - tn_i * per_n;
    const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
    return {tw_i * p.tw, th_i * p.th, tn_i * p.tn};
}
// virtual tile -> (phase, output-channel part, pixel tile); consecutive CTAs work on the same filter slice
__device__ __forceinline__ void decode_vtile(const ConvParams& p, int vt, int& ph, int& part, int& tile) {
    const int per_ph = p.num_tiles * p.co_parts;
    ph = vt / per_ph;
    const int rem = vt - ph * per_ph;
    part = rem / p.num_tiles;
    tile = rem - part * p.num_tiles;
}

template <int KB>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ ConvParams p) {
    constexpr int ROW = KB * 2;                       // bytes per pixel row of a K block == swizzle span
    constexpr int KSUB = KB / 16;                     // UMMA K steps per K block
    constexpr uint32_t LAYOUT = ROW == 128 ? 2u : (ROW == 64 ? 4u : 6u);
    constexpr uint32_t SBO = 8u * ROW;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // align up with arithmetic on the array itself so the compiler keeps the shared address space (LDS/STS, not generic)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.bar_off);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint64_t* tempty = tfull + 2;
    uint64_t* wfull = tempty + 2;                                 // resident filter slice landed (b_res)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
    int* red_flag = reinterpret_cast<int*>(tmem_slot + 2);
    float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);      // [Co_pad]
    float* stat_w = bias_s + p.Co_pad;                            // [4 epilogue warps][2][Co_pad] (when p.stats)
    float* stat_blk = stat_w + 8 * p.Co_pad;                      // [2][Co_pad] this CTA's totals
    float* stat_tot = stat_blk + 2 * p.Co_pad;                    // [2][Co_pad] grid totals (in the last CTA)

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
        mbar_init(wfull, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
        tmem_relinquish();
    }
    // this CTA's virtual tiles: strided over the grid, or (b_res) a contiguous run inside one (phase, part) slice
    int vt_begin = blockIdx.x, vt_end = p.num_vtiles, vt_step = gridDim.x;
    if (p.b_res) {
        const int slice = blockIdx.x / p.ctas_per_slice, j = blockIdx.x - slice * p.ctas_per_slice;
        vt_begin = slice * p.num_tiles + j * p.tiles_per_cta;
        vt_end = min(vt_begin + p.tiles_per_cta, (slice + 1) * p.num_tiles);
        vt_step = 1;
    }
    for (int c = threadIdx.x; c < p.Co_pad; c += blockDim.x) bias_s[c] = (p.bias && c < p.Co) ? p.bias[c] : 0.f;
    if (p.stats)
        for (int c = threadIdx.x; c < 8 * p.Co_pad; c += blockDim.x) stat_w[c] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {
            const bool leader = elect_one_sync();      // role loops stay warp-uniform; only the issue is predicated
            uint32_t st = 0, ph = 0;
            if (p.b_res && vt_begin < vt_end) {       // the whole filter slice of this CTA's (phase, part), once
                int phase, part, tile;
                decode_vtile(p, vt_begin, phase, part, tile);
                const int wrow = phase * p.w_rows_per_phase + p.co_base + part * p.Nc;
                if (leader) mbar_arrive_expect_tx(wfull, (uint32_t)(p.groups * p.kc_blocks * p.nsub * p.b_slice_bytes));
                for (int g = 0; g < p.groups; ++g)
                    for (int kc = 0; kc < p.kc_blocks; ++kc)
                        for (int sm = 0; sm < p.nsub; ++sm)
                            if (leader)
                                tma_load_2d(smem + p.b_res_off + (size_t)((g * p.kc_blocks + kc) * p.nsub + sm) * p.b_slice_stride, &tmW, wfull,
                                            (g * p.nsub + sm) * p.Ci + kc * KB, wrow);
            }
            for (int vt = vt_begin; vt < vt_end; vt += vt_step) {
                int phase, part, tile;
                decode_vtile(p, vt, phase, part, tile);
                const int wrow = phase * p.w_rows_per_phase + p.co_base + part * p.Nc;   // first row of this CTA's filter slice
                const TileCoord t = decode_tile(p, tile);
                for (int g = 0; g < p.groups; ++g) {
                    const short4 tp = p.tap[phase * p.groups + g];
                    const int wtap = g * p.nsub * p.Ci;                   // first filter column of this group
                    for (int kc0 = 0; kc0 < p.kc_blocks; kc0 += p.kcps) {
                        mbar_wait(&empty[st], ph ^ 1);
                        uint8_t* a_dst = smem + (size_t)st * p.stage_stride;
                        uint8_t* b_dst = a_dst + p.kcps * p.a_off_b;
                        if (leader) mbar_arrive_expect_tx(&full[st], (uint32_t)(p.tx_bytes * p.kcps));
                        for (int q = 0; q < p.kcps; ++q) {
                            const int kc = kc0 + q;
                            if (leader) tma_load_5d(a_dst + q * p.a_off_b, &tmX, &full[st], kc * KB + tp.x, t.w0 + tp.y, tp.z, t.h0 + tp.w, t.n0);
                            if (!p.b_res)
                                for (int sm = 0; sm < p.nsub; ++sm)
                                    if (leader)
                                        tma_load_2d(b_dst + (q * p.nsub + sm) * p.b_slice_stride, &tmW, &full[st], wtap + sm * p.Ci + kc * KB, wrow);
                        }
                        if (++st == (uint32_t)p.stages) { st = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        {
            const bool leader = elect_one_sync();
            const uint32_t idesc = umma_idesc_bf16(128, p.Nc, 0, 0);
            const uint64_t desc_hi = umma_smem_desc(0, 16, SBO, LAYOUT);   // template: all fields but the start address
            const uint32_t smem_base = smem_u32(smem);
            const int groups = p.groups * (p.kc_blocks / p.kcps);     // stages per tile
            uint32_t st = 0, ph = 0, tcount = 0;
            // `probe`: the NEXT stage's full barrier, tested (non-blocking) before the current stage's MMAs are issued and
            // consumed after them -- an mbarrier round trip costs the issuing warp 150-300 cycles even when the phase is
            // complete, and the (blocking) issue of 4-12 MMAs per stage hides it
            uint32_t probe = 0;
            const uint32_t a_q16 = (uint32_t)p.a_off_b >> 4, b_slice16 = (uint32_t)p.b_slice_stride >> 4;
            const uint32_t b_stage_off = (uint32_t)(p.kcps * p.a_off_b) >> 4, bres_base = (smem_base + (uint32_t)p.b_res_off) >> 4;
            if (p.b_res && vt_begin < vt_end) mbar_wait(wfull, 0);
            // The whole tile loop is instantiated for the common (K blocks per stage, taps per group) pairs and selected ONCE: a
            // run-time loop nest -- or a per-stage dispatch -- around four MMAs costs the issuing thread ~5-9 % of the kernel (measured)
            auto run = [&](auto q_tag, auto s_tag) {
                constexpr int Q = decltype(q_tag)::value, NS = decltype(s_tag)::value;      // 0 = run-time count
                const int nq = Q ? Q : p.kcps, ns = NS ? NS : p.nsub;
                for (int vt = vt_begin; vt < vt_end; vt += vt_step, ++tcount) {
                    const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
                    mbar_wait(&tempty[acc], aph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.Nc;
                    uint32_t accumulate = 0;
                    uint32_t b_res_lo = bres_base;
                    for (int g = 0; g < groups; ++g) {
                        if (!probe) mbar_wait(&full[st], ph);
                        tc_fence_after();
                        {
                            const uint32_t nst = st + 1 == (uint32_t)p.stages ? 0u : st + 1, nph = st + 1 == (uint32_t)p.stages ? ph ^ 1u : ph;
                            probe = mbar_test_wait(&full[nst], nph);
                        }
                        // descriptor start addresses (16-byte units) advance by additions only
                        uint32_t a_lo = (smem_base + st * (uint32_t)p.stage_stride) >> 4;
                        uint32_t b_lo = p.b_res ? b_res_lo : a_lo + b_stage_off;
#pragma unroll
                        for (int q = 0; q < nq; ++q) {
                            uint32_t a_s = a_lo;
#pragma unroll
                            for (int sm = 0; sm < ns; ++sm) {               // slab: tap sm == slab shifted by sm pixel rows
#pragma unroll
                                for (int j = 0; j < KSUB; ++j) {
                                    if (leader)
                                        tc_mma_f16(d_tmem, desc_hi | (uint64_t)((a_s + 2 * j) & 0x3FFFu), desc_hi | (uint64_t)((b_lo + 2 * j) & 0x3FFFu),
                                                   idesc, accumulate);
                                    accumulate = 1;
                                }
                                a_s += ROW >> 4;
                                b_lo += b_slice16;
                            }
                            a_lo += a_q16;
                        }
                        b_res_lo = b_lo;
                        if (leader) tc_commit(&empty[st]);   // frees the smem stage once these MMAs have read it
                        if (++st == (uint32_t)p.stages) { st = 0; ph ^= 1; }
                    }
                    if (leader) tc_commit(&tfull[acc]);      // accumulator complete -> epilogue
                }
            };
            using I0 = std::integral_constant<int, 0>;
            using I1 = std::integral_constant<int, 1>;
            using I2 = std::integral_constant<int, 2>;
            using I3 = std::integral_constant<int, 3>;
            using I4 = std::integral_constant<int, 4>;
            if (p.nsub == 1) {
                if (p.kcps == 1) run(I1{}, I1{});
                else if (p.kcps == 2) run(I2{}, I1{});
                else run(I4{}, I1{});
            } else if (p.nsub == 3) {
                if (p.kcps == 1) run(I1{}, I3{});
                else if (p.kcps == 2) run(I2{}, I3{});
                else run(I0{}, I0{});
            } else {
                run(I0{}, I0{});
            }
        }
    } else {
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;       // tile row == pixel within the tile
        const int w_l = row % p.tw, h_l = (row / p.tw) % p.th, n_l = row / (p.tw * p.th);
        float* my_stat = stat_w + q * 2 * p.Co_pad;      // this warp's accumulators: lane pair k <-> channel 16 c + k, plain adds
        uint32_t tcount = 0;
        for (int vt = vt_begin; vt < vt_end; vt += vt_step, ++tcount) {
            int phase, part, tile;
            decode_vtile(p, vt, phase, part, tile);
            const int co0 = part * p.Nc;
            const TileCoord t = decode_tile(p, tile);
            const int w = t.w0 + w_l, h = t.h0 + h_l, n = t.n0 + n_l;
            const int oh = h * p.osy + p.oy[phase], ow = w * p.osx + p.ox[phase];
            const bool valid = n < p.N;
            const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * (uint32_t)p.Nc;
            const size_t pix = ((size_t)n * p.Ho + oh) * p.Wo + ow;
            for (int cl = 0; cl < p.Nc; cl += 16) {
                const int c0 = co0 + cl;                       // output channel of column cl within this launch's chunk
                uint32_t v[16];
                tmem_ld16(taddr + cl, v);
                tmem_ld_wait();
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = 0.f;          // rows past the batch contribute nothing to the statistics
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + bias_s[c0 + i];
                    if (p.act != FV_ACT_NONE && !p.residual) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) f[i] = f[i] > 0.f ? f[i] : (p.act == FV_ACT_LEAKY ? 0.2f * f[i] : 0.f);
                    }
                    if (p.out_mode == FV_OUT_NCHW_F32) {
                        float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < p.Co) o[(((size_t)n * p.Co + c0 + i) * p.Ho + oh) * p.Wo + ow] = f[i];
                    } else {
                        const size_t eoff = pix * p.out_cs + p.co_base + c0;
                        if (p.residual) {
                            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + eoff);
                            const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                f[2 * i] += bf16_lo(rr[i]);
                                f[2 * i + 1] += bf16_hi(rr[i]);
                            }
                        }
                        // 256-bit stores: a thread owns 32 (bf16) / 64 (fp32) contiguous bytes of its pixel row; 16-byte
                        // stores at a row-sized lane stride reach L2 as half-written sectors (2x write traffic in ncu)
                        if (p.out_mode == FV_OUT_NHWC_BF16) {
                            uint32_t wv[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) wv[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
                            st_global_256(reinterpret_cast<__nv_bfloat16*>(p.out) + eoff, wv);
                            if (p.stats) {                       // statistics of the values as stored (bf16-rounded)
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    f[2 * i] = bf16_lo(wv[i]);
                                    f[2 * i + 1] = bf16_hi(wv[i]);
                                }
                            }
                        } else {
                            uint32_t wv[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) wv[i] = __float_as_uint(f[i]);
                            float* o = reinterpret_cast<float*>(p.out) + eoff;
                            st_global_256(o, wv);
                            st_global_256(o + 8, wv + 8);
                        }
                    }
                }
                if (p.stats) {                                    // warp-uniform: all 32 lanes take part in the shuffles
                    float fq[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) fq[i] = f[i] * f[i];
                    const float cs = warp_colsum16(f, lane), cq = warp_colsum16(fq, lane);
                    if (!(lane & 1)) {                            // one owner lane per (warp, channel): no atomics, fixed order
                        my_stat[c0 + ((lane >> 1) & 15)] += cs;
                        my_stat[p.Co_pad + c0 + ((lane >> 1) & 15)] += cq;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        if (p.stats) {                                   // warps in order -> CTA totals -> CTAs in order (fv_reduce.cuh)
            const int tid = threadIdx.x - 64, n2 = 2 * p.Co_pad;
            named_bar_sync(1, 128);
            for (int c = tid; c < n2; c += 128) stat_blk[c] = ((stat_w[c] + stat_w[n2 + c]) + stat_w[2 * n2 + c]) + stat_w[3 * n2 + c];
            named_bar_sync(1, 128);
            if (det_reduce<float>(p.red_ws, n2, gridDim.x, blockIdx.x, stat_blk, stat_tot, tid, 128, NamedSync{1, 128}, red_flag))
                for (int c = tid; c < n2; c += 128) p.stats[(c < p.Co_pad ? c : p.stats_c + c - p.Co_pad) + p.co_base] = stat_tot[c];
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

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// Fused statistics cost the epilogue ~250 cycles per 16 output channels and tile (a transposing shuffle butterfly +
// shared-memory adds); they are fused only where the tile's MMAs hide that (measured: ring kernel with 64 channels +70 %,
// enc.2-like N = 128 / K = 576 tiles +75 %, while 32-channel ring tiles and the K >= 1152 layers hide it).
static bool igemm_fuses_stats(int Ci, int Nc, int taps) {
    const long long mma_cycles = (long long)taps * (Ci / 16) * (Nc / 2 > 32 + Nc / 4 ? Nc / 2 : 32 + Nc / 4);
    return 10 * mma_cycles >= 34LL * (Nc / 16) * 250 && env_int("FV_CONV_FUSE_STATS", 1);
}
// x2 / s2 geometries run the tap schedule with one activation box per tap and are bound by L2 -> shared-memory traffic
// rather than by MMA issue, so their epilogue has slack the MMA-cycle model above does not see: FV_X2_FUSE_STATS = 1 / 0
// forces the fused statistics on / off (default: the model).
static bool geom_fuses_stats(int kind, int Ci, int Nc, int taps, int vtiles) {
    if (kind != CONV_SAME) {
        const int f = env_int("FV_X2_FUSE_STATS", -1);
        if (f >= 0) return f != 0;
        // measured (batch 32, 256^2): fusing costs the x2 kernels 14-21 us per layer and saves a statistic pass of 12 / 14 / 22 /
        // 37 us (up.0 .. up.3): it pays for the large outputs only
        if (kind == CONV_X2 && vtiles >= 2048) return env_int("FV_CONV_FUSE_STATS", 1) != 0;
    }
    return igemm_fuses_stats(Ci, Nc, taps);
}
static int igemm_co_parts(int num_tiles, int Co_pad, int out_mode) {
    int parts = 1;
    while (parts < 4 && num_tiles * parts * 2 <= num_sms() && Co_pad % (parts * 2 * 64) == 0 && out_mode != FV_OUT_NCHW_F32 &&
           env_int("FV_CONV_COSPLIT", 1))
        parts *= 2;
    return parts;
}
static bool ring_fuses_stats(int out_mode, int Co_pad) { return out_mode == FV_OUT_NHWC_BF16 && Co_pad <= 32 && env_int("FV_CONV_FUSE_STATS", 1); }

// fv_conv_ring.cu: sliding-window schedule for thin full-resolution layers; -1 when not eligible
int conv2d_ring_eligible(int out_mode, int H, int W, int Ci, int Co_pad, int R, int S, bool residual);
int conv2d_ring_try(const void* x, const void* w, const float* bias, const void* residual, void* y, int out_mode, int N, int H,
                    int W, int Ci, int Co, int Co_pad, int R, int S, int pad, float* stats, int stats_c, void* red_ws, cudaStream_t stream);

// fv_conv_win.cu: resident-filter / slab-window schedule of the x2 and s2 geometries for thin layers on 128-pixel row tiles
int conv_win_eligible(int kind, int out_mode, int W, int Ci, int Co_pad);
int conv_win_try(int kind, const void* x, const void* w, const float* bias, void* y, int out_mode, int N, int H, int W, int Ci, int Co, int Co_pad,
                 float* stats, void* red_ws, cudaStream_t stream);

}  // namespace fv

// 1 when fv_conv2d_stats produces the statistics inside the convolution's epilogue for this shape, 0 when it runs a separate
// fv_bn_stats pass over y (a host that times kernels individually can then issue the two calls itself).
extern "C" __attribute__((visibility("default"))) int fv_conv2d_fuses_stats(int out_mode, int N, int H, int W, int Ci, int Co_pad, int R, int S,
                                                                           int has_residual) {
    using namespace fv;
    if (out_mode == FV_OUT_NCHW_F32) return 0;
    if (Co_pad > 256) return 1;
    if (conv2d_ring_eligible(out_mode, H, W, Ci, Co_pad, R, S, has_residual != 0)) return ring_fuses_stats(out_mode, Co_pad) ? 1 : 0;
    int tw = 0, th = 0, tn = 0;
    if (pick_tile(N, H, W, tw, th, tn)) return 0;
    const int num_tiles = (W / tw) * (H / th) * ((N + tn - 1) / tn);
    return igemm_fuses_stats(Ci, Co_pad / igemm_co_parts(num_tiles, Co_pad, out_mode), R * S) ? 1 : 0;
}
// the same question for the x2 (kind 1) / s2 (kind 2) geometries; H, W = the coarse tiling grid
extern "C" __attribute__((visibility("default"))) int fv_conv2d_geom_fuses_stats(int kind, int out_mode, int N, int H, int W, int Ci, int Co_pad) {
    using namespace fv;
    if (out_mode == FV_OUT_NCHW_F32 || (kind != CONV_X2 && kind != CONV_S2)) return 0;
    if (conv_win_eligible(kind, out_mode, W, Ci, Co_pad)) return kind == CONV_X2 && env_int("FV_CONV_FUSE_STATS", 1) ? 1 : 0;
    if (Co_pad > 256) return 1;
    int tw = 0, th = 0, tn = 0;
    if (pick_tile(N, H, W, tw, th, tn)) return 0;
    const int nph = kind == CONV_X2 ? 4 : 1;
    const int num_tiles = (W / tw) * (H / th) * ((N + tn - 1) / tn);
    return geom_fuses_stats(kind, Ci, Co_pad / igemm_co_parts(num_tiles * nph, Co_pad, out_mode), kind == CONV_X2 ? 4 : 16, num_tiles * nph) ? 1 : 0;
}
namespace fv {

template <int KB>
static int launch_conv(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvParams& p, size_t smem, int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        FV_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    conv_igemm_kernel<KB><<<grid, kConvThreads, smem, stream>>>(tmX, tmW, p);
    FV_LAUNCH_CHECK("conv_igemm_kernel");
    return FV_OK;
}

struct ConvCall {
    int kind;                       // CONV_SAME / CONV_X2 / CONV_S2
    const void* x; const void* w; const float* bias; const void* residual; void* y;
    int out_mode, N, H, W;          // H, W: the tiling grid (X2 / S2: coarse resolution)
    int Ci, Co, Co_pad, R, S, pad;
    float* stats; void* red_ws; void* stream;
    int act = 0;
};

static int conv_chunk(const ConvCall& c, int co_base, int Co_chunk, int Co_real);

// Output channels beyond one UMMA N (256) are produced in chunks of <= 256: each chunk is an independent GEMM on a row
// slice of the K-major filter matrix, written at its channel offset of the NHWC output (channel stride = Co_pad).
static int conv_any(const ConvCall& c) {
    if (!c.x || !c.w || !c.y) return fail(FV_ERR_ARG, "fv_conv2d: null pointer");
    if (c.stats && !c.red_ws) return fail(FV_ERR_ARG, "fv_conv2d: statistics need the reduction workspace (fv_reduce_ws_bytes)");
    if (c.Co_pad <= 256) return conv_chunk(c, 0, c.Co_pad, c.Co);
    if (c.Co_pad % 64 || c.out_mode == FV_OUT_NCHW_F32)
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d: Co_pad=%d > 256 must be a multiple of 64 with an NHWC output", c.Co_pad);
    for (int c0 = 0; c0 < c.Co_pad; c0 += 256) {
        const int cn = c.Co_pad - c0 < 256 ? c.Co_pad - c0 : 256;
        const int co_real = c.Co - c0 < cn ? (c.Co - c0 > 0 ? c.Co - c0 : 1) : cn;
        if (int e = conv_chunk(c, c0, cn, co_real)) return e;
    }
    return FV_OK;
}

static int conv_chunk(const ConvCall& c, int co_base, int Co_pad, int Co) {
    const int N = c.N, H = c.H, W = c.W, Ci = c.Ci, R = c.R, S = c.S, pad = c.pad, out_mode = c.out_mode, out_cs = c.Co_pad;
    if (N < 1 || H < 1 || W < 1) return fail(FV_ERR_ARG, "fv_conv2d: bad shape N=%d H=%d W=%d", N, H, W);
    if (Ci % 16 || Ci < 16 || (Ci > 64 && Ci % 64) || (Ci < 64 && Ci != 16 && Ci != 32))
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d: Ci=%d must be 16, 32 or a multiple of 64 (pad the channels)", Ci);
    if (Co_pad % 16 || Co_pad < 16 || Co_pad > 256 || Co < 1 || Co > Co_pad)
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d: Co=%d Co_pad=%d (Co_pad must be a multiple of 16 in [16,256])", Co, Co_pad);
    if (c.kind == CONV_SAME && (R != S || (R != 1 && R != 3 && R != 5 && R != 7) || pad != (R - 1) / 2))
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d: only odd square filters with same padding (R=%d S=%d pad=%d)", R, S, pad);
    if (out_mode < 0 || out_mode > 2) return fail(FV_ERR_ARG, "fv_conv2d: out_mode %d", out_mode);
    if (out_mode == FV_OUT_NCHW_F32 && c.residual) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d: residual needs an NHWC output");
    if (c.act < 0 || c.act > 2 || (c.act && c.residual)) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d: activation %d (none / relu / leaky, not with a residual)", c.act);
    float* stats = c.stats;
    cudaStream_t s = (cudaStream_t)c.stream;
    const int osy = c.kind == CONV_X2 ? 2 : 1, Ho = H * osy, Wo = W * osy;
    // statistics: fused into the epilogue where that is hidden (see igemm_fuses_stats), a separate fv_bn_stats pass otherwise
    const int y_dtype = out_mode == FV_OUT_NHWC_BF16 ? FV_DT_BF16 : FV_DT_F32;
    if (c.kind == CONV_SAME && out_cs == Co_pad && c.act == FV_ACT_NONE) {
        float* ring_stats = (stats && ring_fuses_stats(out_mode, Co_pad)) ? stats : nullptr;
        const int rr = conv2d_ring_try(c.x, c.w, c.bias, c.residual, c.y, out_mode, N, H, W, Ci, Co, Co_pad, R, S, pad, ring_stats, out_cs, c.red_ws, s);
        if (rr > 0) return rr;
        if (rr == 0) return (stats && !ring_stats) ? fv_bn_stats(c.y, y_dtype, stats, (long long)N * H * W, Co_pad, c.red_ws, c.stream) : FV_OK;
    }
    if (c.kind != CONV_SAME && out_cs == Co_pad && c.act == FV_ACT_NONE && !c.residual) {
        float* win_stats = (stats && c.kind == CONV_X2 && env_int("FV_CONV_FUSE_STATS", 1)) ? stats : nullptr;
        const int rr = conv_win_try(c.kind, c.x, c.w, c.bias, c.y, out_mode, N, H, W, Ci, Co, Co_pad, win_stats, c.red_ws, s);
        if (rr > 0) return rr;
        if (rr == 0) return (stats && !win_stats) ? fv_bn_stats(c.y, y_dtype, stats, (long long)N * Ho * Wo, Co_pad, c.red_ws, c.stream) : FV_OK;
    }
    ConvParams p{};
    p.N = N; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.Co_pad = Co_pad;
    p.Ho = Ho; p.Wo = Wo; p.osy = osy; p.osx = osy;
    if (pick_tile(N, H, W, p.tw, p.th, p.tn))
        return fail(FV_ERR_UNSUPPORTED, "fv_conv2d: H=%d W=%d not tileable (W multiple of 128, or W,H powers of two)", H, W);
    p.tiles_w = W / p.tw; p.tiles_h = H / p.th; p.tiles_n = (N + p.tn - 1) / p.tn;
    p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int KB = Ci >= 64 ? 64 : Ci;
    const int row_bytes = KB * 2;
    p.kc_blocks = Ci / KB;
    p.nph = c.kind == CONV_X2 ? 4 : 1;
    // fewer pixel tiles than half the SMs (the 16x16 ResBlock2D layers: 64 tiles): split the output channels over 2-4 CTAs
    // per tile -- each streams only its slice of the filter, and twice / four times as many SMs work
    p.co_parts = igemm_co_parts(p.num_tiles * p.nph, Co_pad, out_mode);
    p.Nc = Co_pad / p.co_parts;
    p.num_vtiles = p.num_tiles * p.co_parts * p.nph;
    p.b_slice_stride = (p.Nc * row_bytes + 1023) & ~1023;
    // slab schedule: one activation box per filter row, taps = shifted descriptor views (row-segment tiles of SAME convs only)
    int slab = (c.kind == CONV_SAME && S > 1 && p.th == 1 && p.tn == 1) ? env_int("FV_CONV_SLAB", 1) : 0;
    // a slab stage carries S filter slices: with 256 output channels and a 128-byte K block it no longer fits twice -> tap schedule
    if (slab && 2 * ((((p.tw + S - 1) * row_bytes + 1023) & ~1023) + p.b_slice_stride * S) > 200 * 1024) slab = 0;
    int taps = R * S;
    if (c.kind == CONV_X2) taps = 4;
    if (c.kind == CONV_S2) taps = 16;
    p.nsub = slab ? S : 1;
    p.groups = slab ? R : taps;
    if (p.nph * p.groups > 64) return fail(FV_ERR_INTERNAL, "fv_conv2d: tap table overflow");
    for (int ph = 0; ph < p.nph; ++ph) {
        const int a = ph >> 1, b = ph & 1;
        p.oy[ph] = (signed char)(c.kind == CONV_X2 ? a : 0);
        p.ox[ph] = (signed char)(c.kind == CONV_X2 ? b : 0);
        for (int g = 0; g < p.groups; ++g) {
            short4 t = make_short4(0, 0, 0, 0);
            if (c.kind == CONV_SAME) {
                const int r = slab ? g : g / S, sx = slab ? 0 : g % S;
                t.y = (short)(sx - pad); t.w = (short)(r - pad);
            } else if (c.kind == CONV_X2) {        // phase (a, b), tap (u, v): coarse pixel (i + u - 1 + a, j + v - 1 + b)
                const int u = g >> 1, v = g & 1;
                t.y = (short)(v - 1 + b); t.w = (short)(u - 1 + a);
            } else {                               // S2 tap (r4, s4): fine pixel (2i + r4 - 1, 2j + s4 - 1) = parity + coarse offset
                const int r4 = g >> 2, s4 = g & 3;
                const int fr = r4 - 1, fc = s4 - 1;
                const int dh = fr < 0 ? -1 : fr / 2, dw = fc < 0 ? -1 : fc / 2;
                t.w = (short)dh; t.z = (short)(fr - 2 * dh);
                t.y = (short)dw; t.x = (short)((fc - 2 * dw) * Ci);
            }
            p.tap[ph * p.groups + g] = t;
        }
    }
    const int a_rows = slab ? p.tw + S - 1 : 128;
    p.a_off_b = (a_rows * row_bytes + 1023) & ~1023;
    // K blocks per stage: as many as keep >= 3 stages of <= 64 KB (a stage of one tap x 64 channels is 4 MMAs: the barrier round trips
    // of the issuing thread then cost as much as the MMAs)
    p.kcps = 1;
    for (int k = 4; k >= 2; k >>= 1)
        if (env_int("FV_CONV_KCPS", 1) && p.kc_blocks % k == 0 && (long long)k * (p.a_off_b + p.b_slice_stride * p.nsub) <= 64 * 1024) { p.kcps = k; break; }
    p.stage_stride = p.kcps * (p.a_off_b + p.b_slice_stride * p.nsub);
    p.tx_bytes = a_rows * row_bytes + p.nsub * p.Nc * row_bytes;          // per K block
    p.b_slice_bytes = p.Nc * row_bytes;
    const int groups = p.groups * (p.kc_blocks / p.kcps);
    int stages = (196 * 1024) / p.stage_stride;
    if (stages > 8) stages = 8;
    if (stages > groups) stages = groups;
    if (stages < 2) stages = 2;
    if ((size_t)stages * p.stage_stride > 200 * 1024)
        return fail(FV_ERR_INTERNAL, "fv_conv2d: stage of %d bytes does not fit twice in shared memory", p.stage_stride);
    p.stages = stages;
    p.bar_off = p.stages * p.stage_stride;
    int grid = p.num_vtiles < num_sms() ? p.num_vtiles : num_sms();
    {
        // Filter-resident mode: worth it when the filter slice of one (phase, part) fits beside >= 4 activation stages, every slice
        // gets at least one CTA and a CTA amortises the load over several tiles.  Measured on up.2 forward (x2, 128 -> 64) and the
        // enc.2 data gradient (128 -> 64 at 128 x 128): 805 / 1013 MB through L2 -> SM for 35 / 134 MB inputs, most of it the filter.
        const int slices = p.nph * p.co_parts;
        const long long b_total = (long long)p.groups * p.kc_blocks * p.nsub * p.b_slice_stride;
        int rk = 1;                                              // K blocks per (activation-only) stage in this mode
        for (int k = 4; k >= 2; k >>= 1)
            if (env_int("FV_CONV_KCPS", 1) && p.kc_blocks % k == 0 && (long long)k * p.a_off_b <= 64 * 1024) { rk = k; break; }
        const int a_stage = rk * p.a_off_b;
        const long long room = 218LL * 1024 - b_total - (long long)Co_pad * 52 - 2048;
        int a_stages = (int)(room / a_stage);
        if (a_stages > 8) a_stages = 8;
        const int cps = slices <= num_sms() ? num_sms() / slices : 0;
        if (env_int("FV_CONV_BRES", 1) && b_total <= 160 * 1024 && a_stages * rk >= 4 && a_stages >= 2 && cps >= 1 && p.num_tiles >= 2 * cps) {
            p.b_res = 1;
            p.kcps = rk;
            p.ctas_per_slice = cps;
            p.tiles_per_cta = (p.num_tiles + cps - 1) / cps;
            p.ctas_per_slice = (p.num_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
            p.stage_stride = a_stage;
            p.tx_bytes = a_rows * row_bytes;                     // per K block
            p.stages = a_stages;
            p.b_res_off = p.stages * p.stage_stride;
            p.bar_off = p.b_res_off + (int)b_total;
            grid = slices * p.ctas_per_slice;
        }
    }
    p.out_mode = out_mode;
    int cols = 32;
    while (cols < 2 * p.Nc) cols <<= 1;
    p.tmem_cols = cols;
    p.out_cs = out_cs;
    p.co_base = co_base;
    p.w_rows_per_phase = out_cs;
    p.bias = c.bias ? c.bias + co_base : nullptr;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(c.residual);
    p.out = c.y;
    const bool fuse_stats = stats && (out_cs != Co_pad || geom_fuses_stats(c.kind, Ci, p.Nc, taps, p.num_tiles * p.nph));
    p.stats = fuse_stats ? stats : nullptr;
    p.stats_c = out_cs;                          // the statistic block is [2][total Co_pad] also when Co is walked in chunks
    p.red_ws = c.red_ws;
    p.act = c.act;

    CUtensorMap tmX, tmW;
    {
        // (channels, W, row parity, H, N): SAME / X2 read the tensor as it is (parity dimension of extent 1); S2 reads the fine
        // tensor [N, 2H, 2W, Ci] as (2 Ci, W, 2, H, N) -- column parity folded into the channel coordinate
        const uint64_t fw = c.kind == CONV_S2 ? 2 : 1;
        uint64_t dims[5] = {(uint64_t)Ci * fw, (uint64_t)W, fw, (uint64_t)H, (uint64_t)N};
        uint64_t str[4] = {(uint64_t)Ci * fw * 2, (uint64_t)W * fw * Ci * 2, (uint64_t)W * fw * Ci * 2 * fw, (uint64_t)H * fw * W * fw * Ci * 2};
        uint32_t box[5] = {(uint32_t)KB, (uint32_t)(slab ? p.tw + S - 1 : p.tw), 1, (uint32_t)p.th, (uint32_t)p.tn};
        if (int e = encode_tmap_bf16(&tmX, c.x, 5, dims, str, box, row_bytes)) return e;
    }
    {
        uint64_t dims[2] = {(uint64_t)taps * Ci, (uint64_t)out_cs * p.nph};
        uint64_t str[1] = {(uint64_t)taps * Ci * 2};
        uint32_t box[2] = {(uint32_t)KB, (uint32_t)p.Nc};
        if (int e = encode_tmap_bf16(&tmW, c.w, 2, dims, str, box, row_bytes)) return e;
    }
    const size_t smem = (size_t)p.bar_off + 1024 + 256 + (size_t)Co_pad * 52 + 64;
    const int e = KB == 64 ? launch_conv<64>(tmX, tmW, p, smem, grid, s) : (KB == 32 ? launch_conv<32>(tmX, tmW, p, smem, grid, s) : launch_conv<16>(tmX, tmW, p, smem, grid, s));
    if (e || !stats || fuse_stats) return e;
    return fv_bn_stats(c.y, y_dtype, stats, (long long)N * Ho * Wo, Co_pad, c.red_ws, c.stream);
}

}  // namespace fv

using namespace fv;

extern "C" __attribute__((visibility("default"))) int fv_conv2d(const void* x, const void* w, const float* bias, const void* residual, void* y, int out_mode,
                         int N, int H, int W, int Ci, int Co, int Co_pad, int R, int S, int pad, void* stream) {
    ConvCall c{CONV_SAME, x, w, bias, residual, y, out_mode, N, H, W, Ci, Co, Co_pad, R, S, pad, nullptr, nullptr, stream};
    return conv_any(c);
}

// The same with the batch-norm statistics of the output fused into the epilogue: stats[0..Co_pad) = sum_pixels y,
// stats[Co_pad..2 Co_pad) = sum_pixels y^2 of the values as stored (written, not accumulated; NHWC outputs only).
extern "C" __attribute__((visibility("default"))) int fv_conv2d_stats(const void* x, const void* w, const float* bias, const void* residual, void* y,
                               int out_mode, int N, int H, int W, int Ci, int Co, int Co_pad, int R, int S, int pad, float* stats,
                               void* red_ws, void* stream) {
    if (!stats) return fail(FV_ERR_ARG, "fv_conv2d_stats: null stats pointer");
    if (out_mode == FV_OUT_NCHW_F32) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_stats: NHWC outputs only");
    ConvCall c{CONV_SAME, x, w, bias, residual, y, out_mode, N, H, W, Ci, Co, Co_pad, R, S, pad, stats, red_ws, stream};
    return conv_any(c);
}

// x [N,H,W,Ci] -> y [N,2H,2W,Co_pad]: nearest 2x up-sampling + 3x3 "same" convolution as four 2x2 phase convolutions on the
// coarse grid (wp = [4][Co_pad][4*Ci] from fv_weight_prep_up), or the data gradient of a 4x4 stride-2 convolution.
extern "C" __attribute__((visibility("default"))) int fv_conv2d_x2(const void* x, const void* wp, const float* bias, void* y, int out_mode, int N, int H, int W,
                                                                 int Ci, int Co, int Co_pad, float* stats, void* red_ws, void* stream) {
    if (stats && out_mode == FV_OUT_NCHW_F32) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_x2: statistics need an NHWC output");
    ConvCall c{CONV_X2, x, wp, bias, nullptr, y, out_mode, N, H, W, Ci, Co, Co_pad, 2, 2, 0, stats, red_ws, stream};
    return conv_any(c);
}

// x [N,2H,2W,Ci] -> y [N,H,W,Co_pad]: 4x4 stride-2 pad-1 convolution (w = [Co_pad][16*Ci], taps row-major), or -- with the
// summed, mirrored taps of fv_weight_prep_up -- the data gradient of fv_conv2d_x2 (up-sampling backward folded in).
extern "C" __attribute__((visibility("default"))) int fv_conv2d_s2(const void* x, const void* w, const float* bias, void* y, int out_mode, int N, int H, int W,
                                                                 int Ci, int Co, int Co_pad, float* stats, void* red_ws, void* stream) {
    if (stats && out_mode == FV_OUT_NCHW_F32) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_s2: statistics need an NHWC output");
    ConvCall c{CONV_S2, x, w, bias, nullptr, y, out_mode, N, H, W, Ci, Co, Co_pad, 4, 4, 1, stats, red_ws, stream};
    return conv_any(c);
}

// The general entry: geometry kind (0 same, 1 x2, 2 s2), optional activation in the epilogue (conv -> bias -> act, the order of
// Conv2dELR.forward, reference models_utils.py:712-742), optional fused statistics.  H, W: the tiling grid (x2 / s2: coarse).
extern "C" __attribute__((visibility("default"))) int fv_conv2d_ex(int kind, const void* x, const void* w, const float* bias, const void* residual, void* y,
                                                                 int out_mode, int N, int H, int W, int Ci, int Co, int Co_pad, int R, int S, int pad, int act,
                                                                 float* stats, void* red_ws, void* stream) {
    if (kind < 0 || kind > 2) return fail(FV_ERR_ARG, "fv_conv2d_ex: kind %d", kind);
    if (kind != CONV_SAME && residual) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_ex: residual only with the same-size geometry");
    if (stats && out_mode == FV_OUT_NCHW_F32) return fail(FV_ERR_UNSUPPORTED, "fv_conv2d_ex: statistics need an NHWC output");
    ConvCall c{kind, x, w, bias, residual, y, out_mode, N, H, W, Ci, Co, Co_pad, kind == CONV_X2 ? 2 : (kind == CONV_S2 ? 4 : R),
               kind == CONV_X2 ? 2 : (kind == CONV_S2 ? 4 : S), kind == CONV_X2 ? 0 : (kind == CONV_S2 ? 1 : pad), stats, red_ws, stream};
    c.act = act;
    return conv_any(c);
}
