// Sliding-window schedule of the weight-gradient GEMM for the thin full-resolution layers (Ci <= 64, Co_pad <= 64):
//
//   dW[co, (r,s), ci] = sum_pixels X[pixel + (r - pad, s - pad), ci] * dY[pixel, co]
//
// The plain wgrad kernel (fv_wgrad.cu) fetches one shifted activation box per filter tap and pixel block, i.e. it
// pulls the input R*S times through L2 -> SMEM; for the 7x7 out_conv that is 49x and the kernel is L2-bound.  Here a
// CTA walks a run of vertically adjacent 64-pixel row segments and keeps the last R input-row slabs of (64 + S - 1)
// pixels in a shared-memory ring: per segment ONE new slab and one dY tile are fetched.  The S taps of a filter row
// are the same slab read through MN-major UMMA descriptors whose start address is shifted by s pixel rows; the
// 128/cw taps stacked in one 128-row M tile are consecutive shifts, expressed by a leading-dimension byte offset of
// one pixel row.  All R * ceil(S / (128/cw)) accumulators live in TMEM for the whole run; the epilogue STORES them into this
// CTA's slab of the workspace [CTAs][Co_pad][R*S][Ci]; fv_wgrad_finish adds the slabs in CTA order (reproducible).
#include <cstdio>
#include <cstdlib>

#include "../../include/facevae_b200.h"
#include "fv_host.h"
#include "fv_ptx.cuh"

namespace fv {

struct WRingParams {
    int N, H, W, Ci, Co_pad, R, S, pad, taps;
    int cpt, tiles_per_r, mt_total;
    int cols_w, num_blocks, blocks_per_cta;
    int ring, slab_stride, slab_tx;
    int b_off, b_slots, b_stride, b_tx;
    int bar_off, tmem_cols;
    float* dw;
    long long split_stride;   // elements between the slabs of consecutive CTAs
    // D[(tap t, ci), co] -> dw[cta * split_stride + ci + t * st + co * sb]: st = total input channels, sb = taps * st (dw already points at this
    // launch's channel chunk when the wide operand is walked in chunks of 64)
    long long st, sb;
    // x2 mode (weight gradient of the up-sampling convolution, x2_co > 0): dY is the parity-x2_a rows of the fine gradient seen as
    // [N][H][W][(b, co)] and only the filter rows r_lo <= r < r_hi are accumulated; the epilogue scatters D[(r, s), (b, co)] into
    // the phase layout [4][x2_co][4][Ci] of fv_conv2d_wgrad_x2 (phase (a, b), tap (u, v) = (r - a, s - b))
    int r_lo, r_hi, x2_a, x2_co;
    long long* trace;
};

static constexpr int kWRingThreads = 224;   // producer, issuer A, 4 epilogue warps, issuer B

// CW: channels per A chunk == Ci (16 / 32 / 64); S_: filter size (R == S); kPX: pixels (GEMM K) per block
template <int CW, int S_, int kPX>
__global__ void __launch_bounds__(kWRingThreads, 1)
conv_wgrad_ring_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WRingParams p) {
    constexpr int AROW = CW * 2;                                   // bytes per pixel row of the activation slab
    constexpr uint32_t A_LAYOUT = AROW == 128 ? 2u : (AROW == 64 ? 4u : 6u);
    constexpr int CPT = 128 / CW;                                  // taps stacked in one 128-row M tile
    constexpr int TPR = (S_ + CPT - 1) / CPT;                      // M tiles per filter row

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.bar_off);   // [ring]   slab landed
    uint64_t* empty = full + p.ring;                                   // [ring]   slab no longer read
    uint64_t* bfull = empty + p.ring;                                  // [b_slots] dY tile landed
    uint64_t* bempty = bfull + p.b_slots;                              // [b_slots]
    uint64_t* tfull = bempty + p.b_slots;                              // accumulators complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef FV_TRACE
    long long* fv_trace = p.trace;
#endif
    const int b0 = blockIdx.x * p.blocks_per_cta;
    const int b1 = min(b0 + p.blocks_per_cta, p.num_blocks);
    const int brow = p.Co_pad * 2;                                     // bytes per pixel row of the dY tile

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDY);
        // two issuer warps (even / odd M tiles) read every slab and every dY tile: "empty" needs both votes
        for (int i = 0; i < p.ring; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 2);
        }
        for (int i = 0; i < p.b_slots; ++i) {
            mbar_init(&bfull[i], 1);
            mbar_init(&bempty[i], 2);
        }
        mbar_init(tfull, 2);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    // block b -> column (n, 64-pixel segment) = b / H, image row h = b % H: consecutive blocks are vertically adjacent
    if (warp == 0) {
        if (b0 < b1) {
            const bool leader = elect_one_sync();
            uint32_t slot = 0, ph = 0, bs = 0, bph = 0;
            int col = b0 / p.H, h = b0 - col * p.H;
            bool fresh = true;
            for (int b = b0; b < b1; ++b) {
                const int n = col / p.cols_w, w0 = (col - n * p.cols_w) * kPX;
                { FV_T0(tw); mbar_wait(&bempty[bs], bph ^ 1); FV_TACC(0, tw); }
                if (leader) {
                    mbar_arrive_expect_tx(&bfull[bs], (uint32_t)p.b_tx);
                    tma_load_4d(smem + p.b_off + (size_t)bs * p.b_stride, &tmDY, &bfull[bs], 0, w0, h, n);
                }
                if (++bs == (uint32_t)p.b_slots) { bs = 0; bph ^= 1; }
                for (int j = fresh ? 0 : p.R - 1; j < p.R; ++j) {
                    { FV_T0(tw); mbar_wait(&empty[slot], ph ^ 1); FV_TACC(1, tw); }
                    if (leader) {
                        mbar_arrive_expect_tx(&full[slot], (uint32_t)p.slab_tx);
                        tma_load_4d(smem + (size_t)slot * p.slab_stride, &tmX, &full[slot], 0, w0 - p.pad, h + j - p.pad, n);
                    }
                    if (++slot == (uint32_t)p.ring) { slot = 0; ph ^= 1; }
                }
                fresh = false;
                if (++h == p.H) { h = 0; ++col; fresh = true; }
            }
        }
    } else if (warp == 1 || warp == 6) {
        // Two issuer warps: a single warp spends ~40 % of a block on barrier observation and commits while the tensor
        // pipe drains; with the M tiles split by parity the other warp's MMAs fill those gaps.
        const uint32_t issuer = warp == 6 ? 1u : 0u;
        if (b0 < b1) {
            const bool leader = elect_one_sync();
            const uint32_t idesc = umma_idesc_bf16(128, p.Co_pad, 1, 1);           // both operands MN-major
            const uint32_t b_layout = umma_layout_code(brow);
            // A: chunk i of an M tile = the slab shifted by i more pixel rows -> LBO = one pixel row
            const uint64_t a_tmpl = umma_smem_desc(0, AROW, 8u * AROW, A_LAYOUT);
            const uint64_t b_tmpl = umma_smem_desc(0, 16, 8u * brow, b_layout);
            const uint32_t a_hi = (uint32_t)(a_tmpl >> 32), a_lo_base = (uint32_t)a_tmpl;
            const uint32_t b_hi = (uint32_t)(b_tmpl >> 32), b_lo_base = (uint32_t)b_tmpl;
            const uint32_t smem_base = smem_u32(smem);
            const uint32_t a_kstep = (16u * AROW) >> 4, b_kstep = (16u * (uint32_t)brow) >> 4;   // 16 pixel rows per MMA
            uint32_t first = 0, wait_slot = 0, wait_ph = 0, bs = 0, bph = 0, accumulate = 0;
            int h = b0 % p.H;
            bool fresh = true;
            FV_T0(t_all);
            // the barrier waits of block b + 1 are taken in the middle of block b's MMAs (see fv_conv_ring.cu)
            uint32_t wbs = 0, wbph = 0;
            auto wait_block = [&](bool is_fresh) {
                const int n_new = is_fresh ? p.R : 1;
                FV_T0(tw);
                for (int i = 0; i < n_new; ++i) {
                    mbar_wait(&full[wait_slot], wait_ph);
                    if (++wait_slot == (uint32_t)p.ring) { wait_slot = 0; wait_ph ^= 1; }
                }
                FV_TACC(2, tw);
                FV_T0(tw2);
                mbar_wait(&bfull[wbs], wbph);
                if (++wbs == (uint32_t)p.b_slots) { wbs = 0; wbph ^= 1; }
                FV_TACC(3, tw2);
            };
            bool waited = false;
            for (int b = b0; b < b1; ++b) {
                const bool next_fresh = (h + 1 == p.H);
                // a column change needs R fresh slabs, i.e. slots this block still occupies: that wait cannot be taken early
                if (!waited) wait_block(fresh);
                waited = false;
                tc_fence_after();
                // non-blocking probes of block b + 1's barriers before this block's MMAs, consumed in their middle (fv_conv_ring.cu)
                uint32_t probe_full = 0, probe_b = 0;
                if (b + 1 < b1 && !next_fresh) {
                    probe_full = mbar_test_wait(&full[wait_slot], wait_ph);
                    probe_b = mbar_test_wait(&bfull[wbs], wbph);
                }
                FV_T0(t_issue);
                const uint32_t b_lo = b_lo_base | ((smem_base + (uint32_t)p.b_off + bs * (uint32_t)p.b_stride) >> 4);
                uint32_t slot = first;
#pragma unroll
                for (int r = 0; r < S_; ++r) {
                    const uint32_t a_row = a_lo_base | ((smem_base + slot * (uint32_t)p.slab_stride) >> 4);
#pragma unroll
                    for (int part = 0; part < TPR; ++part) {
                        const uint32_t d_col = tmem_base + (uint32_t)((r * TPR + part) * p.Co_pad);
#pragma unroll
                        for (int k4 = 0; k4 < kPX / 16; ++k4)
                            if (leader && (uint32_t)((r * TPR + part) & 1) == issuer && r >= p.r_lo && r < p.r_hi)
                                tc_mma_f16_lohi2(d_col, a_row + (uint32_t)(part * CPT * (AROW >> 4)) + k4 * a_kstep, a_hi, b_lo + k4 * b_kstep,
                                                 b_hi, idesc, accumulate | (uint32_t)(k4 > 0));
                    }
                    if (++slot == (uint32_t)p.ring) slot = 0;
                    if (r == (S_ - 1) / 2 && b + 1 < b1 && !next_fresh) {
                        if (!probe_full) mbar_wait(&full[wait_slot], wait_ph);
                        if (++wait_slot == (uint32_t)p.ring) { wait_slot = 0; wait_ph ^= 1; }
                        if (!probe_b) mbar_wait(&bfull[wbs], wbph);
                        if (++wbs == (uint32_t)p.b_slots) { wbs = 0; wbph ^= 1; }
                        waited = true;
                    }
                }
                accumulate = 1;
                FV_TACC(4, t_issue);
                if (leader) tc_commit(&bempty[bs]);
                if (++bs == (uint32_t)p.b_slots) { bs = 0; bph ^= 1; }
                const int n_rel = (b + 1 < b1) ? (next_fresh ? p.R : 1) : 0;
                for (int i = 0; i < n_rel; ++i) {
                    if (leader) tc_commit(&empty[first]);
                    if (++first == (uint32_t)p.ring) first = 0;
                }
                fresh = next_fresh;
                if (++h == p.H) h = 0;
            }
            if (leader) tc_commit(tfull);
            FV_TACC(5, t_all);
        }
    } else if (b0 < b1) {       // warps 2..5
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int chunk = row / CW, ci = row % CW;
        mbar_wait(tfull, 0);
        tc_fence_after();
        FV_T0(t_epi);
        int mt = 0;
        for (int r = 0; r < p.R; ++r)
            for (int part = 0; part < p.tiles_per_r; ++part, ++mt) {
                const int s = part * p.cpt + chunk;
                const bool valid = s < p.S;
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + (uint32_t)mt * p.Co_pad;
                if (r < p.r_lo || r >= p.r_hi) continue;
                if (p.x2_co) {
                    const int u = r - p.x2_a;
                    float* dst = p.dw + (size_t)blockIdx.x * p.split_stride + ci;
                    for (int c0 = 0; c0 < p.Co_pad; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(taddr + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int b = (c0 + i) / p.x2_co, co = (c0 + i) - b * p.x2_co, vv = s - b;
                            if (valid && vv >= 0 && vv < 2)
                                dst[((size_t)((p.x2_a * 2 + b) * p.x2_co + co) * 4 + (u * 2 + vv)) * CW] = __uint_as_float(v[i]);
                        }
                    }
                    continue;
                }
                const int t = r * p.S + s;
                float* dst = p.dw + (size_t)blockIdx.x * p.split_stride + ci + (size_t)t * p.st;
                for (int c0 = 0; c0 < p.Co_pad; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c0, v);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) dst[(size_t)(c0 + i) * p.sb] = __uint_as_float(v[i]);
                    }
                }
            }
        if (warp == 2) FV_TACC(6, t_epi);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

template <int CW, int S_, int kPX>
static int launch_wring(const CUtensorMap& tmX, const CUtensorMap& tmDY, const WRingParams& p, size_t smem, int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        FV_CUDA(cudaFuncSetAttribute(conv_wgrad_ring_kernel<CW, S_, kPX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    conv_wgrad_ring_kernel<CW, S_, kPX><<<grid, kWRingThreads, smem, stream>>>(tmX, tmDY, p);
    FV_LAUNCH_CHECK("conv_wgrad_ring_kernel");
    return FV_OK;
}

// FV_OK after launching, -1 when not eligible (caller falls through to the generic wgrad kernel).
// x / dy: NHWC operands of which Ci <= 64 / Co_pad <= 64 channels starting at the pointers are used; x_cs / dy_cs = their channel
// strides in elements.  Output mapping: WRingParams.
static int wring_grid(int N, int H, int W, int kPX) {
    const int num_blocks = N * (W / kPX) * H, sms = num_sms();
    const int per = (num_blocks + sms - 1) / sms;
    return (num_blocks + per - 1) / per;
}
static bool wring_shape_ok(int W, int Ci, int Co_pad, int R, int S) {
    if ((S != 3 && S != 5 && S != 7) || R != S || W % 64 || Ci > 64 || Co_pad > 64) return false;
    const int cpt = 128 / Ci, kPX = (W % 128 == 0) ? 128 : 64;
    if (R * ((S + cpt - 1) / cpt) * Co_pad > 512) return false;
    const int slab_stride = ((kPX + S - 1 + cpt) * Ci * 2 + 1023) & ~1023, b_stride = (kPX * Co_pad * 2 + 1023) & ~1023;
    const size_t smem = (size_t)(R + 3) * slab_stride + 4 * b_stride + (2 * (R + 3) + 2 * 4 + 2) * 8 + 16 + 1024 + 64;
    const char* env = getenv("FV_WGRAD_RING");
    return smem <= 225 * 1024 && !(env && atoi(env) == 0);
}

struct WRingX2 {            // x2 mode of conv2d_wgrad_ring_ex (see WRingParams)
    int a, co;
    long long dy_row, dy_img;   // element strides of the dY view between rows / images
};

int conv2d_wgrad_ring_ex(const void* x, int x_cs, const void* dy, int dy_cs, float* dw_acc, long long split_stride, long long st, long long sb, int N,
                         int H, int W, int Ci, int Co_pad, int R, int S, int pad, cudaStream_t stream, const WRingX2* x2 = nullptr) {
    if (!wring_shape_ok(W, Ci, Co_pad, R, S)) return -1;
    const int kPX = (W % 128 == 0) ? 128 : 64;
    WRingParams p{};
    p.N = N; p.H = H; p.W = W; p.Ci = Ci; p.Co_pad = Co_pad; p.R = R; p.S = S; p.pad = pad; p.taps = R * S;
    p.cpt = 128 / Ci;
    p.tiles_per_r = (S + p.cpt - 1) / p.cpt;
    p.mt_total = R * p.tiles_per_r;
    if (p.mt_total * Co_pad > 512) return -1;
    p.cols_w = W / kPX;
    p.num_blocks = N * p.cols_w * H;
    const int arow = Ci * 2, brow = Co_pad * 2;
    p.slab_tx = (kPX + S - 1) * arow;
    // the last (partly unused) M tile of a filter row reads up to cpt - 1 pixel rows past its slab: keep them inside the stride
    p.slab_stride = ((kPX + S - 1 + p.cpt) * arow + 1023) & ~1023;
    p.ring = R + 3;
    p.b_slots = 4;
    p.b_tx = kPX * brow;
    p.b_stride = (p.b_tx + 1023) & ~1023;
    p.b_off = p.ring * p.slab_stride;
    p.bar_off = p.b_off + p.b_slots * p.b_stride;
    const size_t smem = (size_t)p.bar_off + (2 * p.ring + 2 * p.b_slots + 2) * 8 + 16 + 1024 + 64;
    if (smem > 225 * 1024) return -1;
    int cols = 32;
    while (cols < p.mt_total * Co_pad) cols <<= 1;
    p.tmem_cols = cols;
    p.dw = dw_acc;
    p.split_stride = split_stride;
    p.st = st; p.sb = sb;
    p.r_lo = 0; p.r_hi = R;
    if (x2) { p.r_lo = x2->a; p.r_hi = x2->a + 2; p.x2_a = x2->a; p.x2_co = x2->co; }
    p.trace = trace_ptr();
    const int sms = num_sms();
    p.blocks_per_cta = (p.num_blocks + sms - 1) / sms;
    const int grid = (p.num_blocks + p.blocks_per_cta - 1) / p.blocks_per_cta;

    CUtensorMap tmX, tmDY;
    {
        uint64_t dims[4] = {(uint64_t)Ci, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)x_cs * 2, (uint64_t)W * x_cs * 2, (uint64_t)H * W * x_cs * 2};
        uint32_t box[4] = {(uint32_t)Ci, (uint32_t)(kPX + S - 1), 1, 1};
        if (int e = encode_tmap_bf16(&tmX, x, 4, dims, str, box, arow)) return e;
    }
    {
        uint64_t dims[4] = {(uint64_t)Co_pad, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)dy_cs * 2, (uint64_t)W * dy_cs * 2, (uint64_t)H * W * dy_cs * 2};
        if (x2) { str[1] = (uint64_t)x2->dy_row * 2; str[2] = (uint64_t)x2->dy_img * 2; }
        uint32_t box[4] = {(uint32_t)Co_pad, (uint32_t)kPX, 1, 1};
        if (int e = encode_tmap_bf16(&tmDY, dy, 4, dims, str, box, brow)) return e;
    }
#define FV_WR2(CW_, S__) (kPX == 128 ? launch_wring<CW_, S__, 128>(tmX, tmDY, p, smem, grid, stream) : launch_wring<CW_, S__, 64>(tmX, tmDY, p, smem, grid, stream))
#define FV_WR(CW_) (S == 3 ? FV_WR2(CW_, 3) : (S == 5 ? FV_WR2(CW_, 5) : FV_WR2(CW_, 7)))
    if (Ci == 64) return FV_WR(64);
    if (Ci == 32) return FV_WR(32);
    return FV_WR(16);
#undef FV_WR
#undef FV_WR2
}

// Which schedule a shape gets: 1 both channel counts <= 64, 2 dY walked in chunks of 64, 3 x walked in chunks of 64, 0 none.
static int wring_mode(int W, int Ci, int Co_pad, int R, int S) {
    if (Ci <= 64 && Co_pad <= 64) return wring_shape_ok(W, Ci, Co_pad, R, S) ? 1 : 0;
    const char* env = getenv("FV_WGRAD_RING_WIDE");
    if (env && atoi(env) == 0) return 0;
    if (Ci <= 64 && Co_pad % 64 == 0 && Co_pad <= 256) return wring_shape_ok(W, Ci, 64, R, S) ? 2 : 0;
    if (Co_pad <= 64 && Ci % 64 == 0 && Ci <= 256) return wring_shape_ok(W, 64, Co_pad, R, S) ? 3 : 0;
    return 0;
}

// ---- x2: weight gradient of the nearest-2x up-sampling 3x3 convolution on the coarse grid ---------------------------------------
// With y = 2h + a, x = 2w + b the fine pixel (y + kh - 1, x + kw - 1) of the up-sampled input is the coarse pixel
// (h + floor((a + kh - 1) / 2), w + floor((b + kw - 1) / 2)): per row parity a the phase-filter gradients are entries of an ordinary
// 3x3 "same" weight gradient between the coarse input and the parity-a rows of dY read as [N][H][W][(b, co)] (a plain strided NHWC
// view: row stride 4 W Co, pixel stride 2 Co).  Two launches (a = 0, 1) of the ring kernel with N = 2 Co columns, filter rows
// {a, a + 1}; the epilogue keeps column shift s for parity b when v = s - b is 0 or 1.
static bool wring_x2_ok(int W, int Ci, int Co_pad) {
    const char* env = getenv("FV_WGRAD_RING_X2");
    if (env && atoi(env) == 0) return false;
    if (Ci != 16 && Ci != 32 && Ci != 64) return false;
    if (Co_pad != 16 && Co_pad != 32) return false;
    return wring_shape_ok(W, Ci, 2 * Co_pad, 3, 3);
}
int conv2d_wgrad_ring_x2_splits(int N, int H, int W, int Ci, int Co_pad) {
    if (!wring_x2_ok(W, Ci, Co_pad)) return 0;
    return wring_grid(N, H, W, (W % 128 == 0) ? 128 : 64);
}
// part[cta][4][Co_pad][4][Ci]; -1 when not eligible
int conv2d_wgrad_ring_x2_try(const void* x, const void* dy, float* part, long long split_stride, int N, int H, int W, int Ci, int Co_pad,
                             cudaStream_t stream) {
    if (!wring_x2_ok(W, Ci, Co_pad)) return -1;
    for (int a = 0; a < 2; ++a) {
        WRingX2 x2{a, Co_pad, 4LL * W * Co_pad, 4LL * H * W * Co_pad};
        const int e = conv2d_wgrad_ring_ex(x, Ci, static_cast<const char*>(dy) + (size_t)a * 2 * W * Co_pad * 2, 2 * Co_pad, part, split_stride, 0, 0, N,
                                           H, W, Ci, 2 * Co_pad, 3, 3, 1, stream, &x2);
        if (e) return e < 0 ? fail(FV_ERR_INTERNAL, "fv_conv2d_wgrad_x2: ring schedule refused a shape it had accepted") : e;
    }
    return FV_OK;
}

// number of partial slabs (= CTAs) the ring schedule writes for this shape; 0 when it does not take the shape
int conv2d_wgrad_ring_splits(int N, int H, int W, int Ci, int Co_pad, int R, int S) {
    if (!wring_mode(W, Ci, Co_pad, R, S)) return 0;
    return wring_grid(N, H, W, (W % 128 == 0) ? 128 : 64);
}

// part[cta][Co_pad][R*S][Ci] = this CTA's share.  Ring schedule whenever one of the two channel counts is <= 64: the other
// operand is walked in chunks of 64 channels, one launch per chunk (each pass re-reads the narrow operand).
int conv2d_wgrad_ring_try(const void* x, const void* dy, float* part, long long split_stride, int N, int H, int W, int Ci, int Co_pad, int R, int S,
                          int pad, cudaStream_t stream) {
    const long long taps = (long long)R * S;
    const int mode = wring_mode(W, Ci, Co_pad, R, S);
    if (mode == 0) return -1;
    if (mode == 1) return conv2d_wgrad_ring_ex(x, Ci, dy, Co_pad, part, split_stride, Ci, taps * Ci, N, H, W, Ci, Co_pad, R, S, pad, stream);
    if (mode == 2) {                                                         // dY in chunks of 64 channels
        for (int c0 = 0; c0 < Co_pad; c0 += 64) {
            const int e = conv2d_wgrad_ring_ex(x, Ci, static_cast<const char*>(dy) + (size_t)c0 * 2, Co_pad, part + (size_t)c0 * taps * Ci, split_stride,
                                               Ci, taps * Ci, N, H, W, Ci, 64, R, S, pad, stream);
            if (e) return e < 0 ? fail(FV_ERR_INTERNAL, "fv_conv2d_wgrad: ring schedule refused a chunk it had accepted") : e;
        }
        return FV_OK;
    }
    for (int c0 = 0; c0 < Ci; c0 += 64) {                                    // x in chunks of 64 channels
        const int e = conv2d_wgrad_ring_ex(static_cast<const char*>(x) + (size_t)c0 * 2, Ci, dy, Co_pad, part + c0, split_stride, Ci, taps * Ci, N, H, W,
                                           64, Co_pad, R, S, pad, stream);
        if (e) return e < 0 ? fail(FV_ERR_INTERNAL, "fv_conv2d_wgrad: ring schedule refused a chunk it had accepted") : e;
    }
    return FV_OK;
}

}  // namespace fv
