// "Window" schedule of the x2 / s2 convolution geometries (fv_conv.cu) for the thin full-resolution layers -- UpBlock2D(64 -> 32)
// at 256x256 and its data gradient: the layers where the tap schedule of conv_igemm_kernel re-reads every activation box through
// L2 -> shared memory once per (phase, tap) and runs at a third of the tensor-core issue rate.
//
// Like conv_ring_kernel (fv_conv_ring.cu): a tile is a 128-pixel row segment of the tiling (coarse) grid, a CTA walks a run of
// vertically adjacent tiles, the filter is resident in shared memory for the whole kernel, and the activation rows ("lines")
// the tile needs live in a ring of shared-memory slabs of 128 + 2 pixels -- consecutive tiles share most of their lines, so
// only `adv` new slabs are fetched per tile and the L2 -> SMEM traffic equals the layer's input once.  Horizontal taps are
// row-shifted UMMA descriptor views of a slab.  What differs from the ring kernel is described by a small table in the kernel
// parameters, one entry per tcgen05.mma group: which slab of the window, which pixel shift, which K sub-range of the slab's
// channels, which filter slice and which accumulator region:
//
//   x2 (UpBlock2D forward; window of J = 3 coarse rows, adv = 1): 16 entries = 4 phases x 4 taps; phase (a, b), tap (u, v) reads
//      slab u + a at shift v + b with filter slice (phase, tap) of wx2 and accumulates into TMEM region `phase`; the epilogue
//      scatters region (a, b) to the output pixels (2h + a, 2w + b);
//   s2 (its data gradient = 4x4 stride-2 conv of dY; window of J = 4 FINE rows 2h-1 .. 2h+2, adv = 2): a fine row is fetched
//      through the 5-D parity view (2C, W, 2, H, N) as a slab of coarse pixels whose 2C channels are (column parity b, c); tap
//      (r4, s4) reads slab r4 at shift dw + 1, K sub-range b*C .. b*C + C of the slab's channels, filter slice r4*4 + s4 of ws2.
//
// Warp roles, mbarrier protocol, TMEM double buffering and the deterministic fused statistics are those of fv_conv.cu.
#include <cstdio>
#include <cstdlib>

#include "../../include/facevae_b200.h"
#include "fv_host.h"
#include "fv_ptx.cuh"
#include "fv_reduce.cuh"

namespace fv {

struct WinTap {
    unsigned char slab, shift, koff16, ksteps;     // slab of the window, pixel shift, K offset (16-byte units) and K steps inside a slab row
    unsigned short wslice, dcol;                   // resident filter slice, TMEM column offset of the accumulator region
};

struct WinParams {
    int N, H, W;                   // tiling grid (coarse); tiles are 128-pixel row segments
    int Ho, Wo, osy, osx;          // output geometry: region r of tile pixel (h, w) -> (h*osy + oy[r], w*osx + ox[r])
    int Co, Co_pad;                // real / stored channels per output pixel
    int nacc, Nmma;                // accumulator regions per tile, UMMA N
    int tiles_w, num_tiles, tiles_per_cta;
    int J, adv, row0, parity_lines;
    int ring, slab_stride, slab_tx, halo;
    int arow, brow;                // bytes per pixel row of a slab / per row of a filter slice
    int w_off, w_slices, w_slice_stride, w_slice_tx;
    int bar_off, tmem_cols;
    int ntaps;
    WinTap tap[16];
    short wk[16], wr[16];          // TMA coordinates (K element, row) of filter slice i
    signed char oy[4], ox[4];
    const float* bias;
    void* out;
    float* stats;
    int stats_c;
    void* red_ws;
};

static constexpr int kWinThreads = 192;

// NM > 0: Nmma == 16 * NM is a compile-time constant, the epilogue's channel loop is unrolled and the optional batch-norm
// statistics are accumulated PER THREAD in registers across all tiles of the CTA (lane = pixel; reduced across lanes once, at the
// end) -- a per-tile shuffle butterfly would make the epilogue (8 chunks per tile) slower than the tile's MMAs.  NM == 0: run-time
// channel loop, no statistics (the data-gradient use).
// KIND (1 = x2, 2 = s2) makes the 16-entry MMA table a compile-time function of the entry index: the issuing thread's loop is
// fully unrolled and carries no table loads, divisions or address arithmetic beyond one add per MMA -- with the table read from the
// kernel parameters at run time the scalar work per MMA group (~300 cycles) was three times the MMA issue itself.
template <int KIND>
struct WinTapC {
    // x2: entry i = phase * 4 + tap, phase = (a, b), tap = (u, v);  s2: entry i = r4 * 4 + s4
    static __host__ __device__ __forceinline__ constexpr int slab(int i) { return KIND == 1 ? ((i >> 1) & 1) + (i >> 3) : (i >> 2); }
    static __host__ __device__ __forceinline__ constexpr int shift(int i) { return KIND == 1 ? (i & 1) + ((i >> 2) & 1) : ((i & 3) == 0 ? 0 : ((i & 3) == 3 ? 2 : 1)); }
    static __host__ __device__ __forceinline__ constexpr int khalf(int i) { return KIND == 1 ? 0 : (((i & 3) == 0 || (i & 3) == 2) ? 1 : 0); }   // column parity b
    static __host__ __device__ __forceinline__ constexpr int region(int i) { return KIND == 1 ? (i >> 2) : 0; }
};

// KS = K steps of 16 channels per table entry (Ci / 16), compile time as well: the issue loop is straight-line code.
template <int NM, int KIND, int KS>
__global__ void __launch_bounds__(kWinThreads, 1)
conv_win_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ WinParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.bar_off);      // [ring]
    uint64_t* empty = full + p.ring;                                      // [ring]
    uint64_t* wbar = empty + p.ring;
    uint64_t* tfull = wbar + 1;                                           // [2]
    uint64_t* tempty = tfull + 2;                                         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    int* red_flag = reinterpret_cast<int*>(tmem_slot + 2);
    float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);              // [Nmma]
    float* stat_w = bias_s + p.Nmma;                                      // [4 warps][2][Nmma]
    float* stat_blk = stat_w + 8 * p.Nmma;
    float* stat_tot = stat_blk + 2 * p.Nmma;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * p.tiles_per_cta;
    const int t1 = min(t0 + p.tiles_per_cta, p.num_tiles);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        for (int i = 0; i < p.ring; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(wbar, 1);
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
    for (int c = threadIdx.x; c < p.Nmma; c += blockDim.x) bias_s[c] = (p.bias && c < p.Co) ? p.bias[c] : 0.f;
    if (p.stats)
        for (int c = threadIdx.x; c < 8 * p.Nmma; c += blockDim.x) stat_w[c] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile t -> column (n, w segment) = t / H, tiling row h = t % H: consecutive tiles are vertically adjacent
    if (warp == 0) {
        if (t0 < t1) {
            const bool leader = elect_one_sync();
            if (leader) mbar_arrive_expect_tx(wbar, (uint32_t)(p.w_slices * p.w_slice_tx));
            for (int i = 0; i < p.w_slices; ++i)
                if (leader) tma_load_2d(smem + p.w_off + i * p.w_slice_stride, &tmW, wbar, p.wk[i], p.wr[i]);
            uint32_t slot = 0, ph = 0;
            int col = t0 / p.H, h = t0 - col * p.H;
            bool fresh = true;                                   // first tile of a column: the whole window is new
            for (int t = t0; t < t1; ++t) {
                const int n = col / p.tiles_w, w0 = (col - n * p.tiles_w) * 128;
                const int line0 = p.adv * h + p.row0;
                for (int j = fresh ? 0 : p.J - p.adv; j < p.J; ++j) {
                    const int line = line0 + j;
                    mbar_wait(&empty[slot], ph ^ 1);
                    if (leader) {
                        mbar_arrive_expect_tx(&full[slot], (uint32_t)p.slab_tx);
                        if (p.parity_lines)
                            tma_load_5d(smem + (size_t)slot * p.slab_stride, &tmX, &full[slot], 0, w0 - p.halo, line & 1, line >> 1, n);
                        else
                            tma_load_5d(smem + (size_t)slot * p.slab_stride, &tmX, &full[slot], 0, w0 - p.halo, 0, line, n);
                    }
                    if (++slot == (uint32_t)p.ring) { slot = 0; ph ^= 1; }
                }
                fresh = false;
                if (++h == p.H) { h = 0; ++col; fresh = true; }
            }
        }
    } else if (warp == 1) {
        if (t0 < t1) {
            const bool leader = elect_one_sync();
            const uint32_t idesc = umma_idesc_bf16(128, p.Nmma, 0, 0);
            const uint32_t a_hi = (uint32_t)(umma_smem_desc(0, 16, 8u * p.arow, umma_layout_code(p.arow)) >> 32);
            const uint32_t b_hi = (uint32_t)(umma_smem_desc(0, 16, 8u * p.brow, umma_layout_code(p.brow)) >> 32);
            constexpr uint32_t LBO_LO = (16u >> 4) << 16;
            const uint32_t smem_base = smem_u32(smem);
            const uint32_t w_base = ((smem_base + (uint32_t)p.w_off) >> 4) | LBO_LO;
            const uint32_t w_step = (uint32_t)p.w_slice_stride >> 4;
            const uint32_t a_shift = (uint32_t)p.arow >> 4;
            const uint32_t khalf16 = (uint32_t)p.brow >> 4;            // s2: the b = 1 half of a slab row starts C channels (= one filter row) in
            // everything that does not change from tile to tile is formed once: per entry the filter descriptor and the offset of
            // its activation view inside a slab; per tile only the four slab addresses (ring rotation) and the accumulator base
            uint32_t wb[16], aoff[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                wb[i] = w_base + (uint32_t)i * w_step;
                aoff[i] = (uint32_t)WinTapC<KIND>::shift(i) * a_shift + (uint32_t)WinTapC<KIND>::khalf(i) * khalf16;
            }
            const uint32_t nmma = (uint32_t)p.Nmma, ring = (uint32_t)p.ring, slab_stride16 = (uint32_t)p.slab_stride >> 4;
            const uint32_t slab0 = (smem_base >> 4) | LBO_LO;
            mbar_wait(wbar, 0);
            uint32_t first = 0, wait_slot = 0, wait_ph = 0, tcount = 0;
            int h = t0 % p.H;
            bool fresh = true;
            for (int t = t0; t < t1; ++t, ++tcount) {
                const uint32_t acc = tcount & 1;
                const bool next_fresh = (h + 1 == p.H);
                mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);
                const int n_new = fresh ? p.J : p.adv;
                for (int i = 0; i < n_new; ++i) {                 // the new slabs of this tile have landed?
                    mbar_wait(&full[wait_slot], wait_ph);
                    if (++wait_slot == (uint32_t)p.ring) { wait_slot = 0; wait_ph ^= 1; }
                }
                tc_fence_after();
                const uint32_t d_base = tmem_base + acc * (uint32_t)p.nacc * nmma;
                uint32_t sa[4];                                    // descriptor low words of the window's slabs (J <= 4)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t slot = first + j;
                    if (slot >= ring) slot -= ring;
                    sa[j] = slab0 + slot * slab_stride16;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    using T = WinTapC<KIND>;
                    const uint32_t a_lo = sa[T::slab(i)] + aoff[i];
                    const uint32_t d = d_base + (uint32_t)T::region(i) * nmma;
                    // the first entry of every accumulator region overwrites it: x2 -> the first tap of each phase, s2 -> entry 0
                    const bool first_of_region = KIND == 1 ? (i & 3) == 0 : i == 0;
#pragma unroll
                    for (int j = 0; j < KS; ++j)
                        if (leader) tc_mma_f16_lohi2(d, a_lo + 2 * j, a_hi, wb[i] + 2 * j, b_hi, idesc, (first_of_region && j == 0) ? 0u : 1u);
                }
                if (leader) tc_commit(&tfull[acc]);
                // release the slabs the next tile will not read: `adv` when it continues this column, the whole window otherwise
                const int n_rel = (t + 1 < t1) ? (next_fresh ? p.J : p.adv) : 0;
                for (int i = 0; i < n_rel; ++i) {
                    if (leader) tc_commit(&empty[first]);
                    if (++first == (uint32_t)p.ring) first = 0;
                }
                fresh = next_fresh;
                if (++h == p.H) h = 0;
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;                   // pixel within the tile == w offset
        constexpr int NS = NM > 0 ? NM * 16 : 1;
        float ssum[NS], ssq[NS];                         // per-thread statistic accumulators (NM > 0 and p.stats)
#pragma unroll
        for (int i = 0; i < NS; ++i) ssum[i] = ssq[i] = 0.f;
        uint32_t tcount = 0;
        int col = t0 / p.H, h = t0 - col * p.H;
        for (int t = t0; t < t1; ++t, ++tcount) {
            const int n = col / p.tiles_w, w = (col - n * p.tiles_w) * 128 + row;
            const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * (uint32_t)(p.nacc * p.Nmma);
            for (int r = 0; r < p.nacc; ++r) {
                const size_t pix = ((size_t)n * p.Ho + (size_t)(h * p.osy + p.oy[r])) * p.Wo + (size_t)(w * p.osx + p.ox[r]);
                __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.Co_pad;
                if (NM > 0) {
#pragma unroll
                    for (int c = 0; c < (NM > 0 ? NM : 1); ++c) {
                        uint32_t v[16];
                        tmem_ld16(taddr + (uint32_t)(r * p.Nmma + c * 16), v);
                        tmem_ld_wait();
                        uint32_t wv[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            wv[i] = pack_bf16(__uint_as_float(v[2 * i]) + bias_s[c * 16 + 2 * i], __uint_as_float(v[2 * i + 1]) + bias_s[c * 16 + 2 * i + 1]);
                        st_global_256(orow + c * 16, wv);
                        if (p.stats) {                            // statistics of the values as stored (bf16-rounded)
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float lo = bf16_lo(wv[i]), hi = bf16_hi(wv[i]);
                                ssum[(c * 16 + 2 * i) % NS] += lo;
                                ssum[(c * 16 + 2 * i + 1) % NS] += hi;
                                ssq[(c * 16 + 2 * i) % NS] = fmaf(lo, lo, ssq[(c * 16 + 2 * i) % NS]);
                                ssq[(c * 16 + 2 * i + 1) % NS] = fmaf(hi, hi, ssq[(c * 16 + 2 * i + 1) % NS]);
                            }
                        }
                    }
                } else {
                    for (int cl = 0; cl < p.Nmma; cl += 16) {
                        uint32_t v[16];
                        tmem_ld16(taddr + (uint32_t)(r * p.Nmma + cl), v);
                        tmem_ld_wait();
                        uint32_t wv[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            wv[i] = pack_bf16(__uint_as_float(v[2 * i]) + bias_s[cl + 2 * i], __uint_as_float(v[2 * i + 1]) + bias_s[cl + 2 * i + 1]);
                        st_global_256(orow + cl, wv);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++h == p.H) { h = 0; ++col; }
        }
        if (NM > 0 && p.stats) {                         // lanes in order -> warps in order -> CTAs in order (fv_reduce.cuh): reproducible
            const int tid = threadIdx.x - 64, n2 = 2 * p.Nmma;
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                float a = ssum[i], b = ssq[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                }
                if (lane == 0) {
                    stat_w[q * n2 + i] = a;
                    stat_w[q * n2 + p.Nmma + i] = b;
                }
            }
            named_bar_sync(1, 128);
            for (int c = tid; c < n2; c += 128) stat_blk[c] = ((stat_w[c] + stat_w[n2 + c]) + stat_w[2 * n2 + c]) + stat_w[3 * n2 + c];
            named_bar_sync(1, 128);
            if (det_reduce<float>(p.red_ws, n2, gridDim.x, blockIdx.x, stat_blk, stat_tot, tid, 128, NamedSync{1, 128}, red_flag))
                for (int c = tid; c < n2; c += 128) p.stats[c < p.Nmma ? c : p.stats_c + c - p.Nmma] = stat_tot[c];
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

static bool win_enabled() {
    const char* env = getenv("FV_CONV_WIN");
    return !(env && atoi(env) == 0);
}

// 1 when conv_win_try takes this x2 (kind 1) / s2 (kind 2) call (same conditions, no launch); H, W = the tiling grid
int conv_win_eligible(int kind, int out_mode, int W, int Ci, int Co_pad) {
    if (!win_enabled() || out_mode != FV_OUT_NHWC_BF16 || W % 128) return 0;
    if (kind == 1) return (Ci == 16 || Ci == 32 || Ci == 64) && Co_pad <= 64;
    if (kind == 2) return (Ci == 16 || Ci == 32) && Co_pad <= 256 && (size_t)16 * Co_pad * 2 * Ci + 8 * 17 * 1024 <= 200 * 1024;
    return 0;
}

// FV_OK after launching, -1 when not eligible (the caller falls through to conv_igemm_kernel), > 0 on error.
// kind 1: x [N,H,W,Ci] coarse, w = wx2 [4][Co_pad][4*Ci], y [N,2H,2W,Co_pad];  kind 2: x [N,2H,2W,Ci] fine, w = [Co_pad][16*Ci], y [N,H,W,Co_pad]
int conv_win_try(int kind, const void* x, const void* w, const float* bias, void* y, int out_mode, int N, int H, int W, int Ci, int Co, int Co_pad,
                 float* stats, void* red_ws, cudaStream_t stream) {
    if (!conv_win_eligible(kind, out_mode, W, Ci, Co_pad)) return -1;
    WinParams p{};
    p.N = N; p.H = H; p.W = W; p.Co = Co; p.Co_pad = Co_pad;
    p.tiles_w = W / 128;
    p.num_tiles = N * p.tiles_w * H;
    p.halo = 1;
    const int slab_px = 128 + 2;
    p.Nmma = Co_pad;
    if (kind == 1) {
        p.Ho = 2 * H; p.Wo = 2 * W; p.osy = 2; p.osx = 2;
        p.nacc = 4;
        p.J = 3; p.adv = 1; p.row0 = -1; p.parity_lines = 0;
        p.arow = Ci * 2; p.brow = Ci * 2;
        p.w_slices = 16;
        p.w_slice_tx = Co_pad * p.brow;
        p.ntaps = 16;
        for (int ph = 0; ph < 4; ++ph) {
            const int a = ph >> 1, b = ph & 1;
            p.oy[ph] = (signed char)a; p.ox[ph] = (signed char)b;
            for (int t = 0; t < 4; ++t) {
                const int u = t >> 1, v = t & 1, i = ph * 4 + t;
                p.tap[i] = WinTap{(unsigned char)(u + a), (unsigned char)(v + b), 0, (unsigned char)(Ci / 16), (unsigned short)i, (unsigned short)(ph * Co_pad)};
                p.wk[i] = (short)(t * Ci);
                p.wr[i] = (short)(ph * Co_pad);
            }
        }
    } else {
        p.Ho = H; p.Wo = W; p.osy = 1; p.osx = 1;
        p.nacc = 1;
        p.J = 4; p.adv = 2; p.row0 = -1; p.parity_lines = 1;
        p.arow = 2 * Ci * 2; p.brow = Ci * 2;
        p.w_slices = 16;
        p.w_slice_tx = Co_pad * p.brow;
        p.ntaps = 16;
        for (int t = 0; t < 16; ++t) {
            const int r4 = t >> 2, fc = (t & 3) - 1;
            const int dw = fc < 0 ? -1 : fc / 2, b = fc - 2 * dw;
            p.tap[t] = WinTap{(unsigned char)r4, (unsigned char)(dw + 1), (unsigned char)(b * Ci * 2 / 16), (unsigned char)(Ci / 16), (unsigned short)t, 0};
            p.wk[t] = (short)(t * Ci);
            p.wr[t] = 0;
        }
    }
    for (int i = 0; i < 16; ++i) {        // the kernel's compile-time table must be the one described here
        const WinTap& t = p.tap[i];
        const bool ok = kind == 1 ? (t.slab == WinTapC<1>::slab(i) && t.shift == WinTapC<1>::shift(i) && t.koff16 == 0 && t.dcol == WinTapC<1>::region(i) * Co_pad && t.wslice == i)
                                  : (t.slab == WinTapC<2>::slab(i) && t.shift == WinTapC<2>::shift(i) && t.koff16 == WinTapC<2>::khalf(i) * (Ci * 2 / 16) && t.dcol == 0 && t.wslice == i);
        if (!ok) return fail(FV_ERR_INTERNAL, "conv_win: tap table mismatch at entry %d", i);
    }
    p.slab_tx = slab_px * p.arow;
    p.slab_stride = (p.slab_tx + 1023) & ~1023;
    p.ring = p.J + 2 * p.adv;
    p.w_slice_stride = (p.w_slice_tx + 1023) & ~1023;
    p.w_off = p.ring * p.slab_stride;
    p.bar_off = p.w_off + p.w_slices * p.w_slice_stride;
    const size_t smem = (size_t)p.bar_off + (2 * p.ring + 8) * 8 + 16 + (size_t)p.Nmma * 52 + 1024 + 64;
    if (smem > 225 * 1024) return -1;
    int cols = 32;
    while (cols < 2 * p.nacc * p.Nmma) cols <<= 1;
    if (cols > 512) return -1;
    p.tmem_cols = cols;
    p.bias = bias;
    p.out = y;
    p.stats = stats;
    p.stats_c = Co_pad;
    p.red_ws = red_ws;
    const int sms = num_sms();
    p.tiles_per_cta = (p.num_tiles + sms - 1) / sms;
    const int grid = (p.num_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;

    CUtensorMap tmX, tmW;
    {
        const uint64_t f = kind == 2 ? 2 : 1;
        uint64_t dims[5] = {(uint64_t)Ci * f, (uint64_t)W, f, (uint64_t)H, (uint64_t)N};
        uint64_t str[4] = {(uint64_t)Ci * f * 2, (uint64_t)W * f * Ci * 2, (uint64_t)W * f * Ci * 2 * f, (uint64_t)H * f * W * f * Ci * 2};
        uint32_t box[5] = {(uint32_t)(Ci * f), (uint32_t)slab_px, 1, 1, 1};
        if (int e = encode_tmap_bf16(&tmX, x, 5, dims, str, box, p.arow)) return e;
    }
    {
        const uint64_t K = (uint64_t)(kind == 1 ? 4 : 16) * Ci, rows = (uint64_t)(kind == 1 ? 4 : 1) * Co_pad;
        uint64_t dims[2] = {K, rows};
        uint64_t str[1] = {K * 2};
        uint32_t box[2] = {(uint32_t)Ci, (uint32_t)Co_pad};
        if (int e = encode_tmap_bf16(&tmW, w, 2, dims, str, box, p.brow)) return e;
    }
#define FV_WIN_LAUNCH3(NM_, KIND_, KS_)                                                                                              \
    do {                                                                                                                             \
        static bool attr_set = false;                                                                                                \
        if (!attr_set) {                                                                                                             \
            FV_CUDA(cudaFuncSetAttribute(conv_win_kernel<NM_, KIND_, KS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
            attr_set = true;                                                                                                         \
        }                                                                                                                            \
        conv_win_kernel<NM_, KIND_, KS_><<<grid, kWinThreads, smem, stream>>>(tmX, tmW, p);                                          \
    } while (0)
#define FV_WIN_LAUNCH(NM_, KIND_)                                                                                                    \
    do {                                                                                                                             \
        if (Ci == 64) FV_WIN_LAUNCH3(NM_, KIND_, 4);                                                                                 \
        else if (Ci == 32) FV_WIN_LAUNCH3(NM_, KIND_, 2);                                                                            \
        else FV_WIN_LAUNCH3(NM_, KIND_, 1);                                                                                          \
    } while (0)
    if (kind == 1 && p.Nmma == 16) FV_WIN_LAUNCH(1, 1);
    else if (kind == 1 && p.Nmma == 32) FV_WIN_LAUNCH(2, 1);
    else if (kind == 1 && p.Nmma == 64) FV_WIN_LAUNCH(4, 1);
    else {
        if (stats || kind != 2) return fail(FV_ERR_INTERNAL, "conv_win: statistics are fused for the x2 geometry only");
        FV_WIN_LAUNCH(0, 2);
    }
#undef FV_WIN_LAUNCH
#undef FV_WIN_LAUNCH3
    FV_LAUNCH_CHECK("conv_win_kernel");
    return FV_OK;
}

}  // namespace fv
