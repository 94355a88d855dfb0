// Sliding-window ("ring") schedule of the implicit-GEMM convolution for the full-resolution thin layers
// (Ci <= 64, filter resident in shared memory): the layers where the plain schedule is bound by L2->SMEM traffic
// rather than by the tensor core.
//
// A tile is one image-row segment of 128 pixels.  Its R x S taps read R input-row slabs of (128 + S - 1) pixels; the
// S taps of a row are row-shifted UMMA descriptor views of the same slab (see fv_conv.cu).  Each CTA walks a
// contiguous run of vertically adjacent tiles, so consecutive tiles share R - 1 of their R slabs: the slabs live in
// a shared-memory ring and only ONE new slab (plus nothing for the filter, which is loaded once per CTA) is fetched
// per tile -- the SMEM fill traffic equals the HBM traffic of the layer.  Out-of-image rows / columns are zero-filled
// by the TMA unit as before.
//
// Epilogue (NHWC bf16, Co_pad <= 64): TMEM -> registers -> (+bias) -> bf16 -> swizzled shared-memory staging tile ->
// one TMA tensor store per tile (full 128-byte lines instead of 32 scattered 16-byte stores per instruction).
// NCHW fp32 output (the 3-channel out_conv) is stored directly: consecutive lanes are consecutive pixels.
#include <cstdio>
#include <cstdlib>

#include "../../include/facevae_b200.h"
#include "fv_host.h"
#include "fv_ptx.cuh"
#include "fv_reduce.cuh"

namespace fv {

struct RingParams {
    int N, H, W, Ci, Co, Co_pad, R, S, pad;
    int tiles_w, num_tiles, tiles_per_cta;
    int ring, slab_stride, slab_tx;       // slots, placement stride and TMA bytes of one slab
    int w_off, w_slice_stride, w_tx;      // filter region offset, per-tap slice stride, total TMA bytes
    int stage_off, stage_stride;          // epilogue staging buffers (2), 0 stride when unused
    int bar_off;
    int out_mode, tmem_cols;
    const float* bias;
    void* out;
    float* stats;                         // optional [2][stats_c]: per-channel sum and sum of squares of the stored output
    int stats_c;
    void* red_ws;                         // fv_reduce.cuh workspace (with stats)
    int dual;                             // two issuer warps on alternating tiles (3x3 only)
    long long* trace;
};

static constexpr int kRingThreads = 224;   // producer, issuer A, 4 epilogue warps, issuer B (dual mode)

template <int KB, int S_>
__global__ void __launch_bounds__(kRingThreads, 1)
conv_ring_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmY, const RingParams p) {
    constexpr int ROW = KB * 2;
    constexpr int KSUB = KB / 16;
    constexpr uint32_t LAYOUT = ROW == 128 ? 2u : (ROW == 64 ? 4u : 6u);
    constexpr uint32_t SBO = 8u * ROW;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // align up with arithmetic on the array itself so the compiler keeps the shared address space (LDS/STS, not generic)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.bar_off);      // [ring]
    uint64_t* empty = full + p.ring;                                      // [ring]
    uint64_t* wbar = empty + p.ring;
    uint64_t* tfull = wbar + 1;                                           // [2]
    uint64_t* tempty = tfull + 2;                                         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    int* red_flag = reinterpret_cast<int*>(tmem_slot + 2);
    float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);              // [Co_pad]
    float* stat_w = bias_s + p.Co_pad;                                    // [4 epilogue warps][2][Co_pad] (when p.stats)
    float* stat_blk = stat_w + 8 * p.Co_pad;                              // [2][Co_pad] CTA totals
    float* stat_tot = stat_blk + 2 * p.Co_pad;                            // [2][Co_pad] grid totals (last CTA)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef FV_TRACE
    long long* fv_trace = p.trace;
#endif
    const int t0 = blockIdx.x * p.tiles_per_cta;
    const int t1 = min(t0 + p.tiles_per_cta, p.num_tiles);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        if (p.stage_stride) tma_prefetch_desc(&tmY);
        for (int i = 0; i < p.ring; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], p.dual ? 2 : 1);      // dual mode: one arrival from each issuer (see below)
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
    for (int c = threadIdx.x; c < p.Co_pad; c += blockDim.x) bias_s[c] = (p.bias && c < p.Co) ? p.bias[c] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile t -> column (n, w segment) = t / H, image row h = t % H: consecutive tiles are vertically adjacent
    if (warp == 0) {
        if (t0 < t1) {
            const bool leader = elect_one_sync();
            // filter: R*S slices of [Co_pad x KB], resident for the whole kernel
            if (leader) mbar_arrive_expect_tx(wbar, (uint32_t)p.w_tx);
            for (int tap = 0; tap < p.R * p.S; ++tap)
                if (leader) tma_load_2d(smem + p.w_off + tap * p.w_slice_stride, &tmW, wbar, tap * p.Ci, 0);
            uint32_t slot = 0, ph = 0;
            int col = t0 / p.H, h = t0 - col * p.H;
            bool fresh = true;                                   // first tile of a column: all R slabs are new
            for (int t = t0; t < t1; ++t) {
                const int n = col / p.tiles_w, w0 = (col - n * p.tiles_w) * 128;
                for (int j = fresh ? 0 : p.R - 1; j < p.R; ++j) {
                    { FV_T0(tw); mbar_wait(&empty[slot], ph ^ 1); FV_TACC(0, tw); }
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
    } else if ((warp == 1 || warp == 6) && p.dual) {
        // Two issuer warps on alternating tiles (3x3 filters).  A tile's fixed costs in the issuing thread -- observing the accumulator
        // and slab barriers, two commits: ~500-1200 cycles against 860-1600 cycles of MMA issue -- are serial with its MMAs in one
        // thread; with two threads one tile's bookkeeping overlaps the other tile's MMAs (the accumulators are already double-buffered:
        // issuer w owns accumulator w).  Both warps run the same slot / phase bookkeeping over ALL tiles and act on their own.
        // Slab row r is read by tiles r-1, r, r+1, i.e. by both issuers: its `empty` barrier takes two arrivals, one from each issuer
        // after that issuer's LAST tile reading it -- after its own tile t an issuer commits rows t-1 and t (slots first, first+1);
        // the first tile of a run commits row t-1 twice (no predecessor tile), the last tile of a column rows t and t+1 twice.
        if (t0 < t1) {
            const uint32_t me = warp == 6 ? 1u : 0u;
            const bool leader = elect_one_sync();
            const uint32_t idesc = umma_idesc_bf16(128, p.Co_pad, 0, 0);
            const uint32_t desc_hi = (uint32_t)(umma_smem_desc(0, 16, SBO, LAYOUT) >> 32);
            constexpr uint32_t LBO_LO = (16u >> 4) << 16;
            const uint32_t smem_base = smem_u32(smem);
            const uint32_t w_base = ((smem_base + (uint32_t)p.w_off) >> 4) | LBO_LO;
            const uint32_t w_step = (uint32_t)p.w_slice_stride >> 4;
            mbar_wait(wbar, 0);
            // `first` / `first_ph`: ring slot and fill parity of the tile's top slab (row h - 1); the tile reads the S_ consecutive slabs
            uint32_t first = 0, first_ph = 0, tcount = 0;
            int h = t0 % p.H;
            bool fresh = true;
            auto nxt = [&](uint32_t s) { return s + 1 == (uint32_t)p.ring ? 0u : s + 1; };
            for (int t = t0; t < t1; ++t, ++tcount) {
                const uint32_t acc = tcount & 1;
                const bool mine = acc == me;
                const bool last_in_col = (h + 1 == p.H);
                if (mine) {
                    mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);
                    // every slab this tile reads is observed by THIS thread (two of them were new for tiles of the other issuer)
                    {
                        uint32_t ws_ = first, wp_ = first_ph;
#pragma unroll
                        for (int i = 0; i < S_; ++i) {
                            mbar_wait(&full[ws_], wp_);
                            if (++ws_ == (uint32_t)p.ring) { ws_ = 0; wp_ ^= 1; }
                        }
                    }
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.Co_pad;
                    uint32_t accumulate = 0, slot = first, wtap = w_base;
#pragma unroll
                    for (int r = 0; r < S_; ++r) {
                        const uint32_t a_row = ((smem_base + slot * (uint32_t)p.slab_stride) >> 4) | LBO_LO;
#pragma unroll
                        for (int s = 0; s < S_; ++s) {
#pragma unroll
                            for (int j = 0; j < KSUB; ++j) {
                                if (leader)
                                    tc_mma_f16_lohi(d_tmem, a_row + (uint32_t)(s * (ROW >> 4) + 2 * j), wtap + (uint32_t)(2 * j), desc_hi, idesc,
                                                    accumulate);
                                accumulate = 1;
                            }
                            wtap += w_step;
                        }
                        slot = nxt(slot);
                    }
                    if (leader) {
                        tc_commit(&tfull[acc]);
                        if (t + 1 < t1) {          // (the CTA's last tile releases nothing: no load is waiting)
                            const uint32_t s0 = first, s1 = nxt(first), s2 = nxt(s1);
                            tc_commit(&empty[s0]);
                            if (fresh) tc_commit(&empty[s0]);
                            tc_commit(&empty[s1]);
                            if (last_in_col) {
                                tc_commit(&empty[s1]);
                                tc_commit(&empty[s2]);
                                tc_commit(&empty[s2]);
                            }
                        }
                    }
                }
                const int n_rel = last_in_col ? S_ : 1;
                for (int i = 0; i < n_rel; ++i)
                    if (++first == (uint32_t)p.ring) { first = 0; first_ph ^= 1; }
                fresh = last_in_col;
                if (++h == p.H) h = 0;
            }
        }
    } else if (warp == 1) {
        if (t0 < t1) {
            const bool leader = elect_one_sync();
            const uint32_t idesc = umma_idesc_bf16(128, p.Co_pad, 0, 0);
            const uint32_t desc_hi = (uint32_t)(umma_smem_desc(0, 16, SBO, LAYOUT) >> 32);   // low word = LBO | start address
            constexpr uint32_t LBO_LO = (16u >> 4) << 16;
            const uint32_t smem_base = smem_u32(smem);
            const uint32_t w_base = ((smem_base + (uint32_t)p.w_off) >> 4) | LBO_LO;
            const uint32_t w_step = (uint32_t)p.w_slice_stride >> 4;
            mbar_wait(wbar, 0);
            // window = R consecutive ring slots starting at `first`; `wait_slot/wait_ph` track the next slab to arrive
            uint32_t first = 0, wait_slot = 0, wait_ph = 0, tcount = 0;
            int h = t0 % p.H;
            bool fresh = true;
            FV_T0(t_all);
            // barrier waits of tile t + 1 are taken in the MIDDLE of tile t's MMAs: a completed mbarrier still costs
            // ~150-250 cycles to observe, and the tensor pipe would otherwise drain while the issuer polls
            auto wait_tile = [&](uint32_t tc, bool is_fresh) {
                { FV_T0(tw); mbar_wait(&tempty[tc & 1], ((tc >> 1) & 1) ^ 1); FV_TACC(2, tw); }
                const int n_new = is_fresh ? p.R : 1;
                FV_T0(tw2);
                for (int i = 0; i < n_new; ++i) {                 // the new slabs of that tile have landed?
                    mbar_wait(&full[wait_slot], wait_ph);
                    if (++wait_slot == (uint32_t)p.ring) { wait_slot = 0; wait_ph ^= 1; }
                }
                FV_TACC(3, tw2);
            };
            bool waited = false;
            for (int t = t0; t < t1; ++t, ++tcount) {
                const uint32_t acc = tcount & 1;
                const bool next_fresh = (h + 1 == p.H);
                // a column change needs R fresh slabs, i.e. slots this tile still occupies: that wait cannot be taken early
                if (!waited) wait_tile(tcount, fresh);
                waited = false;
                tc_fence_after();
                // non-blocking probes of tile t + 1's barriers, issued before this tile's MMAs and consumed in their middle:
                // the mbarrier round trip overlaps the (blocking) MMA issue instead of stalling between two MMAs
                uint32_t probe_te = 0, probe_full = 0;
                if (t + 1 < t1 && !next_fresh) {
                    probe_te = mbar_test_wait(&tempty[(tcount + 1) & 1], (((tcount + 1) >> 1) & 1) ^ 1);
                    probe_full = mbar_test_wait(&full[wait_slot], wait_ph);
                }
                FV_T0(t_issue);
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.Co_pad;
                uint32_t accumulate = 0, slot = first, wtap = w_base;
#pragma unroll
                for (int r = 0; r < S_; ++r) {
                    const uint32_t a_row = ((smem_base + slot * (uint32_t)p.slab_stride) >> 4) | LBO_LO;
#pragma unroll
                    for (int s = 0; s < S_; ++s) {                                    // tap s == slab shifted by s pixel rows
#pragma unroll
                        for (int j = 0; j < KSUB; ++j) {
                            if (leader)
                                tc_mma_f16_lohi(d_tmem, a_row + (uint32_t)(s * (ROW >> 4) + 2 * j), wtap + (uint32_t)(2 * j), desc_hi, idesc,
                                                accumulate);
                            accumulate = 1;
                        }
                        wtap += w_step;
                    }
                    if (++slot == (uint32_t)p.ring) slot = 0;
                    if (r == (S_ - 1) / 2 && t + 1 < t1 && !next_fresh) {
                        if (!probe_te) mbar_wait(&tempty[(tcount + 1) & 1], (((tcount + 1) >> 1) & 1) ^ 1);
                        if (!probe_full) mbar_wait(&full[wait_slot], wait_ph);
                        if (++wait_slot == (uint32_t)p.ring) { wait_slot = 0; wait_ph ^= 1; }
                        waited = true;
                    }
                }
                FV_TACC(4, t_issue);
                FV_T0(t_commit);
                if (leader) tc_commit(&tfull[acc]);
                // release the slabs the next tile will not read: one when it continues this column, all R otherwise
                const int n_rel = (t + 1 < t1) ? (next_fresh ? p.R : 1) : 0;
                for (int i = 0; i < n_rel; ++i) {
                    if (leader) tc_commit(&empty[first]);
                    if (++first == (uint32_t)p.ring) first = 0;
                }
                fresh = next_fresh;
                if (++h == p.H) h = 0;
                FV_TACC(1, t_commit);
            }
            FV_TACC(5, t_all);
        }
    } else if (warp >= 2 && warp <= 5) {
        const int q = warp & 3;
        const int row = q * 32 + lane;                   // pixel within the tile == w offset
        constexpr int EPI_BAR = 1;
        const bool use_tma_store = p.stage_stride != 0;
        const int out_row = p.Co_pad * 2;                // bytes per pixel of the NHWC bf16 output tile
        // 16-byte chunk swizzle of the staging tile (must match the store tensor map: 128B / 64B / 32B swizzle)
        const int sw_mask = out_row == 128 ? 7 : (out_row == 64 ? 3 : 1);
        const int sw_shift = out_row == 128 ? 0 : (out_row == 64 ? 1 : 2);
        const int sw = (row >> sw_shift) & sw_mask;
        uint32_t tcount = 0;
        int col = t0 / p.H, h = t0 - col * p.H;
        // batch-norm statistics of the stored (bf16-rounded) output, fused: per-lane partial column sums live in
        // registers across the CTA's tiles (lane l <-> channel 16 c + ((l >> 1) & 15) of chunk c), one atomic per lane pair
        // and chunk at the end -- replaces a full re-read of the output by fv_bn_stats
        float st_s[4] = {0.f, 0.f, 0.f, 0.f}, st_q[4] = {0.f, 0.f, 0.f, 0.f};
        for (int t = t0; t < t1; ++t, ++tcount) {
            const int n = col / p.tiles_w, w0 = (col - n * p.tiles_w) * 128;
            const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
            uint8_t* stage = smem + p.stage_off + (size_t)acc * p.stage_stride;
            if (use_tma_store) {
                // the TMA store that last read this staging buffer (two tiles ago) must have drained it
                if (warp == 2 && lane == 0) tma_store_wait_read<1>();
                named_bar_sync(EPI_BAR, 128);
            }
            { FV_T0(tw); mbar_wait(&tfull[acc], aph); if (warp == 2) FV_TACC(6, tw); }
            tc_fence_after();
            FV_T0(t_epi);
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * (uint32_t)p.Co_pad;
            if (use_tma_store) {
                // fast path (NHWC bf16, Co_pad <= 64): pull the whole accumulator row into registers with back-to-back
                // TMEM loads and ONE wait, hand the accumulator back to the issuer immediately, then convert and stage
                const int nc = p.Co_pad >> 4;
                uint32_t v[4][16];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < nc) tmem_ld16(taddr + c * 16, v[c]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                uint8_t* srow = stage + row * out_row;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < nc) {
                        float f[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[c][i]) + bias_s[c * 16 + i];
                        uint32_t wv[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) wv[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
                        *reinterpret_cast<uint4*>(srow + (((2 * c) ^ sw) << 4)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                        *reinterpret_cast<uint4*>(srow + (((2 * c + 1) ^ sw) << 4)) = make_uint4(wv[4], wv[5], wv[6], wv[7]);
                        if (p.stats) {
                            float fr[16], fq[16];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                fr[2 * i] = bf16_lo(wv[i]);
                                fr[2 * i + 1] = bf16_hi(wv[i]);
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) fq[i] = fr[i] * fr[i];
                            st_s[c] += warp_colsum16(fr, lane);
                            st_q[c] += warp_colsum16(fq, lane);
                        }
                    }
                }
                fence_proxy_async();                       // make the staging writes visible to the TMA unit
                named_bar_sync(EPI_BAR, 128);
                if (warp == 2 && lane == 0) {
                    tma_store_4d(&tmY, stage, 0, w0, h, n);
                    tma_store_commit();
                }
            } else {
                for (int c0 = 0; c0 < p.Co_pad; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c0, v);
                    tmem_ld_wait();
                    float f[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + bias_s[c0 + i];
                    if (p.out_mode == FV_OUT_NCHW_F32) {
                        float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < p.Co) o[(((size_t)n * p.Co + c0 + i) * p.H + h) * p.W + w0 + row] = f[i];
                    } else {
                        const size_t pix = ((size_t)n * p.H + h) * p.W + w0 + row;
                        if (p.out_mode == FV_OUT_NHWC_BF16) {
                            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.Co_pad + c0);
                            o[0] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                            o[1] = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
                        } else {
                            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.Co_pad + c0);
#pragma unroll
                            for (int i = 0; i < 4; ++i) o[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);     // accumulator drained: the issuer may start tile t + 2
            }
            if (warp == 2) FV_TACC(7, t_epi);
            if (++h == p.H) { h = 0; ++col; }
        }
        if (use_tma_store && warp == 2 && lane == 0) tma_store_wait_all<0>();
        if (p.stats && use_tma_store) {            // lanes -> warp slots -> warps in order -> CTAs in order (fv_reduce.cuh): reproducible
            const int tid = threadIdx.x - 64, n2 = 2 * p.Co_pad;
            if (!(lane & 1)) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < (p.Co_pad >> 4)) {
                        stat_w[q * n2 + c * 16 + ((lane >> 1) & 15)] = st_s[c];
                        stat_w[q * n2 + p.Co_pad + c * 16 + ((lane >> 1) & 15)] = st_q[c];
                    }
            }
            named_bar_sync(EPI_BAR, 128);
            for (int c = tid; c < n2; c += 128) stat_blk[c] = ((stat_w[c] + stat_w[n2 + c]) + stat_w[2 * n2 + c]) + stat_w[3 * n2 + c];
            named_bar_sync(EPI_BAR, 128);
            if (det_reduce<float>(p.red_ws, n2, gridDim.x, blockIdx.x, stat_blk, stat_tot, tid, 128, NamedSync{EPI_BAR, 128}, red_flag))
                for (int c = tid; c < n2; c += 128) p.stats[c < p.Co_pad ? c : p.stats_c + c - p.Co_pad] = stat_tot[c];
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

template <int KB, int S_>
static int launch_ring(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmY, const RingParams& p, size_t smem,
                       int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        FV_CUDA(cudaFuncSetAttribute(conv_ring_kernel<KB, S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    conv_ring_kernel<KB, S_><<<grid, kRingThreads, smem, stream>>>(tmX, tmW, tmY, p);
    FV_LAUNCH_CHECK("conv_ring_kernel");
    return FV_OK;
}

// Returns FV_OK after launching, or -1 when the configuration is not eligible (caller falls through to the generic kernel).
// 1 when conv2d_ring_try takes this shape (same conditions, no launch)
// ring slots (R + 3 preferred, down to R + 1: one slab of prefetch) with which the kernel's shared memory fits; 0 = does not fit
static size_t ring_smem(int ring, int out_mode, int Ci, int Co_pad, int R, int S) {
    const int row_bytes = Ci * 2;
    const int slab_stride = ((128 + S - 1) * row_bytes + 1023) & ~1023;
    int off = ring * slab_stride + R * S * ((Co_pad * row_bytes + 1023) & ~1023);
    if (out_mode == FV_OUT_NHWC_BF16 && Co_pad <= 64) off += 2 * ((128 * Co_pad * 2 + 1023) & ~1023);
    return (size_t)off + (2 * ring + 8) * 8 + 16 + (size_t)Co_pad * 52 + 1024 + 64;
}
static int ring_slots(int out_mode, int Ci, int Co_pad, int R, int S) {
    for (int ring = R + 3; ring >= R + 1; --ring)
        if (ring_smem(ring, out_mode, Ci, Co_pad, R, S) <= 225 * 1024) return ring;
    return 0;
}

static bool ring_chunked_ok(int out_mode, int Co_pad) {
    // measured on enc.2 (64 -> 128 at 128 x 128, batch 32): 2 x 44.5 us against 86 us for the generic schedule -- no gain, so this is
    // opt-in (FV_CONV_RING_CHUNK=1); the single pass with a 4-slot ring (ring_slots) is what the layer runs
    const char* env = getenv("FV_CONV_RING_CHUNK");
    return out_mode == FV_OUT_NHWC_BF16 && Co_pad > 64 && Co_pad % 64 == 0 && env && atoi(env) == 1;
}

int conv2d_ring_eligible(int out_mode, int H, int W, int Ci, int Co_pad, int R, int S, bool residual) {
    if ((S != 3 && S != 5 && S != 7) || R != S || W % 128 || Ci > 64 || residual) return 0;
    const char* env = getenv("FV_CONV_RING");
    if (env && atoi(env) == 0) return 0;
    if (Co_pad > 256) return 0;
    if (ring_slots(out_mode, Ci, Co_pad, R, S)) return 1;
    if (ring_chunked_ok(out_mode, Co_pad)) return ring_slots(out_mode, Ci, 64, R, S) ? 1 : 0;
    (void)H;
    return ring_slots(out_mode, Ci, Co_pad, R, S) ? 1 : 0;
}

// y_cs: channel stride of the output tensor in elements (== Co_pad unless this launch writes a 64-channel chunk of a wider tensor)
static int ring_launch(const void* x, const void* w, const float* bias, const void* residual, void* y, int y_cs, int out_mode, int N, int H,
                       int W, int Ci, int Co, int Co_pad, int R, int S, int pad, float* stats, int stats_c, void* red_ws, cudaStream_t stream) {
    if ((S != 3 && S != 5 && S != 7) || W % 128 || Ci > 64 || residual) return -1;
    if (stats && !(out_mode == FV_OUT_NHWC_BF16 && Co_pad <= 64)) return -1;      // fused statistics: staged-store epilogue only
    const char* env = getenv("FV_CONV_RING");
    if (env && atoi(env) == 0) return -1;
    const int KB = Ci, row_bytes = KB * 2;
    RingParams p{};
    p.N = N; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.Co_pad = Co_pad; p.R = R; p.S = S; p.pad = pad;
    p.tiles_w = W / 128;
    p.num_tiles = N * p.tiles_w * H;
    const int a_rows = 128 + S - 1;
    p.slab_tx = a_rows * row_bytes;
    p.slab_stride = (p.slab_tx + 1023) & ~1023;
    p.ring = ring_slots(out_mode, Ci, Co_pad, R, S);
    if (!p.ring || Co_pad > 256) return -1;
    p.w_slice_stride = (Co_pad * row_bytes + 1023) & ~1023;
    p.w_tx = R * S * Co_pad * row_bytes;
    p.w_off = p.ring * p.slab_stride;
    int off = p.w_off + R * S * p.w_slice_stride;
    const bool tma_store = (out_mode == FV_OUT_NHWC_BF16 && Co_pad <= 64);
    p.stage_off = off;
    p.stage_stride = tma_store ? ((128 * Co_pad * 2 + 1023) & ~1023) : 0;
    off += 2 * p.stage_stride;
    p.bar_off = off;
    const size_t smem = (size_t)off + (2 * p.ring + 8) * 8 + 16 + (size_t)Co_pad * 52 + 1024 + 64;
    if (smem > 225 * 1024) return -1;
    int cols = 32;
    while (cols < 2 * Co_pad) cols <<= 1;
    p.tmem_cols = cols;
    p.out_mode = out_mode;
    p.bias = bias;
    p.out = y;
    p.stats = stats;
    p.stats_c = stats_c;
    p.red_ws = red_ws;
    {
        const char* denv = getenv("FV_RING_DUAL");
        p.dual = (S == 3 && !(denv && atoi(denv) == 0)) ? 1 : 0;
    }
    p.trace = trace_ptr();
    const int sms = num_sms();
    p.tiles_per_cta = (p.num_tiles + sms - 1) / sms;
    const int grid = (p.num_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;

    CUtensorMap tmX, tmW, tmY;
    {
        uint64_t dims[4] = {(uint64_t)Ci, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)Ci * 2, (uint64_t)W * Ci * 2, (uint64_t)H * W * Ci * 2};
        uint32_t box[4] = {(uint32_t)KB, (uint32_t)a_rows, 1, 1};
        if (int e = encode_tmap_bf16(&tmX, x, 4, dims, str, box, row_bytes)) return e;
    }
    {
        uint64_t dims[2] = {(uint64_t)R * S * Ci, (uint64_t)Co_pad};
        uint64_t str[1] = {(uint64_t)R * S * Ci * 2};
        uint32_t box[2] = {(uint32_t)KB, (uint32_t)Co_pad};
        if (int e = encode_tmap_bf16(&tmW, w, 2, dims, str, box, row_bytes)) return e;
    }
    if (tma_store) {
        uint64_t dims[4] = {(uint64_t)Co_pad, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)y_cs * 2, (uint64_t)W * y_cs * 2, (uint64_t)H * W * y_cs * 2};
        uint32_t box[4] = {(uint32_t)Co_pad, 128, 1, 1};
        if (int e = encode_tmap_bf16(&tmY, y, 4, dims, str, box, Co_pad * 2)) return e;
    } else {
        tmY = tmX;
    }
#define FV_RING(KB_) \
    (S == 3 ? launch_ring<KB_, 3>(tmX, tmW, tmY, p, smem, grid, stream) \
            : (S == 5 ? launch_ring<KB_, 5>(tmX, tmW, tmY, p, smem, grid, stream) : launch_ring<KB_, 7>(tmX, tmW, tmY, p, smem, grid, stream)))
    if (KB == 64) return FV_RING(64);
    if (KB == 32) return FV_RING(32);
    return FV_RING(16);
#undef FV_RING
}

// Layers with Ci <= 64 but MORE than 64 output channels (enc.2: 64 -> 128 at 128 x 128): the resident filter of the whole layer
// (147 KB) leaves no room for the slab ring, and under the generic schedule the filter is re-streamed for every 128-pixel tile
// (ncu: 808 MB through L2 -> SM for a 67 MB input, tensor pipe 44 %).  Run the ring kernel once per 64-channel chunk of the output
// instead: each pass keeps its 74 KB filter slice resident and re-reads the (narrow) input.
int conv2d_ring_try(const void* x, const void* w, const float* bias, const void* residual, void* y, int out_mode, int N, int H,
                    int W, int Ci, int Co, int Co_pad, int R, int S, int pad, float* stats, int stats_c, void* red_ws, cudaStream_t stream) {
    if (!ring_chunked_ok(out_mode, Co_pad) || W % 128 || Ci > 64 || residual || R != S || ring_slots(out_mode, Ci, Co_pad, R, S))
        return ring_launch(x, w, bias, residual, y, Co_pad, out_mode, N, H, W, Ci, Co, Co_pad, R, S, pad, stats, stats_c, red_ws, stream);
    if (stats || !conv2d_ring_eligible(out_mode, H, W, Ci, 64, R, S, false)) return -1;
    for (int c0 = 0; c0 < Co_pad; c0 += 64) {
        const int co = Co - c0 < 64 ? (Co - c0 < 1 ? 1 : Co - c0) : 64;
        const int e = ring_launch(x, static_cast<const char*>(w) + (size_t)c0 * R * S * Ci * 2, bias ? bias + c0 : nullptr, nullptr,
                                  static_cast<char*>(y) + (size_t)c0 * 2, Co_pad, out_mode, N, H, W, Ci, co, 64, R, S, pad, nullptr, 0, red_ws, stream);
        if (e) return e < 0 ? fail(FV_ERR_INTERNAL, "fv_conv2d: ring schedule refused a chunk it had accepted") : e;
    }
    return FV_OK;
}

}  // namespace fv
