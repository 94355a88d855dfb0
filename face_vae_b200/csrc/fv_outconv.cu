// Tap-folded schedule for the 7x7 output convolution (out_conv 32 -> 3, reference models.py:1099) at full resolution,
// forward + fused sigmoid / reconstruction loss, data gradient and weight gradient.
//
// The generic kernels spend one tcgen05.mma per filter tap: 49 taps x (Ci/16) instructions with N = 16 columns, each
// pinned at the ~40-cycle shared-memory operand-fetch floor of a 128-row A operand (profiles/r01_tcgen05_mma_rate.txt),
// i.e. ~2 % of the tensor peak.  Here the horizontal taps are folded into a GEMM dimension so that a 128-pixel row
// segment costs 7 (filter rows) x 2 (K steps) = 14 instructions:
//
//   forward   Q[u, (s,co)] = sum_{r,ci} X[y+r-3, u, ci] * W[co,r,s,ci]         N = 7*Co (<= 32), K = 7 rows x 32 channels
//             out[y, x, co] = bias[co] + sum_s Q[x+s-3, (s,co)]                 shifted sum in the epilogue (via smem);
//             u runs over the image columns only, so the horizontal zero padding is implicit.
//   dgrad     rec[y', u, (s',co)] = dY[y', u+s'-3, co]   ("records": 7 taps x 4 channels = 64 bytes per pixel, built in
//             shared memory by builder warps from the compact 4-channel gradient written by the loss epilogue)
//             dX[y, u, ci] = sum_{r'} sum_k rec[y+r'-3, u, k] * W[co, 6-r', 6-s', ci]       N = 32, K = 7 rows x 32
//   wgrad     dW[co,r,s,ci] = sum_{y,u} X[y+r-3, u, ci] * rec[y, u, (6-s,co)]   both operands MN-major, M = 4 filter rows x
//             32 channels stacked through the leading-dimension offset (one input-row slab), N = 32, K = pixels.
//
// Forward / dgrad: a CTA walks a run of image rows and keeps the last 7 input-row slabs (a whole row: W x 64 bytes) in
// a shared-memory ring, one new slab per output row (smem fill traffic == HBM traffic).  Wgrad: a CTA holds a window
// of 16 + 6 consecutive input-row slabs at consecutive addresses (so the stacked M tile never wraps) and streams the
// records of the 16 output rows past it; its two accumulators live in TMEM for the whole kernel.
#include <cstdio>
#include <cstdlib>

#include "../../include/facevae_b200.h"
#include "fv_host.h"
#include "fv_ptx.cuh"
#include "fv_reduce.cuh"

namespace fv {

// Role-loop timing for FV_TRACE builds: cycle deltas accumulate in REGISTERS and are flushed once when the role ends (the
// generic FV_TACC does a global read-modify-write per sample, ~300 cycles each, which swamps loops of ~1000 cycles).
#ifdef FV_TRACE
#define FVR_DECL long long fvr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define FVR_ACC(slot, var) fvr_acc[slot] += clock64() - (var)
#define FVR_FLUSH(ptr) do { if ((ptr) && threadIdx.x % 32 == 0) { for (int i_ = 0; i_ < 8; ++i_) if (fvr_acc[i_]) atomicAdd((unsigned long long*)&(ptr)[(blockIdx.x % 148) * 8 + i_], (unsigned long long)fvr_acc[i_]); } } while (0)
#else
#define FVR_DECL
#define FVR_ACC(slot, var)
#define FVR_FLUSH(ptr)
#endif

static constexpr int kFR = 7;            // filter rows == filter columns
static constexpr int kFC = 32;           // channels of the wide side (Ci of the forward conv) == record width
static constexpr int kRowB = kFC * 2;    // bytes per pixel row of a slab / record (one 64-byte swizzle span)
static constexpr int kWBytes = kFR * 32 * kRowB;   // folded filter: 7 x [32 x 32] bf16
static constexpr int kAcc = 4;           // TMEM accumulator buffers (64 columns each) of the forward / dgrad kernel

struct FoldParams {
    int N, H, W, halves, Co, NQ, QS;
    int rows_total, rows_per_cta, ring, slab_bytes;
    int w_off, q_off, t_off, bar_off;
    // forward epilogue
    const float* bias;
    float* logits;                  // NCHW fp32 (optional when the loss is fused)
    const float* target;            // NCHW fp32: non-null fuses sigmoid + loss + gradient
    float* pred;                    // NCHW fp32
    __nv_bfloat16* g4;              // [N,H,W,4] bf16: gscale * dloss/dlogits
    float* loss_sum;                // [1]   (written: grid total in a fixed summation order)
    float* gsum;                    // [4]: per-channel sums of the gradient (bias gradient)
    void* red_ws;                   // fv_reduce.cuh workspace (with target)
    int l1, use_sigmoid;
    float gscale;
    // dgrad
    const __nv_bfloat16* dy4;
    __nv_bfloat16* dx;
    const float* scale_ptr;
    long long* trace;
    int dbg;                        // FV_FOLD_DEBUG bit mask (tools/fold_experiments.py only): 1 no MMAs, 2 no epilogue work, 4 no slab fill
};

__device__ __forceinline__ uint32_t swz64(uint32_t byte_addr) {          // 64-byte swizzle: 16-byte chunk ^= address bits [7,9)
    return byte_addr ^ ((byte_addr >> 3) & 0x30u);
}

// 7 taps x 4 channels of a compact gradient row around column u -> one 64-byte record (zero outside the row).
// Split into the global loads and the swizzled shared-memory store so that a builder warp can have the loads of a whole
// slab in flight before it waits for the slot.
struct Rec {
    uint2 v[7];
};
__device__ __forceinline__ void rec_load(Rec& r, const __nv_bfloat16* __restrict__ row4, bool row_ok, int u, int W) {
#pragma unroll
    for (int s = 0; s < 7; ++s) {
        const int xx = u + s - 3;
        r.v[s] = (row_ok && xx >= 0 && xx < W) ? __ldg(reinterpret_cast<const uint2*>(row4) + xx) : make_uint2(0u, 0u);
    }
}
__device__ __forceinline__ void rec_store(const Rec& r, uint32_t rec_addr) {       // rec_addr: unswizzled byte address of the record
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t a = swz64(rec_addr + j * 16);
        const uint2 lo = r.v[2 * j], hi = (j < 3) ? r.v[2 * j + 1] : make_uint2(0u, 0u);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(lo.x), "r"(lo.y), "r"(hi.x), "r"(hi.y) : "memory");
    }
}

// MODE 0: forward (slabs by TMA, shifted-sum epilogue).  MODE 1: data gradient (slabs = records built by warps 10..17).
template <int MODE>
__global__ void __launch_bounds__(MODE == 0 ? 352 : 608, 1)
fold_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const FoldParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.bar_off);
    uint64_t* empty = full + p.ring;
    uint64_t* tfull = empty + p.ring;
    uint64_t* tempty = tfull + 2 * kAcc;                 // [kAcc][2 halves]
    uint64_t* wfull = tempty + 2 * kAcc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
    float* Qs = reinterpret_cast<float*>(smem + p.q_off);
    __shared__ float fred[8][8];                         // per epilogue warp: loss | - | - | - | gradient sums x4
    __shared__ float fblk[8], ftot[8];
    __shared__ int fflag;

    constexpr int kIssuerB = MODE == 0 ? 10 : 18;       // second MMA issuer warp (the last warp of the CTA)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g0 = blockIdx.x * p.rows_per_cta;
    const int g1 = min(g0 + p.rows_per_cta, p.rows_total);
#ifdef FV_TRACE
    long long* fv_trace = p.trace;
#endif

    if (warp == 0 && lane == 0) {
        if (MODE == 0) tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
        for (int i = 0; i < p.ring; ++i) {
            mbar_init(&full[i], MODE == 0 ? 1 : 32);
            mbar_init(&empty[i], p.halves);              // one vote per issuer warp
        }
        for (int i = 0; i < 2 * kAcc; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);
        }
        mbar_init(wfull, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 64 * kAcc);
        tmem_relinquish();
    }
    if (MODE == 0) {                       // zero padding rows of the exchange buffer: u + 3 in [0,3) and [W+3, W+6)
        for (int i = threadIdx.x; i < 6 * p.QS; i += blockDim.x) {
            const int row = i / p.QS, c = i - row * p.QS;
            Qs[(row < 3 ? row : p.W + row) * p.QS + c] = 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        const bool leader = elect_one_sync();
        if (leader) {
            mbar_arrive_expect_tx(wfull, (uint32_t)kWBytes);
            tma_load_2d(smem + p.w_off, &tmW, wfull, 0, 0);
        }
        if (MODE == 0) {
            uint32_t slot = 0, ph = 0;
            for (int g = g0; g < g1; ++g) {
                const int n = g / p.H, y = g - n * p.H;
                const bool fresh = (g == g0) || (y == 0);
                for (int j = fresh ? 0 : kFR - 1; j < kFR; ++j) {
                    mbar_wait(&empty[slot], ph ^ 1);
                    if (leader) {
                        if (p.dbg & 4) {
                            mbar_arrive(&full[slot]);
                        } else {
                            mbar_arrive_expect_tx(&full[slot], (uint32_t)p.slab_bytes);
                            tma_load_4d(smem + (size_t)slot * p.slab_bytes, &tmA, &full[slot], 0, 0, y - 3 + j, n);
                        }
                    }
                    if (++slot == (uint32_t)p.ring) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1 || warp == kIssuerB) {
        // Two MMA issuer warps, one per 128-pixel half of the row.  The issuing thread is blocked for the ~45 cycles of
        // every tcgen05.mma (tools/fold_experiments.py: time is linear in the MMA count, with the per-row bookkeeping --
        // barrier waits, descriptor words, commits, ~600-1000 cycles -- purely additive), so a single issuer leaves the
        // tensor pipe idle during its bookkeeping; with two, one warp's MMAs run during the other's bookkeeping.
        const uint32_t ih = warp == 1 ? 0u : 1u;
        if ((int)ih < p.halves) {
            const bool leader = elect_one_sync();
            const uint32_t idesc = umma_idesc_bf16(128, 32, 0, 0);
            const uint64_t tmpl = umma_smem_desc(0, 16, 8u * kRowB, 4u);          // K-major, 64-byte swizzle
            const uint32_t desc_hi = (uint32_t)(tmpl >> 32), lo_base = (uint32_t)tmpl;
            const uint32_t smem_base = smem_u32(smem);
            const uint32_t w_lo = lo_base | ((smem_base + (uint32_t)p.w_off) >> 4);
            const uint32_t a_half = ih * (uint32_t)((128 * kRowB) >> 4);
            uint32_t first = 0, wslot = 0, wph = 0, tcount = 0;
            uint32_t probe_full = 0, probe_te = 0;       // the next row's barriers, probed before this row's MMAs were issued
            int y = g0 % p.H;
            FVR_DECL;
            mbar_wait(wfull, 0);
            FV_T0(t_all);
            for (int g = g0; g < g1; ++g, ++tcount) {
                const bool fresh = (g == g0) || (y == 0);
                const bool next_fresh = (g + 1 == g1) || (y + 1 == p.H);
                if (++y == p.H) y = 0;
                if (fresh) first = wslot;
                {
                    FV_T0(tw);
                    for (int i = fresh ? 0 : kFR - 1; i < kFR; ++i) {
                        if (fresh || !probe_full) mbar_wait(&full[wslot], wph);
                        if (++wslot == (uint32_t)p.ring) { wslot = 0; wph ^= 1; }
                    }
                    if (ih == 0) FVR_ACC(1, tw);
                }
                const uint32_t acc = tcount % kAcc, aph = (tcount / kAcc) & 1;
                { FV_T0(tw); if (!probe_te) mbar_wait(&tempty[acc * 2 + ih], aph ^ 1); if (ih == 0) FVR_ACC(2, tw); }
                tc_fence_after();
                // probe the NEXT row's barriers now: the results are consumed after this row's MMAs have been issued
                probe_full = probe_te = 0;
                if (g + 1 < g1) {
                    const uint32_t nacc = (tcount + 1) % kAcc, naph = ((tcount + 1) / kAcc) & 1;
                    if (!next_fresh) probe_full = mbar_test_wait(&full[wslot], wph);
                    probe_te = mbar_test_wait(&tempty[nacc * 2 + ih], naph ^ 1);
                }
                FV_T0(t_desc);
                // all descriptor words first, then ONE predicated block of back-to-back MMAs: a per-MMA address computation
                // in front of every tcgen05.mma is a serial IMAD -> R2UR -> uniform-ALU chain of ~50 cycles (ncu, round 1)
                uint32_t a_lo[kFR];
                {
                    uint32_t slot = first;
#pragma unroll
                    for (int r = 0; r < kFR; ++r) {
                        a_lo[r] = (lo_base | ((smem_base + slot * (uint32_t)p.slab_bytes) >> 4)) + a_half;
                        if (++slot == (uint32_t)p.ring) slot = 0;
                    }
                }
                const uint32_t d_tmem = tmem_base + acc * 64u + ih * 32u;
                if (ih == 0) FVR_ACC(6, t_desc);
                FV_T0(t_issue);
                if (leader && !(p.dbg & 1)) {
#pragma unroll
                    for (int r = 0; r < kFR; ++r)
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            tc_mma_f16_lohi(d_tmem, a_lo[r] + 2 * k, w_lo + (uint32_t)(r * ((32 * kRowB) >> 4) + 2 * k), desc_hi, idesc,
                                            (uint32_t)(r | k));
                }
                if (ih == 0) FVR_ACC(3, t_issue);
                FV_T0(t_commit);
                if (leader) tc_commit(&tfull[acc * 2 + ih]);
                const int n_rel = next_fresh ? kFR : 1;
                for (int i = 0; i < n_rel; ++i) {
                    if (leader) tc_commit(&empty[first]);
                    if (++first == (uint32_t)p.ring) first = 0;
                }
                if (ih == 0) FVR_ACC(4, t_commit);
            }
            if (ih == 0) FVR_ACC(5, t_all);
            FVR_FLUSH(p.trace);
        }
    } else if (warp < 10) {
        const int q = warp & 3, h = (warp - 2) >> 2;
        if (h < p.halves) {
            const int u = h * 128 + q * 32 + lane;
            float loss_acc = 0.f, gs[4] = {0.f, 0.f, 0.f, 0.f};
            float bias_r[4] = {0.f, 0.f, 0.f, 0.f};
            if (MODE == 0 && p.bias) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < p.Co) bias_r[c] = __ldg(p.bias + c);
            }
            const float dscale = (MODE == 1 && p.scale_ptr) ? __ldg(p.scale_ptr) : 1.f;
            const int nthr = 128 * p.halves;
            const size_t plane = (size_t)p.H * p.W;
            // The targets of a row are fetched three rows ahead with cp.async into a per-thread slot of shared memory.  Register
            // prefetching does not work here: however far ahead the loads are issued (1, 2 and 4 rows were measured), the
            // warp ends up waiting a full memory latency per row -- with six scoreboards per warp the long-lived load shares
            // one with the row's own tcgen05.ld / LDS traffic and is waited on at its next reuse.  cp.async completion is
            // tracked per thread by commit groups instead.
            constexpr int kTRows = 3;
            float* Ts = reinterpret_cast<float*>(smem + p.t_off);
            const bool fused = MODE == 0 && p.target != nullptr;
            auto fetch_targets = [&](int gg, int slot) {
                if (fused && gg < g1 && !(p.dbg & 32)) {
                    const int nn = gg / p.H, yy = gg - nn * p.H;
                    const float* tp = p.target + (size_t)nn * p.Co * plane + (size_t)yy * p.W + u;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < p.Co)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(Ts + (slot * p.Co + c) * p.W + u)),
                                         "l"(tp + c * plane)
                                         : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");          // one group per row, empty or not
            };
            if (fused)
                for (int i = 0; i < kTRows; ++i) fetch_targets(g0 + i, i);
            uint32_t tcount = 0;
            float t_cur[4] = {0.f, 0.f, 0.f, 0.f};
            for (int g = g0; g < g1; ++g, ++tcount) {
                const uint32_t acc = tcount % kAcc, aph = (tcount / kAcc) & 1;
                mbar_wait(&tfull[acc * 2 + h], aph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * 64u + (uint32_t)h * 32u;
                uint32_t v0[16], v1[16];
                tmem_ld16(taddr, v0);
                tmem_ld16(taddr + 16, v1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc * 2 + h]);
                if (p.dbg & 2) continue;
                if (MODE == 1) {
                    // one pixel = 64 contiguous bytes per thread: two 256-bit stores (full 32-byte sectors; 16-byte stores at a
                    // 64-byte lane stride send half-written sectors to L2: 264 MB for a 134 MB tensor in the round-1 profile)
                    uint32_t w[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        w[i] = pack_bf16(dscale * __uint_as_float(v0[2 * i]), dscale * __uint_as_float(v0[2 * i + 1]));
                        w[8 + i] = pack_bf16(dscale * __uint_as_float(v1[2 * i]), dscale * __uint_as_float(v1[2 * i + 1]));
                    }
                    __nv_bfloat16* o = p.dx + ((size_t)g * p.W + u) * kFC;
                    st_global_256(o, w);
                    st_global_256(o + 16, w + 8);
                } else {
                    const int n = g / p.H, y = g - n * p.H;
                    float* qrow = Qs + (u + 3) * p.QS;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < p.NQ) qrow[j] = __uint_as_float(v0[j]);
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (16 + j < p.NQ) qrow[16 + j] = __uint_as_float(v1[j]);
                    named_bar_sync(1, nthr);
                    float o[4] = {bias_r[0], bias_r[1], bias_r[2], bias_r[3]};
#pragma unroll
                    for (int s = 0; s < kFR; ++s) {
                        const float* qs = Qs + (u + s) * p.QS + s * p.Co;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (c < p.Co) o[c] += qs[c];
                    }
                    named_bar_sync(2, nthr);                       // all reads done before the next row's writes
                    const size_t idx0 = (size_t)n * p.Co * plane + (size_t)y * p.W + u;
                    if (p.logits) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (c < p.Co) p.logits[idx0 + c * plane] = o[c];
                    }
                    if (p.target) {
                        const int tslot = (int)(tcount % kTRows);
                        asm volatile("cp.async.wait_group %0;" ::"n"(kTRows - 1) : "memory");     // this row's group has landed
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (c < p.Co) t_cur[c] = Ts[(tslot * p.Co + c) * p.W + u];
                        fetch_targets(g + kTRows, tslot);                                       // refill the slot just consumed
                        float gd4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (c < p.Co) {
                            const float t = t_cur[c];
                            const float sg = p.use_sigmoid ? 1.f / (1.f + expf(-o[c])) : o[c];
                            const float d = sg - t;
                            loss_acc += p.l1 ? fabsf(d) : d * d;
                            float gd = p.l1 ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : 2.f * d;
                            if (p.use_sigmoid) gd *= sg * (1.f - sg);
                            gd *= p.gscale;
                            if (p.pred && !(p.dbg & 64)) p.pred[idx0 + c * plane] = sg;
                            gd4[c] = gd;
                            gs[c] += gd;
                            }
                        }
                        if (p.g4 && !(p.dbg & 64))
                            *reinterpret_cast<uint2*>(p.g4 + ((size_t)g * p.W + u) * 4) =
                                make_uint2(pack_bf16(gd4[0], gd4[1]), pack_bf16(gd4[2], gd4[3]));
                    }
                }
            }
            if (MODE == 0 && p.target) {                 // per-warp partials; combined in warp order at the end of the kernel
                loss_acc = warp_sum(loss_acc);
                if (lane == 0) fred[warp - 2][0] = loss_acc;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float t = warp_sum(gs[c]);
                    if (lane == 0) fred[warp - 2][4 + c] = c < p.Co ? t : 0.f;
                }
            }
        } else if (MODE == 0 && lane == 0) {
            fred[warp - 2][0] = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) fred[warp - 2][4 + c] = 0.f;
        }
    } else if (MODE == 1 && warp < 18) {
        // record builders (warps 10..17): the forward producer's slab sequence, slab k built by warp k % 8, so that up to
        // (ring - 7) slabs are in flight; the global loads of a slab are issued before the wait for its slot
        const uint32_t bw = (uint32_t)(warp - 10);
        const uint32_t smem_base = smem_u32(smem);
        uint32_t slot = 0, ph = 0, k = 0;
        for (int g = g0; g < g1; ++g) {
            const int n = g / p.H, y = g - n * p.H;
            const bool fresh = (g == g0) || (y == 0);
            for (int j = fresh ? 0 : kFR - 1; j < kFR; ++j, ++k) {
                if ((k & 7u) == bw) {
                    const int yy = y - 3 + j;
                    const bool row_ok = yy >= 0 && yy < p.H;
                    const __nv_bfloat16* row4 = p.dy4 + ((size_t)n * p.H + (row_ok ? yy : 0)) * p.W * 4;
                    const uint32_t slab = smem_base + slot * (uint32_t)p.slab_bytes;
                    if (p.dbg & 4) {
                        mbar_wait(&empty[slot], ph ^ 1);
                        mbar_arrive(&full[slot]);
                        if (++slot == (uint32_t)p.ring) { slot = 0; ph ^= 1; }
                        continue;
                    }
                    Rec ra[2], rb[2];
                    rec_load(ra[0], row4, row_ok, lane, p.W);
                    rec_load(ra[1], row4, row_ok, lane + 32, p.W);
                    mbar_wait(&empty[slot], ph ^ 1);
                    for (int u0 = 0; u0 < p.W; u0 += 128) {                 // 4 records per lane and pass, double-buffered
                        rec_load(rb[0], row4, row_ok, u0 + 64 + lane, p.W);
                        rec_load(rb[1], row4, row_ok, u0 + 96 + lane, p.W);
                        rec_store(ra[0], slab + (uint32_t)(u0 + lane) * kRowB);
                        rec_store(ra[1], slab + (uint32_t)(u0 + 32 + lane) * kRowB);
                        if (u0 + 128 < p.W) {
                            rec_load(ra[0], row4, row_ok, u0 + 128 + lane, p.W);
                            rec_load(ra[1], row4, row_ok, u0 + 160 + lane, p.W);
                        }
                        rec_store(rb[0], slab + (uint32_t)(u0 + 64 + lane) * kRowB);
                        rec_store(rb[1], slab + (uint32_t)(u0 + 96 + lane) * kRowB);
                    }
                    fence_proxy_async();
                    mbar_arrive(&full[slot]);
                }
                if (++slot == (uint32_t)p.ring) { slot = 0; ph ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64 * kAcc);
    }
    if (MODE == 0 && p.target) {                         // warps in order -> CTA totals -> CTAs in order (fv_reduce.cuh): reproducible
        if (threadIdx.x < 8) {
            float a = 0.f;
            if (threadIdx.x == 0 || threadIdx.x >= 4)
                for (int w = 0; w < 8; ++w) a += fred[w][threadIdx.x];
            fblk[threadIdx.x] = a;
        }
        __syncthreads();
        if (det_reduce<float>(p.red_ws, 8, gridDim.x, blockIdx.x, fblk, ftot, threadIdx.x, blockDim.x, BlockSync{}, &fflag)) {
            if (threadIdx.x == 0) p.loss_sum[0] = ftot[0];
            if (p.gsum && threadIdx.x >= 4 && threadIdx.x < 4 + p.Co) p.gsum[threadIdx.x - 4] = ftot[threadIdx.x];
        }
    }
}

// ------------------------------------------------------------------------------------------------ weight gradient
static constexpr int kWT = 16;                     // output rows per window
static constexpr int kWSlots = kWT + 7;            // input-row slabs of a window (+1: the unused 4th chunk of M tile 1)
static constexpr int kWSlab = 128 * kRowB;         // 8 KB: one half row
static constexpr int kRecSlots = 5;

struct FoldWgradParams {
    int N, H, W, halves, Co;
    int wins_per_col, units_total, units_per_cta;
    int rec_off, bar_off;
    const __nv_bfloat16* dy4;
    float* dw;                     // [CTAs][Co][32][7][7] fp32 partial slabs (stored; fv_slab_sum adds them in CTA order)
    long long slab_stride;
    const float* scale_ptr;
    long long* trace;
    int dbg;                       // FV_FOLD_DEBUG (tools/fold_experiments.py only): 1 no MMAs, 4 no record build, 8 no slab loads
};

__global__ void __launch_bounds__(352, 1)
fold_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const FoldWgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.bar_off);     // [kWSlots]
    uint64_t* empty = full + kWSlots;                                    // [kWSlots]
    uint64_t* rfull = empty + kWSlots;                                   // [kRecSlots]
    uint64_t* rempty = rfull + kRecSlots;                                // [kRecSlots]
    uint64_t* tfull = rempty + kRecSlots;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u0 = blockIdx.x * p.units_per_cta;
    const int u1 = min(u0 + p.units_per_cta, p.units_total);
#ifdef FV_TRACE
    long long* fv_trace = p.trace;
#endif

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        for (int i = 0; i < kWSlots; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 2);                     // one vote per issuer warp
        }
        for (int i = 0; i < kRecSlots; ++i) {
            mbar_init(&rfull[i], 32);
            mbar_init(&rempty[i], 2);
        }
        mbar_init(tfull, 2);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // unit -> (image n, half h, window start y0, rows T)
    auto decode = [&](int unit, int& n, int& h, int& y0, int& T) {
        const int col = unit / p.wins_per_col, wi = unit - col * p.wins_per_col;
        n = col / p.halves;
        h = col - n * p.halves;
        y0 = wi * kWT;
        T = min(kWT, p.H - y0);
    };

    if (warp == 0) {
        const bool leader = elect_one_sync();
        uint32_t phmask = 0;                                  // bit i: parity of the NEXT wait on empty[i] is (bit ^ 1)
        for (int unit = u0; unit < u1; ++unit) {
            int n, h, y0, T;
            decode(unit, n, h, y0, T);
            for (int i = 0; i < T + 6; ++i) {
                mbar_wait(&empty[i], ((phmask >> i) & 1u) ^ 1u);
                phmask ^= 1u << i;
                if (leader) {
                    if (p.dbg & 8) {
                        mbar_arrive(&full[i]);
                    } else {
                        mbar_arrive_expect_tx(&full[i], (uint32_t)kWSlab);
                        tma_load_4d(smem + (size_t)i * kWSlab, &tmX, &full[i], 0, h * 128, y0 - 3 + i, n);
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 10) {
        // two issuer warps, one per M tile (filter rows 0-3 / 4-6): see fold_conv_kernel
        const uint32_t mt = warp == 1 ? 0u : 1u;
        if (u0 < u1) {
            const bool leader = elect_one_sync();
            const uint32_t idesc = umma_idesc_bf16(128, 32, 1, 1);                       // both operands MN-major
            const uint64_t a_tmpl = umma_smem_desc(0, (uint32_t)kWSlab, 8u * kRowB, 4u);  // chunk i of the M tile = slab + i
            const uint64_t b_tmpl = umma_smem_desc(0, 16, 8u * kRowB, 4u);
            const uint32_t a_hi = (uint32_t)(a_tmpl >> 32), a_lo_base = (uint32_t)a_tmpl;
            const uint32_t b_hi = (uint32_t)(b_tmpl >> 32), b_lo_base = (uint32_t)b_tmpl;
            const uint32_t smem_base = smem_u32(smem);
            constexpr uint32_t kstep = (16u * kRowB) >> 4;                                // 16 pixels per MMA
            uint32_t phmask = 0, rs = 0, rph = 0, accumulate = 0;
            uint32_t probe_full = 0, probe_rf = 0;       // the next row's barriers, probed before this row's MMAs were issued
            FVR_DECL;
            FV_T0(t_all);
            for (int unit = u0; unit < u1; ++unit) {
                int n, h, y0, T;
                decode(unit, n, h, y0, T);
                for (int j = 0; j < T; ++j) {
                    {
                        FV_T0(tw);
                        for (int i = (j == 0 ? 0 : j + 6); i <= j + 6; ++i) {
                            if (j == 0 || !probe_full) mbar_wait(&full[i], (phmask >> i) & 1u);
                            phmask ^= 1u << i;
                        }
                        if (mt == 0) FVR_ACC(1, tw);
                    }
                    { FV_T0(tw); if (!probe_rf) mbar_wait(&rfull[rs], rph); if (mt == 0) FVR_ACC(2, tw); }
                    tc_fence_after();
                    // probe the NEXT row's barriers now: the results are consumed after this row's MMAs have been issued
                    probe_full = probe_rf = 0;
                    {
                        const uint32_t nrs = rs + 1 == kRecSlots ? 0u : rs + 1, nrph = rs + 1 == kRecSlots ? rph ^ 1u : rph;
                        if (j + 1 < T) {
                            probe_full = mbar_test_wait(&full[j + 7], (phmask >> (j + 7)) & 1u);
                            probe_rf = mbar_test_wait(&rfull[nrs], nrph);
                        } else if (unit + 1 < u1) {
                            probe_rf = mbar_test_wait(&rfull[nrs], nrph);
                        }
                    }
                    FV_T0(t_issue);
                    const uint32_t b_lo = b_lo_base | ((smem_base + (uint32_t)p.rec_off + rs * (uint32_t)kWSlab) >> 4);
                    const uint32_t a_lo = a_lo_base | ((smem_base + ((uint32_t)j + 4u * mt) * (uint32_t)kWSlab) >> 4);
                    if (leader && !(p.dbg & 1)) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            tc_mma_f16_lohi2(tmem_base + mt * 32u, a_lo + k * kstep, a_hi, b_lo + k * kstep, b_hi, idesc,
                                             accumulate | (uint32_t)(k > 0));
                    }
                    accumulate = 1;
                    if (mt == 0) FVR_ACC(3, t_issue);
                    FV_T0(t_commit);
                    if (leader) tc_commit(&rempty[rs]);
                    if (++rs == kRecSlots) { rs = 0; rph ^= 1; }
                    if (leader) tc_commit(&empty[j]);                       // slab j: last used by output row j
                    if (j == T - 1)
                        for (int i = T; i < T + 6; ++i)
                            if (leader) tc_commit(&empty[i]);
                    if (mt == 0) FVR_ACC(4, t_commit);
                }
            }
            if (leader) tc_commit(tfull);
            if (mt == 0) FVR_ACC(5, t_all);
            FVR_FLUSH(p.trace);
        }
    } else if (warp < 6) {
        if (u0 < u1) {
            const int q = warp & 3;
            const float sc = p.scale_ptr ? __ldg(p.scale_ptr) : 1.f;
            mbar_wait(tfull, 0);
            tc_fence_after();
            for (int mt = 0; mt < 2; ++mt) {
                const int r = 4 * mt + q;
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + (uint32_t)mt * 32u;
                uint32_t v0[16], v1[16];
                tmem_ld16(taddr, v0);
                tmem_ld16(taddr + 16, v1);
                tmem_ld_wait();
                if (r < kFR) {
#pragma unroll
                    for (int k = 0; k < 28; ++k) {
                        const int sp = k >> 2, co = k & 3;
                        if (co < p.Co) {
                            const float v = __uint_as_float(k < 16 ? v0[k & 15] : v1[k & 15]) * sc;
                            p.dw[(size_t)blockIdx.x * p.slab_stride + (((size_t)co * kFC + lane) * kFR + r) * kFR + (6 - sp)] = v;
                        }
                    }
                }
            }
        }
    } else if (warp < 10) {
        // record builders (warps 6..9): the records of output row k (in processing order) are built by warp k % 4, four
        // records per lane, with all global loads issued before the wait for the slot
        const uint32_t bw = (uint32_t)(warp - 6);
        const uint32_t smem_base = smem_u32(smem);
        uint32_t rs = 0, rph = 0, k = 0;
        for (int unit = u0; unit < u1; ++unit) {
            int n, h, y0, T;
            decode(unit, n, h, y0, T);
            for (int j = 0; j < T; ++j, ++k) {
                if ((k & 3u) == bw) {
                    const __nv_bfloat16* row4 = p.dy4 + ((size_t)n * p.H + y0 + j) * p.W * 4;
                    Rec r[4];
                    if (!(p.dbg & 4)) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) rec_load(r[i], row4, true, h * 128 + 32 * i + lane, p.W);
                    }
                    mbar_wait(&rempty[rs], rph ^ 1);
                    const uint32_t base = smem_base + (uint32_t)p.rec_off + rs * (uint32_t)kWSlab;
                    if (!(p.dbg & 4)) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) rec_store(r[i], base + (uint32_t)(32 * i + lane) * kRowB);
                    }
                    fence_proxy_async();
                    mbar_arrive(&rfull[rs]);
                }
                if (++rs == kRecSlots) { rs = 0; rph ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

// ------------------------------------------------------------------------------------------------ filter operands
// w fp32 [Co][32][7][7] -> wq  [7 r ][32 rows j = s*Co + co][32 ci] (forward B operand, K-major)
//                          wdq [7 r'][32 rows ci           ][32 k = s'*4 + co] = w[co][ci][6-r'][6-s'] (dgrad B operand)
__global__ void outconv_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wq, __nv_bfloat16* __restrict__ wdq,
                                    int Co, int Ci) {
    const int total = kFR * 32 * 32;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i & 31, row = (i >> 5) & 31, r = i >> 10;
        if (wq) {
            const int s = row / Co, co = row - s * Co;
            const float v = (row < kFR * Co && c < Ci) ? w[(((size_t)co * Ci + c) * kFR + r) * kFR + s] : 0.f;
            wq[i] = __float2bfloat16(v);
        }
        if (wdq) {
            const int sp = c >> 2, co = c & 3;
            const float v = (sp < kFR && co < Co && row < Ci) ? w[(((size_t)co * Ci + row) * kFR + (6 - r)) * kFR + (6 - sp)] : 0.f;
            wdq[i] = __float2bfloat16(v);
        }
    }
}

// FV_OUTCONV_MAX_CTAS (tests): fewer CTAs -> longer runs per CTA, so small shapes exercise ring / window wrap-around
static int fold_ctas() {
    const char* v = getenv("FV_OUTCONV_MAX_CTAS");
    const int cap = v ? atoi(v) : 0;
    const int sms = num_sms();
    return (cap > 0 && cap < sms) ? cap : sms;
}

static int fold_geometry(FoldParams& p, int N, int H, int W, int mode) {
    p.N = N; p.H = H; p.W = W; p.halves = W / 128;
    { const char* v = getenv("FV_FOLD_DEBUG"); p.dbg = v ? atoi(v) : 0; }
    p.rows_total = N * H;
    const int sms = fold_ctas();
    p.rows_per_cta = (p.rows_total + sms - 1) / sms;
    p.slab_bytes = W * kRowB;
    const int q_bytes = mode == 0 ? (((W + 6) * p.QS * 4 + 1023) & ~1023) : 0;
    const int t_bytes = (mode == 0 && p.target) ? 3 * p.Co * W * 4 : 0;          // cp.async staging of three target rows
    int ring = (221 * 1024 - kWBytes - q_bytes - t_bytes) / p.slab_bytes;
    if (ring > 16) ring = 16;
    if (ring < kFR + 1) return -1;
    p.ring = ring;
    p.w_off = ring * p.slab_bytes;
    p.q_off = p.w_off + kWBytes;
    p.t_off = p.q_off + q_bytes;
    p.bar_off = p.t_off + t_bytes;
    return (p.rows_total + p.rows_per_cta - 1) / p.rows_per_cta;
}

static int encode_w(CUtensorMap* tm, const void* wq) {
    uint64_t dims[2] = {32, (uint64_t)kFR * 32};
    uint64_t str[1] = {64};
    uint32_t box[2] = {32, (uint32_t)kFR * 32};
    return encode_tmap_bf16(tm, wq, 2, dims, str, box, kRowB);
}

static bool fold_shape_ok(int N, int H, int W, int Ci, int Co, int R, int S) {
    return N >= 1 && H >= 1 && (W == 128 || W == 256) && Ci == kFC && Co >= 1 && Co <= 4 && R == kFR && S == kFR &&
           (long long)N * H * W < (1LL << 30);
}

}  // namespace fv

using namespace fv;

extern "C" __attribute__((visibility("default"))) int fv_outconv_supported(int N, int H, int W, int Ci, int Co, int R, int S) {
    return fold_shape_ok(N, H, W, Ci, Co, R, S) ? 1 : 0;
}

extern "C" __attribute__((visibility("default"))) int fv_outconv_prep(const float* w, void* wq, void* wdq, int Co, int Ci, void* stream) {
    if (!w || (!wq && !wdq) || Co < 1 || Co > 4 || Ci < 1 || Ci > kFC) return fail(FV_ERR_ARG, "fv_outconv_prep: bad arguments (Co=%d Ci=%d)", Co, Ci);
    outconv_prep_kernel<<<7, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wq, (__nv_bfloat16*)wdq, Co, Ci);
    FV_LAUNCH_CHECK("outconv_prep_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_outconv_fwd(const void* x, const void* wq, const float* bias, float* logits, const float* target,
                                                                    float* pred, void* g4, float* loss_sum, float* gsum, int N, int H, int W,
                                                                    int Ci, int Co, int l1, int use_sigmoid, float gscale, void* red_ws, void* stream) {
    if (!x || !wq || (!logits && !target)) return fail(FV_ERR_ARG, "fv_outconv_fwd: null pointer");
    if (target && (!loss_sum || !red_ws)) return fail(FV_ERR_ARG, "fv_outconv_fwd: the fused loss needs loss_sum and the reduction workspace");
    if (!fold_shape_ok(N, H, W, Ci, Co, kFR, kFR))
        return fail(FV_ERR_UNSUPPORTED, "fv_outconv_fwd: needs a 7x7 filter, Ci = 32, Co <= 4, W = 128 or 256 (got Ci=%d Co=%d W=%d)", Ci, Co, W);
    FoldParams p{};
    p.Co = Co; p.NQ = kFR * Co; p.QS = p.NQ | 1;
    p.target = target;                           // sizes the target staging area
    const int grid = fold_geometry(p, N, H, W, 0);
    if (grid < 1) return fail(FV_ERR_INTERNAL, "fv_outconv_fwd: shared-memory budget");
    p.bias = bias; p.logits = logits; p.target = target; p.pred = pred; p.g4 = (__nv_bfloat16*)g4; p.loss_sum = loss_sum; p.gsum = gsum; p.red_ws = red_ws;
    p.l1 = l1; p.use_sigmoid = use_sigmoid; p.gscale = gscale; p.trace = trace_ptr();
    CUtensorMap tmA, tmW;
    {
        uint64_t dims[4] = {(uint64_t)kFC, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)kRowB, (uint64_t)W * kRowB, (uint64_t)H * W * kRowB};
        uint32_t box[4] = {(uint32_t)kFC, (uint32_t)W, 1, 1};
        if (int e = encode_tmap_bf16(&tmA, x, 4, dims, str, box, kRowB)) return e;
    }
    if (int e = encode_w(&tmW, wq)) return e;
    const size_t smem = (size_t)p.bar_off + (2 * p.ring + 4 * kAcc + 1) * 8 + 16 + 1024 + 64;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncAttributes fa;                      // the opt-in limit covers static + dynamic shared memory
        FV_CUDA(cudaFuncGetAttributes(&fa, fold_conv_kernel<0>));
        FV_CUDA(cudaFuncSetAttribute(fold_conv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes));
        attr_set = true;
    }
    fold_conv_kernel<0><<<grid, 352, smem, (cudaStream_t)stream>>>(tmA, tmW, p);
    FV_LAUNCH_CHECK("fold_conv_kernel<fwd>");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_outconv_dgrad(const void* dy4, const void* wdq, const float* scale_ptr, void* dx, int N, int H, int W,
                                                                      int Ci, int Co, void* stream) {
    if (!dy4 || !wdq || !dx) return fail(FV_ERR_ARG, "fv_outconv_dgrad: null pointer");
    if (!fold_shape_ok(N, H, W, Ci, Co, kFR, kFR))
        return fail(FV_ERR_UNSUPPORTED, "fv_outconv_dgrad: needs a 7x7 filter, Ci = 32, Co <= 4, W = 128 or 256 (got Ci=%d Co=%d W=%d)", Ci, Co, W);
    FoldParams p{};
    p.Co = Co; p.NQ = 0; p.QS = 1;
    const int grid = fold_geometry(p, N, H, W, 1);
    if (grid < 1) return fail(FV_ERR_INTERNAL, "fv_outconv_dgrad: shared-memory budget");
    p.dy4 = (const __nv_bfloat16*)dy4; p.dx = (__nv_bfloat16*)dx; p.scale_ptr = scale_ptr; p.trace = trace_ptr();
    CUtensorMap tmW;
    if (int e = encode_w(&tmW, wdq)) return e;
    const size_t smem = (size_t)p.bar_off + (2 * p.ring + 4 * kAcc + 1) * 8 + 16 + 1024 + 64;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncAttributes fa;                      // the opt-in limit covers static + dynamic shared memory
        FV_CUDA(cudaFuncGetAttributes(&fa, fold_conv_kernel<1>));
        FV_CUDA(cudaFuncSetAttribute(fold_conv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes));
        attr_set = true;
    }
    fold_conv_kernel<1><<<grid, 608, smem, (cudaStream_t)stream>>>(tmW, tmW, p);
    FV_LAUNCH_CHECK("fold_conv_kernel<dgrad>");
    return FV_OK;
}

static int fold_wgrad_grid(int N, int H, int W, int& units_total, int& units_per_cta) {
    units_total = N * (W / 128) * ((H + kWT - 1) / kWT);
    const int sms = fold_ctas();
    units_per_cta = (units_total + sms - 1) / sms;
    return (units_total + units_per_cta - 1) / units_per_cta;
}
// partial slabs ([Co][32][7][7] floats each) fv_outconv_wgrad writes for this shape
extern "C" __attribute__((visibility("default"))) int fv_outconv_wgrad_splits(int N, int H, int W) {
    int a, b;
    return (N < 1 || H < 1 || (W != 128 && W != 256)) ? 0 : fold_wgrad_grid(N, H, W, a, b);
}

extern "C" __attribute__((visibility("default"))) int fv_outconv_wgrad(const void* x, const void* dy4, const float* scale_ptr, float* part, int splits, int N,
                                                                      int H, int W, int Ci, int Co, void* stream) {
    if (!x || !dy4 || !part) return fail(FV_ERR_ARG, "fv_outconv_wgrad: null pointer");
    if (!fold_shape_ok(N, H, W, Ci, Co, kFR, kFR))
        return fail(FV_ERR_UNSUPPORTED, "fv_outconv_wgrad: needs a 7x7 filter, Ci = 32, Co <= 4, W = 128 or 256 (got Ci=%d Co=%d W=%d)", Ci, Co, W);
    FoldWgradParams p{};
    p.N = N; p.H = H; p.W = W; p.halves = W / 128; p.Co = Co;
    p.wins_per_col = (H + kWT - 1) / kWT;
    const int grid = fold_wgrad_grid(N, H, W, p.units_total, p.units_per_cta);
    if (grid != splits) return fail(FV_ERR_ARG, "fv_outconv_wgrad: splits=%d does not match fv_outconv_wgrad_splits() = %d", splits, grid);
    p.rec_off = kWSlots * kWSlab;
    p.bar_off = p.rec_off + kRecSlots * kWSlab;
    p.dy4 = (const __nv_bfloat16*)dy4; p.dw = part; p.slab_stride = (long long)Co * kFC * kFR * kFR; p.scale_ptr = scale_ptr; p.trace = trace_ptr();
    { const char* v = getenv("FV_FOLD_DEBUG"); p.dbg = v ? atoi(v) : 0; }
    CUtensorMap tmX;
    {
        uint64_t dims[4] = {(uint64_t)kFC, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)kRowB, (uint64_t)W * kRowB, (uint64_t)H * W * kRowB};
        uint32_t box[4] = {(uint32_t)kFC, 128, 1, 1};
        if (int e = encode_tmap_bf16(&tmX, x, 4, dims, str, box, kRowB)) return e;
    }
    const size_t smem = (size_t)p.bar_off + (2 * kWSlots + 2 * kRecSlots + 1) * 8 + 16 + 1024 + 64;
    static bool attr_set = false;
    if (!attr_set) {
        FV_CUDA(cudaFuncSetAttribute(fold_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    fold_wgrad_kernel<<<grid, 352, smem, (cudaStream_t)stream>>>(tmX, p);
    FV_LAUNCH_CHECK("fold_wgrad_kernel");
    return FV_OK;
}
