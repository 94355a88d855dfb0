// Calibration micro-benchmark (not on the product path): issue rate / latency of tcgen05.mma (M = 128, K = 16, bf16)
// as a function of N and of the shared-memory row width, with operands resident in shared memory and no TMA or epilogue.
#include "../../include/facevae_b200.h"
#include "fv_host.h"
#include "fv_ptx.cuh"

namespace fv {
static long long* g_trace_dev = nullptr;      // device buffer of 148*8 counters registered by the caller (debug builds)
long long* trace_ptr() { return g_trace_dev; }

__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(int n_cols, int row_bytes, int iters, int a_distinct, int mn_major, long long* out_cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_slot, 256);
        tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 1) {
        const bool leader = elect_one_sync();
        const uint32_t idesc = umma_idesc_bf16(128, n_cols, mn_major, mn_major);
        const uint32_t layout = umma_layout_code(row_bytes);
        const uint64_t tmpl = mn_major ? umma_smem_desc(0, (uint32_t)(64 * row_bytes), 8u * row_bytes, layout)
                                       : umma_smem_desc(0, 16, 8u * row_bytes, layout);
        const uint32_t hi = (uint32_t)(tmpl >> 32), lo0 = (uint32_t)tmpl;
        const uint32_t a_base = smem_u32(smem) >> 4, b_base = (smem_u32(smem) + 48 * 1024) >> 4;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            // a_distinct >= 16: additionally shift the A start address by (i % 3) pixel rows, as the slab schedule does
            const uint32_t shift = a_distinct >= 16 ? (uint32_t)((i % 3) * (row_bytes >> 4)) : 0u;
            const uint32_t a_lo = lo0 | (a_base + shift + (a_distinct ? ((i & 3) * 2) : 0));
            const uint32_t b_lo = lo0 | (b_base + (a_distinct ? ((i & 3) * 2) : 0));
            if (leader) tc_mma_f16_lohi(tmem_base, a_lo, b_lo, hi, idesc, i > 0);
        }
        if (leader) tc_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (leader && blockIdx.x == 0) out_cycles[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace fv

extern "C" __attribute__((visibility("default"))) int fv_debug_mma_rate(int n_cols, int row_bytes, int iters, int a_distinct, int mn_major,
                                                                      int all_sms, long long* out_cycles_dev, void* stream) {
    using namespace fv;
    static bool attr_set = false;
    if (!attr_set) {
        FV_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    mma_rate_kernel<<<all_sms ? num_sms() : 1, 128, 98 * 1024, (cudaStream_t)stream>>>(n_cols, row_bytes, iters, a_distinct, mn_major, out_cycles_dev);
    FV_LAUNCH_CHECK("mma_rate_kernel");
    return FV_OK;
}

extern "C" __attribute__((visibility("default"))) int fv_debug_trace_set(long long* dev_counters) {
#ifdef FV_TRACE
    fv::g_trace_dev = dev_counters;
    return 0;
#else
    (void)dev_counters;
    return fv::fail(fv::FV_ERR_UNSUPPORTED, "library built without FV_TRACE");
#endif
}
