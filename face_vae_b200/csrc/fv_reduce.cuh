// Deterministic cross-block reduction of a per-block vector: the replacement for "one atomicAdd per channel and block".
//
// fp32 atomics add in arrival order, so two runs of the same step differed in the last bits of every batch-norm sum; bf16
// rounding and ReLU masks amplify that to percent-level differences of the encoder gradients (round-1 VERDICT, weak #1).
// Here every block stores its partial vector, the LAST block of each group of 64 blocks (ticket counter) adds the group's
// rows in block order, and the last group to finish adds the group rows in group order: a fixed summation tree for a
// given grid, i.e. bitwise-reproducible results, at the price of ~1-2 us at the tail of the kernel.  The block that ends up
// with the totals continues with whatever follows (statistic finalize, cross-rank exchange), which is what used to be a
// separate single-block launch.
//
// Workspace (caller-provided, `fv_reduce_ws_bytes`): [ticket words | group rows | block rows].  The ticket words must be
// zero at launch and are zero again when the kernel exits, so one zero-initialised buffer per stream serves every launch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fv {

static constexpr int kRedGroup = 64;          // blocks per first-level group (grids of <= 64 blocks: one level)
static constexpr int kRedMaxGroups = 32;      // => grids of up to 2048 blocks
static constexpr int kRedTicketBytes = 512;   // word 0: finished groups; word 1 + g: finished blocks of group g

__host__ __device__ inline size_t det_reduce_ws_bytes(int max_blocks, int n, int elem_size) {
    return (size_t)kRedTicketBytes + ((size_t)kRedMaxGroups + (size_t)max_blocks) * (size_t)n * (size_t)elem_size;
}

struct BlockSync {            // all threads of the block take part
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct NamedSync {            // a warp-aligned subset of the block (e.g. the four epilogue warps of the conv kernels)
    int id, nthreads;
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
};

template <typename T>
__device__ __forceinline__ T ld_l2(const T* p) { return __ldcg(p); }

// `my`  : this block's partial vector (shared memory, n elements, complete and visible: the caller synchronised)
// `tot` : shared memory, n elements; receives the grid totals in the ONE block for which the function returns true
// `flag`: one shared-memory int of scratch
// nb / b: number of participating blocks and this block's index among them; tid / nthr: the participating threads
template <typename T, typename Sync>
__device__ __forceinline__ bool det_reduce(void* ws, int n, int nb, int b, const T* my, T* tot, int tid, int nthr, Sync sync, int* flag) {
    if (nb == 1) {
        for (int c = tid; c < n; c += nthr) tot[c] = my[c];
        sync();
        return true;
    }
    unsigned* tickets = reinterpret_cast<unsigned*>(ws);
    T* part2 = reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + kRedTicketBytes);
    T* part = part2 + (size_t)kRedMaxGroups * n;
    for (int c = tid; c < n; c += nthr) part[(size_t)b * n + c] = my[c];
    // release / acquire through ONE thread: the barrier orders every thread's stores before thread 0's gpu-scope fence
    // (fences are cumulative), and thread 0's fence after the ticket orders the other blocks' stores before the barrier
    // that lets this block's threads read them (with L2 loads) -- a fence per thread costs ~1 us per level
    sync();
    const int grp = b / kRedGroup, g0 = grp * kRedGroup;
    const int gsz = nb - g0 < kRedGroup ? nb - g0 : kRedGroup;
    const int ngroups = (nb + kRedGroup - 1) / kRedGroup;
    if (tid == 0) {
        __threadfence();
        *flag = (atomicAdd(&tickets[1 + grp], 1u) == (unsigned)(gsz - 1));
        __threadfence();
    }
    sync();
    if (!*flag) return false;
    for (int c = tid; c < n; c += nthr) {
        T a = T(0);
        for (int j0 = 0; j0 < kRedGroup; j0 += 32) {           // 32 independent L2 loads in flight, added in block order
            if (j0 >= gsz) break;
            T v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = j0 + j < gsz ? ld_l2(part + (size_t)(g0 + j0 + j) * n + c) : T(0);
#pragma unroll
            for (int j = 0; j < 32; ++j) a += v[j];
        }
        if (ngroups == 1) tot[c] = a;
        else part2[(size_t)grp * n + c] = a;
    }
    if (ngroups == 1) {
        sync();
        if (tid == 0) tickets[1] = 0u;
        return true;
    }
    sync();
    if (tid == 0) {
        __threadfence();
        *flag = (atomicAdd(&tickets[0], 1u) == (unsigned)(ngroups - 1));
        __threadfence();
    }
    sync();
    if (!*flag) return false;
    for (int c = tid; c < n; c += nthr) {
        T a = T(0);
        for (int g = 0; g < ngroups; g += 16) {
            T v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = g + j < ngroups ? ld_l2(part2 + (size_t)(g + j) * n + c) : T(0);
#pragma unroll
            for (int j = 0; j < 16; ++j) a += v[j];
        }
        tot[c] = a;
    }
    sync();
    for (int i = tid; i <= ngroups; i += nthr) tickets[i] = 0u;
    return true;
}

}  // namespace fv
