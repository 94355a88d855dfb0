// Host-side helpers shared by the C-ABI entry points: error state, CUDA checks, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fv {

// Error convention of the C ABI (include/facevae_b200.h): 0 = success, non-zero = failure with a
// message retrievable through fv_last_error().  No exception crosses the boundary.
enum : int { FV_OK = 0, FV_ERR_ARG = 1, FV_ERR_UNSUPPORTED = 2, FV_ERR_CUDA = 3, FV_ERR_INTERNAL = 4 };

int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define FV_CUDA(call)                                          \
    do {                                                       \
        cudaError_t _e = (call);                               \
        if (_e != cudaSuccess) return fv::cuda_fail(_e, #call); \
    } while (0)
#define FV_LAUNCH_CHECK(name)                                   \
    do {                                                        \
        cudaError_t _e = cudaGetLastError();                    \
        if (_e != cudaSuccess) return fv::cuda_fail(_e, name);  \
    } while (0)

int num_sms();
long long* trace_ptr();   // fv_debug.cu: device counters for FV_TRACE builds (null otherwise)
// rank <= 5; dims innermost first; strides_bytes[i] = byte stride of dim i+1; box per dim; swizzle 0/32/64/128.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes);

}  // namespace fv
