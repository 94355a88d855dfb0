// Error state, device queries and TMA descriptor encoding for the C ABI.
#include "fv_host.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/facevae_b200.h"

namespace fv {

static thread_local char g_err[1024] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    (void)cudaGetLastError();      // reported here: do not leave it behind for the next entry point's launch check
    return fail(FV_ERR_CUDA, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
}

int num_sms() {
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        // resolved through the runtime so the library carries no link-time dependency on libcuda
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(FV_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(FV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u sw %d",
                    (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                    (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0), bdim[0],
                    rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0, rank > 3 ? bdim[3] : 0, swizzle_bytes);
    return FV_OK;
}

}  // namespace fv

extern "C" __attribute__((visibility("default"))) const char* fv_last_error(void) { return fv::g_err; }
extern "C" __attribute__((visibility("default"))) const char* fv_version(void) { return "face_vae_b200 0.2 (sm_100a)"; }
// bumped whenever an exported signature changes: the ctypes binding refuses a library built from other sources
extern "C" __attribute__((visibility("default"))) int fv_abi_version(void) { return FV_ABI_VERSION; }
extern "C" __attribute__((visibility("default"))) int fv_device_ok(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess) return fv::fail(fv::FV_ERR_CUDA, "no CUDA device");
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return fv::fail(fv::FV_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (p.major != 10) return fv::fail(fv::FV_ERR_UNSUPPORTED, "device %s is sm_%d%d; this library is sm_100a only", p.name, p.major, p.minor);
    return 0;
}
