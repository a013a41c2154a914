// common.cu — library-level entry points and error plumbing of libebfi_b200.so.
#include "common.cuh"
#include "tma.cuh"

#include <cstring>

namespace ebfi {

char *last_error_buf()
{
    static thread_local char buf[512] = "";
    return buf;
}

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int sm_count()
{
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace ebfi

namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int encode_3d(CUtensorMap &tm, const void *base, CUtensorMapDataType dtype, const uint64_t (&dims)[3],
              const uint64_t (&strides)[2], const uint32_t (&box)[3])
{
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return ebfi::fail(EBFI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[3] = {dims[0], dims[1], dims[2]};
    const cuuint64_t gstr[2] = {strides[0], strides[1]};
    const cuuint32_t bx[3] = {box[0], box[1], box[2]};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&tm, dtype, 3, const_cast<void *>(base), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ebfi::fail(EBFI_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return EBFI_OK;
}

}  // namespace tma

extern "C" {

int ebfi_abi_version(void) { return EBFI_ABI_VERSION; }

const char *ebfi_last_error(void) { return ebfi::last_error_buf(); }

int ebfi_device_arch(void)
{
    int dev = 0, major = 0, minor = 0;
    EBFI_CUDA_OK(cudaGetDevice(&dev));
    EBFI_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    EBFI_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    return major * 10 + minor;
}

}  // extern "C"
