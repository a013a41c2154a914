// common.cu — library-level entry points and error plumbing of libebfi_b200.so.
#include "common.cuh"

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
