// common.cuh — shared helpers of the ebfi_b200 kernels (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/ebfi_b200.h"

namespace ebfi {

// Thread-local message behind ebfi_last_error().
char *last_error_buf();
int fail(int code, const char *fmt, ...);

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T> __host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T> __host__ __device__ constexpr T round_up(T a, T b) { return ceil_div(a, b) * b; }

int sm_count();      // multiprocessors of the current device (cached per device)

}  // namespace ebfi

// Kernel-launch / runtime-call check used by every entry point.
#define EBFI_CUDA_OK(expr)                                                              \
    do {                                                                                \
        cudaError_t e_ = (expr);                                                        \
        if (e_ != cudaSuccess)                                                          \
            return ebfi::fail(EBFI_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                              __FILE__, __LINE__);                                      \
    } while (0)

#define EBFI_LAUNCH_OK(what)                                                            \
    do {                                                                                \
        cudaError_t e_ = cudaGetLastError();                                            \
        if (e_ != cudaSuccess)                                                          \
            return ebfi::fail(EBFI_ERR_CUDA, "launch of %s failed: %s", what,           \
                              cudaGetErrorString(e_));                                  \
    } while (0)

#define EBFI_REQUIRE(cond, ...)                                                         \
    do {                                                                                \
        if (!(cond)) return ebfi::fail(EBFI_ERR_INVALID, __VA_ARGS__);                  \
    } while (0)
