// ffma2_probe.cu — issue rate of packed fp32x2 FMA (FFMA2, sm_100) against scalar FFMA, per SM sub-partition.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE> __global__ void k(float *out, float s, int iters, long long *clk)
{
    float a[16]; u64 p[8];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < 8; ++i) p[i] = pack(a[2 * i], a[2 * i + 1]);
    const u64 s2 = pack(s, s);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 0.5f);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], s2, s2);
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < 16; ++i) acc += a[i];
    for (int i = 0; i < 8; ++i) acc += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main()
{
    float *o; long long *c, h;
    cudaMalloc(&o, 1 << 24); cudaMalloc(&c, 8);
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int iters = 4096;
        k<0><<<148, warps * 32>>>(o, 0.999f, iters, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        const double f1 = (double)h / iters;
        k<1><<<148, warps * 32>>>(o, 0.999f, iters, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        const double f2 = (double)h / iters;
        printf("%2d warps/SM: 16 FFMA per thread-iter %.1f clk | 8 FFMA2 (same flops) %.1f clk -> fp32 FMA lanes/clk/SM: %.0f vs %.0f\n",
               warps, f1, f2, 16.0 * warps * 32 / f1, 16.0 * warps * 32 / f2);
    }
    return 0;
}
