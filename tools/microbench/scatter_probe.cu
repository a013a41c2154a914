// scatter_probe.cu — B200 microbenchmarks behind the DCNv2 gather/scatter design:
//   1. shared-memory int32 atomics (ATOMS.ADD): conflict-free vs random banks
//   2. global red.v4.f32 scatter with DCN-like locality: one lane per 16 B vs lane pairs covering one 32-B sector
//   3. LDS.128 gather of 2x2x8ch neighbourhoods from a staged box: natural order vs bank-rotated order
//   4. LDG.256 gather from the group-blocked layout: one corner per lane vs lane pairs on x-adjacent corners
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scatter_probe scatter_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int BOX_H = 22, BOX_W = 30, PITCH = BOX_W * 8;   // floats; 960 B pitch = 64 mod 128
constexpr int BOX_WORDS = BOX_H * PITCH;

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// ---- 1. shared int atomics: mode 0 conflict-free (lane L -> bank (L+i)%32 at a random row), 1 random word, 2 random pixel + lane-rotated channel
template <int MODE>
__global__ void __launch_bounds__(384, 2) k_atoms(int *out, int iters)
{
    __shared__ int box[BOX_WORDS];
    for (int i = threadIdx.x; i < BOX_WORDS; i += blockDim.x) box[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t h = hash32(threadIdx.x * 9781u + blockIdx.x * 77u + 1u);
    for (int it = 0; it < iters; ++it) {
        h = hash32(h);
        const int y = h % (BOX_H - 1), x = (h >> 8) % (BOX_W - 1);
        const int base = y * PITCH + x * 8;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            int w;
            if (MODE == 0) {         // rotated: word index inside the 2x2x8 neighbourhood chosen so that bank = (lane + i) % 32
                const int bank0 = base & 31;
                const int j = (lane + i - bank0) & 31;       // j-th bank after base; neighbourhood covers 32 consecutive banks
                w = base + (j >> 4) * PITCH - ((j >> 4) * (PITCH & 31)) + (j & 15) + ((j >> 4) * 16) - (j >> 4) * 16;
                // row 1 starts at base + PITCH whose bank is bank0 + 16: word j>=16 lives at base + PITCH + (j - 16)
                w = (j < 16) ? base + j : base + PITCH + (j - 16);
            } else if (MODE == 1) {  // natural order: corner i>>3, channel i&7 (all lanes same channel -> same bank inside an octet)
                const int k = i >> 3, c = i & 7;
                w = base + (k >> 1) * PITCH + (k & 1) * 8 + c;
            } else {                 // natural corner order, lane-rotated channel
                const int k = i >> 3, c = (i + lane) & 7;
                w = base + (k >> 1) * PITCH + (k & 1) * 8 + c;
            }
            atomicAdd(&box[w], it + i);
        }
    }
    __syncthreads();
    int s = 0;
    for (int i = threadIdx.x; i < BOX_WORDS; i += blockDim.x) s += box[i];
    if (s == 0x7fffffff) out[0] = s;
}

// ---- 2. global red.v4 scatter: image 256x256, 8 groups, blocked [g][y][x][8]; thread = (pixel, tap) like the DCN backward
__device__ __forceinline__ void red_v4(float *a, float x, float y, float z, float w)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float gauss(uint32_t &h)
{
    float s = 0.f;
    for (int i = 0; i < 4; ++i) { h = hash32(h); s += (h & 0xffff) * (1.f / 65536.f); }
    return (s - 2.f) * 1.732f;     // ~N(0,1)
}
template <int PAIR>
__global__ void __launch_bounds__(384, 2) k_red(float *gin, int H, int W, int ntile_x, float sigma)
{
    const int p = threadIdx.x & 127, r = threadIdx.x >> 7, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, g = blockIdx.y;
    const int ty0 = (tile / ntile_x) * 8, tx0 = (tile % ntile_x) * 16;
    const int ho = ty0 + (p >> 4), wo = tx0 + (p & 15);
    float *gb = gin + (size_t)g * H * W * 8;
    uint32_t h = hash32((tile * 8 + g) * 384 + threadIdx.x + 12345u);
    for (int s = 0; s < 3; ++s) {
        const float y = ho + (r - 1) + sigma * gauss(h), x = wo + (s - 1) + sigma * gauss(h);
        int y0 = (int)floorf(y), x0 = (int)floorf(x);
        y0 = min(max(y0, 0), H - 2); x0 = min(max(x0, 0), W - 2);
        const float v = 1.f;
        if (PAIR == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float *a = gb + ((size_t)(y0 + (k >> 1)) * W + x0 + (k & 1)) * 8;
                red_v4(a, v, v, v, v); red_v4(a + 4, v, v, v, v);
            }
        } else {
            // lanes 2k / 2k+1 cover halves 0 / 1 of the same pixel: both samples of the pair, 4 corners each
            const int py0 = __shfl_xor_sync(0xffffffffu, y0, 1), px0 = __shfl_xor_sync(0xffffffffu, x0, 1);
            const int half = lane & 1;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int yy = (q == half) ? y0 : py0, xx = (q == half) ? x0 : px0;    // sample q of the pair
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float *a = gb + ((size_t)(yy + (k >> 1)) * W + xx + (k & 1)) * 8 + half * 4;
                    red_v4(a, v, v, v, v);
                }
            }
        }
    }
}

// ---- 3. LDS.128 gather from a box: natural vs rotated
template <int ROT>
__global__ void __launch_bounds__(384, 2) k_lds(float *out, int iters)
{
    __shared__ __align__(128) float box[BOX_WORDS];
    for (int i = threadIdx.x; i < BOX_WORDS; i += blockDim.x) box[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t h = hash32(threadIdx.x * 9781u + blockIdx.x * 77u + 1u);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int it = 0; it < iters; ++it) {
        h = hash32(h);
        const int y = h % (BOX_H - 1), x = (h >> 8) % (BOX_W - 1);
        const int base = y * PITCH + x * 8;            // floats
        if (ROT == 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = *reinterpret_cast<const float4 *>(box + base + (c >> 2) * PITCH + (c & 3) * 4);
                acc[(c & 1) * 4 + 0] += v.x; acc[(c & 1) * 4 + 1] += v.y; acc[(c & 1) * 4 + 2] += v.z; acc[(c & 1) * 4 + 3] += v.w;
            }
        } else {
            const int b = (base >> 2) & 7;             // bank group of chunk 0
            const int c0 = ((lane & 7) - b) & 7;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = (c0 + i) & 7;
                const float4 v = *reinterpret_cast<const float4 *>(box + base + (c >> 2) * PITCH + (c & 3) * 4);
                acc[(i & 1) * 4 + 0] += v.x; acc[(i & 1) * 4 + 1] += v.y; acc[(i & 1) * 4 + 2] += v.z; acc[(i & 1) * 4 + 3] += v.w;
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 1.2345f) out[0] = s;
}

// ---- 4. LDG.256 gather from the blocked layout, DCN-like addresses
struct f8 { float v[8]; };
__device__ __forceinline__ f8 ldg8(const float *p)
{
    f8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
template <int PAIR>
__global__ void __launch_bounds__(384, 2) k_ldg(const float *in, float *out, int H, int W, int ntile_x, float sigma)
{
    const int p = threadIdx.x & 127, r = threadIdx.x >> 7, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int ty0 = (tile / ntile_x) * 8, tx0 = (tile % ntile_x) * 16;
    const int ho = ty0 + (p >> 4), wo = tx0 + (p & 15);
    float acc = 0.f;
    for (int g = 0; g < 8; ++g) {
        const float *gb = in + (size_t)g * H * W * 8;
        uint32_t h = hash32((tile * 8 + g) * 384 + threadIdx.x + 12345u);
        for (int s = 0; s < 3; ++s) {
            const float y = ho + (r - 1) + sigma * gauss(h), x = wo + (s - 1) + sigma * gauss(h);
            int y0 = (int)floorf(y), x0 = (int)floorf(x);
            y0 = min(max(y0, 0), H - 2); x0 = min(max(x0, 0), W - 2);
            if (PAIR == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const f8 v = ldg8(gb + ((size_t)(y0 + (k >> 1)) * W + x0 + (k & 1)) * 8);
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc += v.v[c];
                }
            } else {
                const int py0 = __shfl_xor_sync(0xffffffffu, y0, 1), px0 = __shfl_xor_sync(0xffffffffu, x0, 1);
                const int dx = lane & 1;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int yy = (q == dx) ? y0 : py0, xx = (q == dx) ? x0 : px0;
#pragma unroll
                    for (int dy = 0; dy < 2; ++dy) {
                        const f8 v = ldg8(gb + ((size_t)(yy + dy) * W + xx + dx) * 8);
#pragma unroll
                        for (int c = 0; c < 8; ++c) acc += v.v[c];
                    }
                }
            }
        }
    }
    if (acc == 1.2345f) out[0] = acc;
}

template <typename F> float time_ms(F f, int rep = 5)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < rep; ++i) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = ms < best ? ms : best;
    }
    CK(cudaGetLastError());
    return best;
}

int main()
{
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int nsm = pr.multiProcessorCount; const double ghz = pr.clockRate * 1e-6;
    printf("device %s, %d SMs, %.3f GHz\n", pr.name, nsm, ghz);
    int *dout; CK(cudaMalloc(&dout, 1024));
    const int H = 256, W = 256;
    float *img; CK(cudaMalloc(&img, (size_t)8 * H * W * 8 * 4)); CK(cudaMemset(img, 0, (size_t)8 * H * W * 8 * 4));

    // 1. ATOMS: grid = 2 CTAs/SM, 384 threads, iters samples per thread, 32 atomics per sample
    const int iters = 200;
    {
        const double nwarp_instr = (double)iters * 32 * 12 * 2;   // warp-level atomic instructions per SM
        float t0 = time_ms([&] { k_atoms<0><<<2 * nsm, 384>>>(dout, iters); });
        float t1 = time_ms([&] { k_atoms<1><<<2 * nsm, 384>>>(dout, iters); });
        float t2 = time_ms([&] { k_atoms<2><<<2 * nsm, 384>>>(dout, iters); });
        printf("ATOMS.ADD s32 clk per warp instruction per SM: rotated(conflict-free) %.2f | natural(same channel) %.2f | lane-rotated channel %.2f\n",
               t0 * 1e-3 * ghz * 1e9 / nwarp_instr, t1 * 1e-3 * ghz * 1e9 / nwarp_instr, t2 * 1e-3 * ghz * 1e9 / nwarp_instr);
        printf("   -> clk per sample (32 atomics x 32 lanes): %.2f | %.2f | %.2f\n", t0 * 1e-3 * ghz * 1e9 / nwarp_instr, t1 * 1e-3 * ghz * 1e9 / nwarp_instr, t2 * 1e-3 * ghz * 1e9 / nwarp_instr);
    }
    // 2. global RED: 512 tiles x 8 groups x 384 threads x 3 taps = 4.72 M samples
    for (float sigma : {0.f, 0.5f, 2.f}) {
        float t0 = time_ms([&] { k_red<0><<<dim3(512, 8), 384>>>(img, H, W, 16, sigma); });
        float t1 = time_ms([&] { k_red<1><<<dim3(512, 8), 384>>>(img, H, W, 16, sigma); });
        printf("global red.v4.f32 scatter, sigma %.1f: one lane per corner %.1f us | lane pairs per 32-B sector %.1f us (4.72 M samples, 37.7 M REDs)\n", sigma, t0 * 1e3, t1 * 1e3);
    }
    // 3. LDS gather
    {
        const double nsamp = (double)iters * 10 * 384 * 2;     // samples per SM
        float t0 = time_ms([&] { k_lds<0><<<2 * nsm, 384>>>((float *)dout, iters * 10); });
        float t1 = time_ms([&] { k_lds<1><<<2 * nsm, 384>>>((float *)dout, iters * 10); });
        printf("LDS.128 2x2x8ch gather clk per sample per SM: natural %.2f | rotated %.2f\n", t0 * 1e-3 * ghz * 1e9 / nsamp, t1 * 1e-3 * ghz * 1e9 / nsamp);
    }
    // 4. LDG gather
    for (float sigma : {0.f, 0.5f, 2.f}) {
        float t0 = time_ms([&] { k_ldg<0><<<512, 384>>>(img, (float *)dout, H, W, 16, sigma); });
        float t1 = time_ms([&] { k_ldg<1><<<512, 384>>>(img, (float *)dout, H, W, 16, sigma); });
        printf("LDG.256 gather, sigma %.1f: one corner per lane %.1f us | lane pairs on x-adjacent corners %.1f us (4.72 M samples)\n", sigma, t0 * 1e3, t1 * 1e3);
    }
    return 0;
}
