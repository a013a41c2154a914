// tcgemm_selftest.cu — one-CTA 3xTF32 GEMM on the tcgen05 path, C[MxN] = A[MxK] * B[NxK]^T.
//
// Test-only (libebfi_b200_selftest.so, include/ebfi_b200_selftest.h): it exists to pin the plumbing of umma.cuh (shared-memory descriptors for
// K-major and MN-major operands in the no-swizzle core-matrix layout, instruction descriptor,
// TMEM allocation / tcgen05.ld lane mapping for M = 128 and M = 64, commit -> mbarrier) against
// an fp64 reference, independently of the DCN kernels that are built on the same pieces
// (tests/test_tcgen05_gpu.py).
#include "common.cuh"
#include "../../include/ebfi_b200_selftest.h"
#include "umma.cuh"

namespace {

constexpr int KS = 72;            // K elements per smem stage (= one DCN deformable group: 8 ch x 9 taps)
constexpr int KSC = KS / 4;       // 16-byte chunks per row per stage
constexpr int MAXN = 128;

// Stage image sizes (bytes) for 128 rows x KS columns of fp32
constexpr int OPER_BYTES = 128 * KS * 4;

__global__ void __launch_bounds__(128)
tc_gemm_selftest_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C,
                        int M, int N, int K, int a_mn)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float *a_hi = reinterpret_cast<float *>(smem);
    float *a_lo = reinterpret_cast<float *>(smem + OPER_BYTES);
    float *b_hi = reinterpret_cast<float *>(smem + 2 * OPER_BYTES);
    float *b_lo = reinterpret_cast<float *>(smem + 3 * OPER_BYTES);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    if (warp == 0) umma::tmem_alloc<512>(&tmem_slot);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = umma::instr_desc_tf32(M, N, a_mn, 0);

    // The tensor core's fp32 accumulate rounds toward zero, so the error of a long dependent
    // chain of MMAs into ONE accumulator grows linearly with its length. Spread the chain:
    // hi*hi k-steps go round-robin into `nacc` accumulators, both small cross terms into one
    // more; the epilogue sums them in fp32 registers. Accumulator j sits at column j*N.
    const int nacc = min(4, 512 / N - 1);
    uint32_t parity = 0;
    int step = 0;                       // global k-step counter (thread 0 only)
    for (int k0 = 0; k0 < K; k0 += KS) {
        const int kl = min(KS, K - k0);
        // ---- fill the stage (generic proxy), splitting into hi / lo
        if (!a_mn) {    // A K-major: [row/8][chunk][8][4]
            for (int e = tid; e < M * kl; e += blockDim.x) {
                const int r = e / kl, k = e % kl;
                float hi, lo;
                umma::split_tf32(A[(size_t)r * K + k0 + k], hi, lo);
                const int off = (r / 8) * (KSC * 32) + (k / 4) * 32 + (r % 8) * 4 + (k % 4);
                a_hi[off] = hi; a_lo[off] = lo;
            }
        } else {        // A MN-major (memory [K][M]): [k/8][m/4][8 k][4 m]
            for (int e = tid; e < M * kl; e += blockDim.x) {
                const int k = e / M, m = e % M;
                float hi, lo;
                umma::split_tf32(A[(size_t)(k0 + k) * M + m], hi, lo);
                const int off = (k / 8) * (32 * 32) + (m / 4) * 32 + (k % 8) * 4 + (m % 4);
                a_hi[off] = hi; a_lo[off] = lo;
            }
        }
        for (int e = tid; e < N * kl; e += blockDim.x) {
            const int r = e / kl, k = e % kl;
            float hi, lo;
            umma::split_tf32(B[(size_t)r * K + k0 + k], hi, lo);
            const int off = (r / 8) * (KSC * 32) + (k / 4) * 32 + (r % 8) * 4 + (k % 4);
            b_hi[off] = hi; b_lo[off] = lo;
        }
        umma::fence_smem_to_async();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            for (int ks = 0; ks < kl / 8; ++ks) {
                const uint32_t a_off = a_mn ? ks * 4096 : ks * 256, b_off = ks * 256;
                const uint32_t a_lbo = a_mn ? 4096 : 128, a_sbo = a_mn ? 128 : KSC * 128;
                const uint64_t dah = umma::smem_desc(umma::smem_u32(a_hi) + a_off, a_lbo, a_sbo);
                const uint64_t dal = umma::smem_desc(umma::smem_u32(a_lo) + a_off, a_lbo, a_sbo);
                const uint64_t dbh = umma::smem_desc(umma::smem_u32(b_hi) + b_off, 128, KSC * 128);
                const uint64_t dbl = umma::smem_desc(umma::smem_u32(b_lo) + b_off, 128, KSC * 128);
                const uint32_t d_x = tmem + nacc * N, d_h = tmem + (step % nacc) * N;
                umma::mma_tf32(d_x, dal, dbh, idesc, step > 0);
                umma::mma_tf32(d_x, dah, dbl, idesc, true);
                umma::mma_tf32(d_h, dah, dbh, idesc, step >= nacc);
                ++step;
            }
            umma::commit(&bar);
        }
        umma::mbar_wait(&bar, parity);      // MMAs of this stage done: smem reusable, accumulator valid
        parity ^= 1;
        umma::fence_after_sync();
    }

    // ---- epilogue: warp w owns TMEM lanes 32w..32w+31
    // M = 128: row = lane index. M = 64: rows 16q..16q+15 live in lanes 32q..32q+15.
    const int row = (M == 128) ? tid : (lane < 16 ? warp * 16 + lane : -1);
    const int nused = min(nacc, K / 8);         // hi*hi accumulators that received at least one MMA
    for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8], u[8];
        umma::tmem_ld8(umma::tmem_addr(tmem, warp * 32, nacc * N + c0), v);
        umma::tmem_ld_wait();
        for (int j = 0; j < nused; ++j) {
            umma::tmem_ld8(umma::tmem_addr(tmem, warp * 32, j * N + c0), u);
            umma::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += u[i];
        }
        if (row >= 0 && row < M)
#pragma unroll
            for (int j = 0; j < 8; ++j) C[(size_t)row * N + c0 + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

}  // namespace

extern "C" int ebfi_selftest_gemm_tf32x3(void *stream, const float *A, const float *B, float *C, int M, int N,
                                         int K, int a_mn_major)
{
    EBFI_REQUIRE(A && B && C, "selftest_gemm: null pointer");
    EBFI_REQUIRE(M == 128 || M == 64, "selftest_gemm: M must be 64 or 128");
    EBFI_REQUIRE(N >= 8 && N <= MAXN && N % (M == 128 ? 16 : 8) == 0, "selftest_gemm: bad N=%d for M=%d", N, M);
    EBFI_REQUIRE(K > 0 && K % 8 == 0, "selftest_gemm: K must be a positive multiple of 8");
    EBFI_REQUIRE(!a_mn_major || M == 128, "selftest_gemm: MN-major A is exercised with M=128 only");
    const int smem = 4 * OPER_BYTES;
    EBFI_CUDA_OK(cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_gemm_selftest_kernel<<<1, 128, smem, ebfi::as_stream(stream)>>>(A, B, C, M, N, K, a_mn_major);
    EBFI_LAUNCH_OK("tc_gemm_selftest_kernel");
    return EBFI_OK;
}

// ---- bf16x3 variant ---------------------------------------------------------------------
// Same GEMM with bf16 hi/lo pairs and kind::f16 MMAs (K = 16 per instruction); B may use a padded
// LBO (144 B between K chunks), which the DCN backward uses to keep its 2-byte operand stores free
// of bank conflicts.
namespace {
__global__ void __launch_bounds__(128)
tc_gemm_bf16_selftest_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C,
                             int M, int N, int K, int b_lbo, int a_mn)
{
    extern __shared__ __align__(128) unsigned char smem[];
    // A: [M/8][K/8 chunks][8][8] bf16, LBO 128; B: [N/8][K/8 chunks at b_lbo][8 rows x 16 B]
    const int kch = K / 8;
    const int a_part = 128 * K * 2, b_sbo = kch * b_lbo, b_part = (128 / 8) * b_sbo;
    unsigned short *a_hi = reinterpret_cast<unsigned short *>(smem), *a_lo = reinterpret_cast<unsigned short *>(smem + a_part);
    unsigned char *b_hi = smem + 2 * a_part, *b_lo = smem + 2 * a_part + b_part;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    if (warp == 0) umma::tmem_alloc<256>(&tmem_slot);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    for (int e = tid; e < M * K; e += blockDim.x) {
        const int r = e / K, k = e % K;
        unsigned short hi, lo;
        umma::split_bf16(A[(size_t)r * K + k], hi, lo);
        // K-major: 16-byte chunks run along K. MN-major (a_mn): the same 128-byte core matrices hold 8 k-rows of
        // 8 consecutive m: addr = (m/8)*SBO + (k/8)*LBO + (k%8)*16 + (m%8)*2 with SBO = 128, LBO = 16 * 128
        const int off = a_mn ? (r / 8) * 64 + (k / 8) * (16 * 64) + (k % 8) * 8 + (r % 8)
                             : (r / 8) * (kch * 64) + (k / 8) * 64 + (r % 8) * 8 + (k % 8);
        a_hi[off] = hi; a_lo[off] = lo;
    }
    for (int e = tid; e < N * K; e += blockDim.x) {
        const int r = e / K, k = e % K;
        unsigned short hi, lo;
        umma::split_bf16(B[(size_t)r * K + k], hi, lo);
        const int off = (r / 8) * b_sbo + (k / 8) * b_lbo + (r % 8) * 16 + (k % 8) * 2;
        *reinterpret_cast<unsigned short *>(b_hi + off) = hi;
        *reinterpret_cast<unsigned short *>(b_lo + off) = lo;
    }
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = umma::instr_desc_bf16(M, N) | ((uint32_t)a_mn << 15);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t ao = a_mn ? ks * 2 * (16 * 128) : ks * 256, bo = ks * 2 * b_lbo;
            const uint32_t a_lbo = a_mn ? 16 * 128 : 128, a_sbo = a_mn ? 128 : kch * 128;
            const uint64_t dah = umma::smem_desc(umma::smem_u32(a_hi) + ao, a_lbo, a_sbo);
            const uint64_t dal = umma::smem_desc(umma::smem_u32(a_lo) + ao, a_lbo, a_sbo);
            const uint64_t dbh = umma::smem_desc(umma::smem_u32(b_hi) + bo, b_lbo, b_sbo);
            const uint64_t dbl = umma::smem_desc(umma::smem_u32(b_lo) + bo, b_lbo, b_sbo);
            umma::mma_f16(tmem + N, dal, dbh, idesc, ks > 0);
            umma::mma_f16(tmem + N, dah, dbl, idesc, true);
            umma::mma_f16(tmem, dah, dbh, idesc, ks > 0);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    const int row = (M == 128) ? tid : (lane < 16 ? warp * 16 + lane : -1);
    for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8], u[8];
        umma::tmem_ld8(umma::tmem_addr(tmem, warp * 32, c0), v);
        umma::tmem_ld8(umma::tmem_addr(tmem, warp * 32, N + c0), u);
        umma::tmem_ld_wait();
        if (row >= 0 && row < M)
            for (int j = 0; j < 8; ++j) C[(size_t)row * N + c0 + j] = v[j] + u[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<256>(tmem);
}
}  // namespace

extern "C" int ebfi_selftest_gemm_bf16x3(void *stream, const float *A, const float *B, float *C, int M, int N,
                                         int K, int b_lbo_bytes)
{
    const int a_mn = (b_lbo_bytes >> 16) & 1;      // bit 16: A operand stored MN-major (test-only switch)
    b_lbo_bytes &= 0xFFFF;
    EBFI_REQUIRE(A && B && C, "selftest_gemm_bf16: null pointer");
    EBFI_REQUIRE(M == 128 || M == 64, "selftest_gemm_bf16: M must be 64 or 128");
    EBFI_REQUIRE(N >= 8 && N <= 128 && N % (M == 128 ? 16 : 8) == 0, "selftest_gemm_bf16: bad N");
    EBFI_REQUIRE(K > 0 && K % 16 == 0 && K <= 128, "selftest_gemm_bf16: K must be a multiple of 16, <= 128");
    EBFI_REQUIRE(b_lbo_bytes >= 128 && b_lbo_bytes % 16 == 0 && b_lbo_bytes <= 256, "selftest_gemm_bf16: bad LBO");
    const int smem = 2 * 128 * K * 2 + 2 * 16 * (K / 8) * b_lbo_bytes;
    EBFI_CUDA_OK(cudaFuncSetAttribute(tc_gemm_bf16_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_gemm_bf16_selftest_kernel<<<1, 128, smem, ebfi::as_stream(stream)>>>(A, B, C, M, N, K, b_lbo_bytes, a_mn);
    EBFI_LAUNCH_OK("tc_gemm_bf16_selftest_kernel");
    return EBFI_OK;
}

// ---- layout probe -----------------------------------------------------------------------
// Fills 64 KB of shared memory with "my own float index" (split into low 11 bits / high bits so
// both halves are exact TF32 numbers) and multiplies it, as operand A (128 x 8) under the given
// descriptor fields, with an 8x8 identity B. C[m][k] is then the float index the tensor core
// fetched for A(m, k): a direct read-out of how the hardware interprets LBO / SBO / major-ness.
namespace {
__global__ void __launch_bounds__(128)
umma_probe_kernel(float *__restrict__ C, uint32_t lbo, uint32_t sbo, int a_mn)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float *a_low = reinterpret_cast<float *>(smem);                 // 16384 floats
    float *a_high = reinterpret_cast<float *>(smem + 65536);
    float *b_id = reinterpret_cast<float *>(smem + 131072);         // [16][8] K-major, 2 core matrices x 2 chunks
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid / 32;
    if (warp == 0) umma::tmem_alloc<32>(&tmem_slot);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    for (int e = tid; e < 16384; e += blockDim.x) { a_low[e] = (float)(e % 2048); a_high[e] = (float)(e / 2048); }
    for (int e = tid; e < 16 * 8; e += blockDim.x) {
        const int n = e / 8, k = e % 8;
        b_id[(n / 8) * 64 + (k / 4) * 32 + (n % 8) * 4 + (k % 4)] = (n == k) ? 1.f : 0.f;   // SBO=256 B, LBO=128 B
    }
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = umma::instr_desc_tf32(128, 16, a_mn, 0);
        const uint64_t db = umma::smem_desc(umma::smem_u32(b_id), 128, 256);
        umma::mma_tf32(tmem, umma::smem_desc(umma::smem_u32(a_low), lbo, sbo), db, idesc, false);
        umma::mma_tf32(tmem + 16, umma::smem_desc(umma::smem_u32(a_high), lbo, sbo), db, idesc, false);
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    float lo[8], hi[8];
    umma::tmem_ld8(umma::tmem_addr(tmem, warp * 32, 0), lo);
    umma::tmem_ld8(umma::tmem_addr(tmem, warp * 32, 16), hi);
    umma::tmem_ld_wait();
    for (int j = 0; j < 8; ++j) C[tid * 8 + j] = hi[j] * 2048.f + lo[j];
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<32>(tmem);
}
}  // namespace

extern "C" int ebfi_selftest_umma_probe(void *stream, float *C, int lbo_bytes, int sbo_bytes, int a_mn_major)
{
    EBFI_REQUIRE(C != nullptr, "umma_probe: null pointer");
    const int smem = 131072 + 1024;
    EBFI_CUDA_OK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_probe_kernel<<<1, 128, smem, ebfi::as_stream(stream)>>>(C, (uint32_t)lbo_bytes, (uint32_t)sbo_bytes, a_mn_major);
    EBFI_LAUNCH_OK("umma_probe_kernel");
    return EBFI_OK;
}

// ---- MMA issue-interval probe -----------------------------------------------------------------------------------
// One CTA per SM issues `iters` back-to-back kind::f16 (bf16) MMAs of shape 128 x N x 16 on fixed shared-memory
// operands (contents irrelevant) and reports clock64 cycles per MMA: the steady-state rate of the tensor pipe for
// that shape, including its shared-memory operand reads. Used to decide whether small-N GEMMs (N = 80 in kpn.cu)
// are bound by operand traffic rather than by math.
namespace {
template <int COMMIT_EVERY>
__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(float *__restrict__ cycles_per_mma, int N, int iters, int a_sbo_bytes, int b_sbo_bytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (64 * 1024) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u;
    if (warp == 0) umma::tmem_alloc<256>(&tmem_slot);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_init(&bar2, 1); umma::mbar_fence_init(); }
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const bool leader = umma::elect_one();
        const uint32_t idesc = umma::instr_desc_bf16(128, N);
        const uint64_t da = umma::smem_desc(umma::smem_u32(smem), 2880, (uint32_t)a_sbo_bytes);          // A: like kpn.cu's halo view
        // B: K-major rows, 8-row groups b_sbo_bytes apart (256 = dense N x 16; 18432 = kpn.cu's 9*128-channel weight image,
        // in which consecutive MMAs also walk along K)
        const uint64_t db = umma::smem_desc(umma::smem_u32(smem) + 32768, 128, (uint32_t)b_sbo_bytes);
        const uint64_t bstep = b_sbo_bytes > 256 ? 16 : 512;
        const long long t0 = clock64();
        if (b_sbo_bytes == 18432) {
            // the exact operand walk of kpn.cu: 4 halo parts x 9 taps x 2 K steps = 72 distinct (A, B) pairs per item
            const uint64_t da_k = umma::smem_desc(umma::smem_u32(smem), 2880, 160);
            const uint64_t db_k = umma::smem_desc(umma::smem_u32(smem) + 46080, 128, 18432);
            for (int i = 0; i < iters; i += 72) {
#pragma unroll
                for (int part = 0; part < 4; ++part)
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            if (leader)
                                umma::mma_f16(tmem, da_k + (uint64_t)(part * 720 + (tap / 3) * 10 + tap % 3 + j * 360),
                                              db_k + (uint64_t)(((part * 9 + tap) * 2 + j) * 16), idesc, true);
                            // commit_every > 0: a tcgen05.commit (to a barrier nobody waits on) after every n-th MMA
                            if (COMMIT_EVERY > 0 && ((part * 9 + tap) * 2 + j + 1) % COMMIT_EVERY == 0 && leader) umma::commit(&bar2);
                        }
            }
        } else
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (leader) umma::mma_f16(tmem, da + (uint64_t)(u & 3), db + bstep * (uint64_t)(u & 7), idesc, true);
        }
        if (leader) umma::commit(&bar);
        umma::mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (leader) cycles_per_mma[blockIdx.x] = (float)(t1 - t0) / (float)iters;
        __syncwarp();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<256>(tmem);
}
}  // namespace

extern "C" int ebfi_selftest_mma_rate(void *stream, float *cycles_per_mma, int n_ctas, int N, int iters, int a_sbo_bytes,
                                      int b_sbo_bytes, int commit_every)
{
    EBFI_REQUIRE(b_sbo_bytes >= 256 && b_sbo_bytes % 16 == 0 &&
                     (b_sbo_bytes == 18432 ? (N <= 80 && iters % 72 == 0) : 32768 + (N / 8) * (long)b_sbo_bytes <= 200 * 1024),
                 "mma_rate: B operand does not fit the probe's shared memory");
    EBFI_REQUIRE(cycles_per_mma != nullptr && n_ctas > 0, "mma_rate: bad arguments");
    EBFI_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && iters >= 8 && iters % 8 == 0, "mma_rate: N multiple of 16 in [16, 256], iters multiple of 8");
    const int smem = 225 * 1024;      // one CTA per SM
#define EBFI_RATE(CE)                                                                                          \
    do {                                                                                                       \
        EBFI_CUDA_OK(cudaFuncSetAttribute(mma_rate_kernel<CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        mma_rate_kernel<CE><<<n_ctas, 128, smem, ebfi::as_stream(stream)>>>(cycles_per_mma, N, iters, a_sbo_bytes, b_sbo_bytes); \
    } while (0)
    switch (commit_every) {
    case 0: EBFI_RATE(0); break;
    case 1: EBFI_RATE(1); break;
    case 9: EBFI_RATE(9); break;
    case 18: EBFI_RATE(18); break;
    case 72: EBFI_RATE(72); break;
    default: return ebfi::fail(EBFI_ERR_INVALID, "mma_rate: commit_every must be 0, 1, 9, 18 or 72");
    }
#undef EBFI_RATE
    EBFI_LAUNCH_OK("mma_rate_kernel");
    return EBFI_OK;
}
