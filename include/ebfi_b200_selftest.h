/* ebfi_b200_selftest.h — hardware probes of the tcgen05 / TMEM plumbing the DCN and KernelConv kernels are built on.
 * TEST-ONLY: these entry points live in libebfi_b200_selftest.so (built next to libebfi_b200.so, linked against it),
 * not in the product library, and are bound by tests/test_tcgen05_gpu.py and tools/probe_*.py only. */
#ifndef EBFI_B200_SELFTEST_H_
#define EBFI_B200_SELFTEST_H_
#include "ebfi_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* One-CTA GEMM on the tcgen05 tensor-core path with the 3xTF32 split the DCN kernels use:
 * C[M x N] = A[M x K] * B[N x K]^T, fp32 in / fp32 out, ~fp32 accuracy.
 * M in {64, 128}; N <= 128, multiple of 16 (M=128) or 8 (M=64); K multiple of 8.
 * A is row-major [M][K], or [K][M] when a_mn_major != 0 (M = 128 only); B is [N][K].
 * Exists to validate descriptors / TMEM plumbing on real hardware; not a product entry. */
int ebfi_selftest_gemm_tf32x3(void *stream, const float *A, const float *B, float *C,
                              int M, int N, int K, int a_mn_major);

/* Same GEMM with bf16 hi/lo pairs on kind::f16 (the DCN backward's operand format, ~2^-16
 * relative accuracy). K multiple of 16, <= 128. b_lbo_bytes: byte distance between consecutive
 * 16-byte K chunks of B (128 = dense, 144 = the padded layout the backward kernel uses). */
int ebfi_selftest_gemm_bf16x3(void *stream, const float *A, const float *B, float *C,
                              int M, int N, int K, int b_lbo_bytes);

/* Layout probe: C[128][8] receives, for every element A(m, k) of a 128 x 8 TF32 operand described
 * by (lbo, sbo, major-ness), the float index inside shared memory that the tensor core fetched.
 * Documents how the hardware interprets the descriptor fields. */
int ebfi_selftest_umma_probe(void *stream, float *C, int lbo_bytes, int sbo_bytes, int a_mn_major);

/* Tensor-pipe rate probe: n_ctas CTAs (one per SM) each issue `iters` back-to-back bf16 MMAs of shape 128 x N x 16
 * on shared-memory operands and write the measured clock cycles per MMA (n_ctas floats). a_sbo_bytes: stride between
 * the 8-row groups of the A operand (128 = dense; 160 = the halo view of the fused KernelConv kernel); b_sbo_bytes:
 * the same for B (256 = dense N x 16; 18432 = the weight image and operand walk of the fused KernelConv kernel, where
 * commit_every > 0 additionally issues a tcgen05.commit after every commit_every-th MMA of the 72-MMA item). */
int ebfi_selftest_mma_rate(void *stream, float *cycles_per_mma, int n_ctas, int N, int iters, int a_sbo_bytes,
                           int b_sbo_bytes, int commit_every);

#ifdef __cplusplus
}
#endif
#endif /* EBFI_B200_SELFTEST_H_ */
