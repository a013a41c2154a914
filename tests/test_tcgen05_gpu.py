"""GPU: the tcgen05 / TMEM plumbing (ebfi-be_b200/csrc/umma.cuh) on its own — a one-CTA 3xTF32
GEMM against an fp64 reference. The DCN GEMMs are built from exactly these pieces."""
import pytest
import torch

pytestmark = pytest.mark.gpu


# K-major operands only: with kind::tf32, an MN-major A in the no-swizzle layout reads back as zeros
# on B200 (probed with ebfi_selftest_umma_probe), so the DCN kernels keep every operand K-major.
@pytest.mark.parametrize("M,N,K,a_mn", [(128, 64, 72, 0), (128, 64, 576, 0), (128, 16, 8, 0), (64, 72, 128, 0),
                                        (64, 8, 512, 0), (128, 80, 64, 0), (128, 128, 200, 0)])
def test_gemm_tf32x3_matches_fp64(M, N, K, a_mn):
    from ebfi_be_b200 import _lib as L
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(M * 1000 + N * 10 + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    want = A.double() @ B.double().t()
    Ad = (A.t().contiguous() if a_mn else A).to(dev)
    Bd = B.to(dev)
    C = torch.full((M, N), float("nan"), device=dev)
    L.check(L.load_selftest().ebfi_selftest_gemm_tf32x3(L.stream_ptr(dev), L.ptr(Ad), L.ptr(Bd), L.ptr(C), M, N, K, a_mn),
            "selftest_gemm")
    torch.cuda.synchronize()
    err = (C.double().cpu() - want).abs().max() / want.abs().max()
    print(f"M={M} N={N} K={K} a_mn={a_mn}: rel err {float(err):.3e}")
    assert err < 2e-6, float(err)          # plain TF32 would sit near 1e-3


@pytest.mark.parametrize("M,N,K,lbo", [(128, 80, 64, 128), (64, 80, 128, 144), (128, 16, 16, 128), (64, 72, 128, 144)])
def test_gemm_bf16x3_matches_fp64(M, N, K, lbo):
    """bf16 hi/lo pairs on kind::f16, dense and padded (LBO = 144 B) B operand."""
    from ebfi_be_b200 import _lib as L
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    want = A.double() @ B.double().t()
    Ad, Bd = A.to(dev), B.to(dev)            # keep the device copies alive across the call
    C = torch.full((M, N), float("nan"), device=dev)
    L.check(L.load_selftest().ebfi_selftest_gemm_bf16x3(L.stream_ptr(dev), L.ptr(Ad), L.ptr(Bd), L.ptr(C), M, N, K, lbo),
            "selftest_gemm_bf16")
    torch.cuda.synchronize()
    err = (C.double().cpu() - want).abs().max() / want.abs().max()
    print(f"bf16x3 M={M} N={N} K={K} lbo={lbo}: rel err {float(err):.3e}")
    assert err < 3e-5, float(err)


@pytest.mark.parametrize("M,N,K", [(128, 80, 64), (128, 16, 16), (128, 64, 128)])
def test_gemm_bf16x3_mn_major_a(M, N, K):
    """kind::f16 accepts an MN-major A operand in the no-swizzle core-matrix layout (kind::tf32 does not): the same
    128-byte core matrices serve a K-major read over one index and an MN-major read over the other, which is how the
    DCN backward uses ONE grad_output copy for both of its contractions (dcn_bwd_box.cu). Bit 16 of the LBO argument
    selects the MN-major store + descriptor in the self test."""
    from ebfi_be_b200 import _lib as L
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    want = A.double() @ B.double().t()
    Ad, Bd = A.to(dev), B.to(dev)
    outs = []
    for amn in (0, 1):
        C = torch.full((M, N), float("nan"), device=dev)
        L.check(L.load_selftest().ebfi_selftest_gemm_bf16x3(L.stream_ptr(dev), L.ptr(Ad), L.ptr(Bd), L.ptr(C), M, N, K, 128 | (amn << 16)),
                "selftest_gemm_bf16")
        torch.cuda.synchronize()
        outs.append(C.cpu())
        assert float((C.double().cpu() - want).abs().max() / want.abs().max()) < 3e-5
    assert torch.equal(outs[0], outs[1])         # same products, same accumulation order


def test_probe_documents_the_k_major_core_matrix_layout():
    """addr(row, k) = (row/8)*SBO + (k/4)*LBO + (row%8)*16 + (k%4)*4 bytes — read back from the hardware."""
    from ebfi_be_b200 import _lib as L
    dev = torch.device("cuda:0")
    C = torch.zeros(128, 8, device=dev)
    L.check(L.load_selftest().ebfi_selftest_umma_probe(L.stream_ptr(dev), L.ptr(C), 128, 256, 0), "probe")
    torch.cuda.synchronize()
    got = C.cpu().long()
    for m in (0, 1, 7, 8, 9, 64, 127):
        for k in range(8):
            assert int(got[m, k]) * 4 == (m // 8) * 256 + (k // 4) * 128 + (m % 8) * 16 + (k % 4) * 4
