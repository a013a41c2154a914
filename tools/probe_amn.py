"""Probe: kind::f16 (bf16) MMA with an MN-major A operand in the no-swizzle core-matrix layout (bit 16 of the LBO argument)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ebfi_be_b200 import _lib as L
dev = torch.device("cuda:0")
for (M, N, K) in [(128, 80, 64), (128, 16, 16), (128, 64, 128)]:
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    want = A.double() @ B.double().t()
    Ad, Bd = A.to(dev), B.to(dev)
    for amn in (0, 1):
        C = torch.full((M, N), float("nan"), device=dev)
        L.check(L.load_selftest().ebfi_selftest_gemm_bf16x3(L.stream_ptr(dev), L.ptr(Ad), L.ptr(Bd), L.ptr(C), M, N, K, 128 | (amn << 16)), "selftest")
        torch.cuda.synchronize()
        err = (C.double().cpu() - want).abs().max() / want.abs().max()
        print(f"M={M} N={N} K={K} a_mn={amn}: rel err {float(err):.3e}", flush=True)
