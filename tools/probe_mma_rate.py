"""tcgen05 bf16 MMA issue interval vs N (M = 128, K = 16), one CTA per SM: is a small-N GEMM math- or operand-bound?"""
import sys, torch
sys.path.insert(0, "/root/repo")
from ebfi_be_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda:0")
out = torch.zeros(148, device=dev)
for sbo, bsbo in ((128, 256), (160, 256), (160, 18432), (160, 2048), (160, 2304)):
    for N in (32, 64, 80, 96, 128, 160, 256):
        if 32768 + (N // 8) * bsbo > 200 * 1024:
            continue
        for n_ctas in (1, 148):
            L.check(lib.ebfi_selftest_mma_rate(L.stream_ptr(dev), L.ptr(out), n_ctas, N, 4096, sbo, bsbo), "mma_rate")
            torch.cuda.synchronize()
            c = out[:n_ctas].mean().item()
            flops = 2 * 128 * N * 16
            print(f"A SBO {sbo:3d}  B SBO {bsbo:5d}  N={N:3d}  CTAs={n_ctas:3d}: {c:6.1f} clk/MMA  -> {flops / c:7.0f} flop/clk/SM ({flops / c / 8192 * 100:5.1f} % of 8192)")
