"""tcgen05 bf16 MMA issue interval vs N (M = 128, K = 16), one CTA per SM: is a small-N GEMM math- or operand-bound?"""
import sys, torch
sys.path.insert(0, "/root/repo")
from ebfi_be_b200 import _lib as L
lib = L.load_selftest()
dev = torch.device("cuda:0")
out = torch.zeros(148, device=dev)
for sbo, bsbo in ((128, 256), (160, 256), (160, 18432)):
    for N in (32, 64, 80, 96, 128, 160, 256):
        if bsbo == 18432 and N > 80:
            continue
        for n_ctas in (1, 148):
            L.check(lib.ebfi_selftest_mma_rate(L.stream_ptr(dev), L.ptr(out), n_ctas, N, 4608 if bsbo == 18432 else 4096, sbo, bsbo, 0), "mma_rate")
            torch.cuda.synchronize()
            c = out[:n_ctas].mean().item()
            flops = 2 * 128 * N * 16
            print(f"A SBO {sbo:3d}  B SBO {bsbo:5d}  N={N:3d}  CTAs={n_ctas:3d}: {c:6.1f} clk/MMA  -> {flops / c:7.0f} flop/clk/SM ({flops / c / 8192 * 100:5.1f} % of 8192)")

print("commit frequency, kpn.cu operand walk, N = 80:")
for every in (0, 72, 18, 9, 1):
    L.check(lib.ebfi_selftest_mma_rate(L.stream_ptr(dev), L.ptr(out), 148, 80, 4608, 160, 18432, every), "mma_rate")
    torch.cuda.synchronize()
    print(f"  tcgen05.commit after every {every:2d} MMAs: {out.mean().item():6.1f} clk/MMA")
