"""FAC forward/backward timings at the shapes the model really runs (dev tool, L2 flushed between calls)."""
import sys, torch
sys.path.insert(0, "/root/repo")
from ebfi_be_b200.shims import kernelconv2d_cuda as kc
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=10):
    ts = []
    for _ in range(n):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for (B, C, K, H, W) in [(4, 64, 5, 256, 256), (8, 64, 5, 128, 128), (1, 64, 5, 360, 640), (2, 64, 5, 720, 1280), (4, 64, 3, 256, 256)]:
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(B, C, H + K - 1, W + K - 1, device=dev).to(dt); ker = torch.randn(B, C * K * K, H, W, device=dev).to(dt)
        go = torch.randn(B, C, H, W, device=dev).to(dt)
        out = torch.empty(B, C, H, W, device=dev, dtype=dt); gi = torch.empty_like(x); gk = torch.empty_like(ker)
        es = x.element_size()
        tf = timed(lambda: kc.forward(x, ker, K, out)); tb = timed(lambda: kc.backward(x, ker, K, go, gi, gk))
        fb = es * (x.numel() + ker.numel() + out.numel()); bb = es * (2 * ker.numel() + go.numel() + 2 * x.numel())
        print(f"B{B} C{C} K{K} {H}x{W} {str(dt)[6:]:9s} fwd {tf:.4f} ms {fb/tf/1e6:6.0f} GB/s   bwd {tb:.4f} ms {bb/tb/1e6:6.0f} GB/s")
        del x, ker, go, out, gi, gk
