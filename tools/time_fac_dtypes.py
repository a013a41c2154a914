import sys, torch
sys.path.insert(0, "/root/repo")
from ebfi_be_b200.shims import kernelconv2d_cuda as kc
dev = torch.device("cuda:0")
B, C, K, H, W = 4, 64, 5, 256, 256
for dt in (torch.float32, torch.bfloat16):
    x = torch.randn(B, C, H + 4, W + 4, device=dev).to(dt); ker = torch.randn(B, C * 25, H, W, device=dev).to(dt)
    go = torch.randn(B, C, H, W, device=dev).to(dt)
    out = torch.empty(B, C, H, W, device=dev, dtype=dt); gi = torch.empty_like(x); gk = torch.empty_like(ker)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    def timed(fn, n=10):
        ts = []
        for _ in range(n):
            flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]
    es = 4 if dt == torch.float32 else 2
    tf = timed(lambda: kc.forward(x, ker, K, out)); tb = timed(lambda: kc.backward(x, ker, K, go, gi, gk))
    fb = es * (x.numel() + ker.numel() + out.numel()); bb = es * (2 * ker.numel() + go.numel() + 2 * x.numel())
    print(dt, f"fwd {tf:.4f} ms {fb/tf/1e6:.0f} GB/s   bwd {tb:.4f} ms {bb/tb/1e6:.0f} GB/s")
