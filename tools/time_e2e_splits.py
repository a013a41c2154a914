"""e2e leg of bench.py (pinned host buffers in and out, FAC fwd+bwd at cfg2) vs the pipeline's channel split."""
import sys, time, torch
sys.path.insert(0, "/root/repo")
from ebfi_be_b200.host_pipeline import HostPipeline
dev = torch.device("cuda:0")
B, C, K, H, W = 4, 64, 5, 256, 256
g = torch.Generator().manual_seed(0)
xi = torch.randn(B, C, H + 4, W + 4, generator=g).pin_memory()
ker = (0.1 * torch.randn(B, C * 25, H, W, generator=g)).pin_memory()
go = torch.randn(B, C, H, W, generator=g).pin_memory()
out, gi, gk = torch.empty_like(go).pin_memory(), torch.empty_like(xi).pin_memory(), torch.empty_like(ker).pin_memory()
pipe = HostPipeline(dev)
nbytes = sum(t.numel() * 4 for t in (xi, ker, go))
for ns in (1, 2, 4, 8, 16, 32):
    for _ in range(2):
        pipe.fac_forward_backward(xi, ker, go, K, out, gi, gk, channel_splits=ns)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        pipe.fac_forward_backward(xi, ker, go, K, out, gi, gk, channel_splits=ns)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"channel_splits {ns:2d}: {dt * 1e3:7.2f} ms  ({nbytes / dt / 1e9:5.1f} GB/s each way)")
