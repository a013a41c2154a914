"""Fused KernelConv -> FAC forward only, cfg2 shape: median ms over 20 L2-flushed calls (dev tool)."""
import sys, torch
sys.path.insert(0, "/root/repo")
from ebfi_be_b200 import modification
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
torch.manual_seed(0)
B, C, K, H, W = 4, 64, 5, 256, 256
ev, fr = torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)
w = torch.randn(C * K * K, 2 * C, 3, 3, device=dev) * 0.03
b = torch.randn(C * K * K, device=dev) * 0.1
ts = []
with torch.no_grad():
    for i in range(23):
        flush.zero_(); a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); modification.kernelconv_fac_fused(ev, fr, w, b, K, 0.01); e.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(e))
ts.sort(); print("kpn fused ms: median %.4f  min %.4f" % (ts[len(ts) // 2], ts[0]))
