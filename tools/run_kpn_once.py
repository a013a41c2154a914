"""One fused KernelConv -> FAC forward at BASELINE cfg2 (for ncu captures)."""
import sys, torch
sys.path.insert(0, "/root/repo")
from ebfi_be_b200 import modification
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, C, K, H, W = 4, 64, 5, 256, 256
ev, fr = torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)
w = torch.randn(C * K * K, 2 * C, 3, 3, device=dev) * 0.03
b = torch.randn(C * K * K, device=dev) * 0.1
with torch.no_grad():
    for _ in range(3):
        modification.kernelconv_fac_fused(ev, fr, w, b, K, 0.01)
torch.cuda.synchronize()
