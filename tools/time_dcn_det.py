"""DCNv2 backward at BASELINE cfg1: default (vector float reductions) vs EBFI_DCN_DETERMINISTIC (int64 fixed point)."""
import os, sys, torch
sys.path.insert(0, "/root/repo")
from ebfi_be_b200.shims import _ext
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
torch.manual_seed(0)
x = torch.randn(1, 64, 256, 256, device=dev); w = torch.randn(64, 64, 3, 3, device=dev) / 24; b = torch.randn(64, device=dev)
off = 2 * torch.randn(1, 144, 256, 256, device=dev); msk = torch.sigmoid(torch.randn(1, 72, 256, 256, device=dev))
go = torch.randn(1, 64, 256, 256, device=dev)
geom = (3, 3, 1, 1, 1, 1, 1, 1, 8)
def timed(fn, n=20):
    ts = []
    for _ in range(n):
        flush.zero_(); a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]
for mode in ("0", "1"):
    os.environ["EBFI_DCN_DETERMINISTIC"] = mode
    print("deterministic =", mode, "bwd ms", round(timed(lambda: _ext.dcn_v2_backward(x, w, b, off, msk, go, *geom)), 4))
