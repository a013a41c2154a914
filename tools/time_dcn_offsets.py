"""DCNv2 forward/backward time vs offset spread (dev tool): the op is bound by how many cache lines a warp's
gathers touch, which depends on the offset field, not by HBM."""
import sys, torch
sys.path.insert(0, "/root/repo")
from ebfi_be_b200.shims import _ext
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=10):
    ts = []
    for _ in range(n):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
geom = (3, 3, 1, 1, 1, 1, 1, 1, 8)
for B in (1, 8):
    x = torch.randn(B, 64, 256, 256, device=dev); w = torch.randn(64, 64, 3, 3, device=dev) / 24; bias = torch.randn(64, device=dev)
    msk = torch.sigmoid(torch.randn(B, 72, 256, 256, device=dev)); go = torch.randn(B, 64, 256, 256, device=dev)
    for name, off in [("zero", torch.zeros(B, 144, 256, 256, device=dev)),
                      ("0.5*randn", 0.5 * torch.randn(B, 144, 256, 256, device=dev)),
                      ("smooth 2px", 2 * torch.nn.functional.interpolate(torch.randn(B, 144, 16, 16, device=dev), size=256, mode="bilinear")),
                      ("2*randn (bench)", 2 * torch.randn(B, 144, 256, 256, device=dev)),
                      ("10*randn", 10 * torch.randn(B, 144, 256, 256, device=dev))]:
        off = off.contiguous()
        tf = timed(lambda: _ext.dcn_v2_forward(x, w, bias, off, msk, *geom))
        tb = timed(lambda: _ext.dcn_v2_backward(x, w, bias, off, msk, go, *geom))
        print(f"B={B} offsets {name:16s} fwd {tf:.4f} ms  bwd {tb:.4f} ms  -> {B*65536/1e6/((tf+tb)*1e-3):7.1f} Mpix/s")
