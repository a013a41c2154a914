import sys, torch, statistics
sys.path.insert(0, "/root/repo")
import ebfi_be_b200
from ebfi_be_b200.shims import _ext
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1234)
r = lambda *s: torch.randn(*s, generator=g)
x, off, msk = r(1, 64, 256, 256), 2 * r(1, 144, 256, 256), torch.sigmoid(r(1, 72, 256, 256))
w, b, go = (torch.rand(64, 64, 3, 3, generator=g) * 2 - 1) / 24, r(64), r(1, 64, 256, 256)
x, off, msk, w, b, go = (v.to(dev) for v in (x, off, msk, w, b, go))
geom = (3, 3, 1, 1, 1, 1, 1, 1, 8)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=20, cold=True):
    ts = []
    for _ in range(n):
        if cold: flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c))
    return statistics.median(ts)
for scale in (2.0, 0.0, 0.5, 5.0):
    o2 = (off * (scale / 2.0)).contiguous()
    f = lambda: _ext.dcn_v2_forward(x, w, b, o2, msk, *geom)
    bw = lambda: _ext.dcn_v2_backward(x, w, b, o2, msk, go, *geom)
    f(); bw(); torch.cuda.synchronize()
    print(f"offset sigma {scale}: fwd cold {timed(f):.4f} warm {timed(f, cold=False):.4f} ms | bwd cold {timed(bw):.4f} warm {timed(bw, cold=False):.4f} ms", flush=True)
