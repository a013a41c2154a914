"""Two backward runs of BASELINE configs[0] must agree bit for bit in all five gradients (default mode)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ebfi_be_b200.shims import _ext
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1234)
r = lambda *s: torch.randn(*s, generator=g)
for sigma in (2.0, 6.0):
    x, off, msk = r(1, 64, 256, 256), sigma * r(1, 144, 256, 256), torch.sigmoid(r(1, 72, 256, 256))
    w, b, go = (torch.rand(64, 64, 3, 3, generator=g) * 2 - 1) / 24, r(64), r(1, 64, 256, 256)
    x, off, msk, w, b, go = (v.to(dev) for v in (x, off, msk, w, b, go))
    geom = (3, 3, 1, 1, 1, 1, 1, 1, 8)
    a = _ext.dcn_v2_backward(x, w, b, off, msk, go, *geom)
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    c = _ext.dcn_v2_backward(x, w, b, off, msk, go, *geom)
    torch.cuda.synchronize()
    print(f"sigma {sigma}:", [bool(torch.equal(u, v)) for u, v in zip(a, c)], flush=True)
