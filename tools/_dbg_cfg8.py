import sys, torch, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from ebfi_be_b200.shims import _ext
dev = torch.device("cuda:0")
def run(B, C, Co, H, W, kh, kw, sh, sw, ph, pw, dh, dw, dg, osc):
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    x = torch.randn(B, C, H, W, device=dev); w = torch.randn(Co, C, kh, kw, device=dev) / 10; b = torch.randn(Co, device=dev)
    off = osc * torch.randn(B, 2 * dg * kh * kw, Ho, Wo, device=dev); msk = torch.rand(B, dg * kh * kw, Ho, Wo, device=dev)
    o = _ext.dcn_v2_forward(x, w, b, off, msk, kh, kw, sh, sw, ph, pw, dh, dw, dg)
    torch.cuda.synchronize(); print(" ok", float(o.abs().sum()), flush=True)
run(1, 16, 64, 24, 24, 2, 3, 1, 1, 0, 1, 1, 1, 2, 0.5)     # TPR 2
