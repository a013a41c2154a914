import sys, torch, os
sys.path.insert(0, "/root/repo")
from ebfi_be_b200.shims import _ext
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(1, 64, 256, 256, device=dev); w = torch.randn(64, 64, 3, 3, device=dev) / 24; b = torch.randn(64, device=dev)
off = 2 * torch.randn(1, 144, 256, 256, device=dev); msk = torch.sigmoid(torch.randn(1, 72, 256, 256, device=dev))
geom = (3, 3, 1, 1, 1, 1, 1, 1, 8)
for _ in range(3):
    _ext.dcn_v2_forward(x, w, b, off, msk, *geom)
torch.cuda.synchronize()
os.environ["EBFI_DCN_TIMELINE"] = "1"
_ext.dcn_v2_forward(x, w, b, off, msk, *geom)
torch.cuda.synchronize()
