"""DCN_sep module (models/DCNv2/dcn_v2.py:197-227) after the offset conv, forward + backward at BASELINE cfg1:
the reference's op sequence (chunk, cat, |.|.mean() + host sync, sigmoid, dcn_v2_conv and their autograd)
vs the packed kernels (raw conv output in, one gradient tensor out, deferred warning). Prints one JSON line."""
import json, sys, torch
sys.path.insert(0, "/root/repo")
from ebfi_be_b200 import dcn_v2
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
torch.manual_seed(0)
B, C, H, W, dg = 1, 64, 256, 256, 8
x = torch.randn(B, C, H, W, device=dev, requires_grad=True)
om = torch.randn(B, 3 * dg * 9, H, W, device=dev)
om[:, :144] *= 2
om.requires_grad_()
w = (torch.randn(C, C, 3, 3, device=dev) / 24).requires_grad_()
b = torch.randn(C, device=dev, requires_grad=True)
go = torch.randn(B, C, H, W, device=dev)
watch = dcn_v2._OffsetWatch()

def unfused():
    o1, o2, mask = torch.chunk(om, 3, dim=1)
    offset = torch.cat((o1, o2), dim=1)
    if torch.mean(torch.abs(offset)) > 100:
        print("warn")
    out = dcn_v2.dcn_v2_conv(x, offset, torch.sigmoid(mask), w, b, 1, 1, 1, dg)
    out.backward(go)

def fused():
    watch.poll()
    stat = torch.empty(1, device=dev)
    out = dcn_v2.dcn_v2_conv_packed(x, om, w, b, 1, 1, 1, dg, stat)
    watch.submit(stat, om.numel() // 3 * 2)
    out.backward(go)

def timed(fn, n=30):
    ts = []
    for _ in range(n + 3):
        for v in (x, om, w, b):
            v.grad = None
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    ts = sorted(ts[3:])
    return ts[len(ts) // 2]

tu, tf = timed(unfused), timed(fused)
print(json.dumps({"workload": "DCN_sep tail fwd+bwd, cfg1 (B=1, 64->64, 3x3, dg=8, 256x256, fp32), L2 flushed",
                  "reference_op_sequence_ms": round(tu, 4), "packed_ms": round(tf, 4),
                  "speedup": round(tu / tf, 3), "Mpix_s_packed": round(B * H * W / 1e6 / (tf * 1e-3), 1)}))
