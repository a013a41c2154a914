"""KernelConv producer -> FAC: fused kernel vs the reference's op sequence (cuDNN conv + LeakyReLU + FAC) at
BASELINE cfg2 (B=4, C=64, K=5, 256x256) and at the inference shape of cfg4 (B=1, 360x640). Prints JSON lines."""
import json, sys, torch
sys.path.insert(0, "/root/repo")
from ebfi_be_b200 import modification
from ebfi_be_b200.kernelconv2d import KernelConv2D
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def timed(fn, n=10):
    ts = []
    for _ in range(n + 2):
        flush.zero_(); a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]

torch.manual_seed(0)
for (B, H, W) in [(4, 256, 256), (1, 360, 640)]:
    C, K = 64, 5
    ev, fr = torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)
    conv = torch.nn.Conv2d(2 * C, C * K * K, 3, 1, 1).to(dev)
    act, kpn = torch.nn.LeakyReLU(), KernelConv2D(kernel_size=K)
    with torch.no_grad():
        def unfused():
            return kpn(ev, act(conv(torch.cat([ev, fr], 1))))
        def fused():
            return modification.kernelconv_fac_fused(ev, fr, conv.weight, conv.bias, K, 0.01)
        res = {}
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            res["unfused_ms_cudnn_" + ("tf32" if tf32 else "fp32")] = round(timed(unfused), 4)
        torch.backends.cudnn.allow_tf32 = True
        evh, frh, convh = ev.bfloat16(), fr.bfloat16(), conv.bfloat16()
        def unfused_bf16():
            return kpn(evh, act(convh(torch.cat([evh, frh], 1))))
        res["unfused_ms_bf16_tensors"] = round(timed(unfused_bf16), 4)
        conv = conv.float()
        t = timed(fused)
        err = float((fused() - unfused()).abs().max() / unfused().abs().max())
    flops = 2.0 * B * H * W * (C * K * K) * (2 * C * 9)
    res.update({"workload": f"B={B} C={C} K={K} {H}x{W}", "fused_ms": round(t, 4), "fused_TFLOPs": round(flops / (t * 1e-3) / 1e12, 1),
                "Mpix_s": round(B * H * W / 1e6 / (t * 1e-3), 1), "max_rel_diff_vs_unfused_tf32": err})
    print(json.dumps(res))
