"""bench.py legs for BASELINE configs[3] (full `models/Ours` 720p inference, bf16) and configs[4] (training step,
256x256 crops, B=8 per GPU, data-parallel with the weight-gradient all-reduce enabled).

The model is the reference's own, unmodified `EVFIAutoEx` (staged by tools/refmodel.py into the git-ignored
baseline/_ref/ebfi_be) running on this repo's `_ext` / `kernelconv2d_cuda` shims; everything that is not the FAC
operator is stock torch / cuDNN (out of scope). A/B arm: the same model with the reference's own FAC
CUDA kernels compiled for sm_100a (oracle/_ref/fac_cuda, fp32 only — they have no bf16 path).

Reference: models/Ours/model_singleframe.py:226-348 (model), config/train_ours.yml:26-65 (hyper-parameters),
train_ours.py:221-277,754-765 (step, losses, optimizer).
"""
import contextlib
import os
import statistics
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import refmodel  # noqa: E402


def _fac_timer(torch, model):
    """CUDA-event pairs around every call of the model's KPN (= KernelConv2D) module."""
    pairs = []

    def pre(m, inp):
        e = torch.cuda.Event(enable_timing=True); e.record(); pairs.append([e, None])

    def post(m, inp, out):
        e = torch.cuda.Event(enable_timing=True); e.record(); pairs[-1][1] = e

    kpn = model.Modification.KPN
    return pairs, [kpn.register_forward_pre_hook(pre), kpn.register_forward_hook(post)]


def _synthetic_720p(torch, dev, B=1, H=720, W=1280, TB=16, seed=4):
    g = torch.Generator(device="cpu").manual_seed(seed)
    frame = torch.rand(B, 3, H, W, generator=g).to(dev)
    event = torch.round(2 * torch.rand(B, TB, 2, H, W, generator=g)).to(dev)    # non-negative counts, like events_to_stack
    t = torch.rand(B, 1, generator=g).to(dev)
    return frame, event, t


def cfg4_inference(torch, dev, reps=5, warm=2):
    """BASELINE configs[3]: EVFIAutoEx(**train_ours.yml) .eval(), random init (seed 0), one 1280x720 frame."""
    if not refmodel.available():
        return {"unavailable": "baseline/_ref/ebfi_be not staged (run __graft_entry__.build() where /root/reference exists)"}
    mod = refmodel.load()
    import ebfi_be_b200
    from ebfi_be_b200 import frame_ops, modification
    model = refmodel.build_model(seed=0).to(dev).eval()
    frame, event, t = _synthetic_720p(torch, dev)
    orig_lap, orig_mod_fwd = mod.Frame2Lap, mod.Modification.forward

    def fused_modification_forward(self, FrameTensor, EventTensor):
        # model_singleframe.py:151-165 with KernelConv -> KPN replaced by the fused kernel
        EventTensor = self.Conv1(EventTensor)
        conv = self.KernelConv.conv2d
        if EventTensor.dtype != FrameTensor.dtype:
            FrameTensor = FrameTensor.to(EventTensor.dtype)
        ev1 = modification.kernelconv_fac_fused(EventTensor, FrameTensor, conv.weight.to(EventTensor.dtype),
                                                conv.bias.to(EventTensor.dtype), 5, self.KernelConv.activation.negative_slope)
        EventTensor1 = self.Conv3(ev1)
        return FrameTensor * EventTensor1 + self.Conv2(EventTensor1)

    def run(bf16):
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if bf16 else contextlib.nullcontext()
        with torch.no_grad(), ctx:
            return model(frame, event, t)

    def measure(bf16):
        pairs, hooks = _fac_timer(torch, model)
        for _ in range(warm):
            run(bf16)
        torch.cuda.synchronize()
        pairs.clear()
        ts = []
        out = None
        for _ in range(reps):
            t0 = time.perf_counter()
            out = run(bf16)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        for h in hooks:
            h.remove()
        fac = statistics.median([a.elapsed_time(b) for a, b in pairs]) if pairs else None
        ms = statistics.median(ts)
        return {"ms_per_frame": round(ms, 3), "Mpix_s": round(0.9216 / (ms * 1e-3), 2),
                "fac_ms": None if fac is None else round(fac, 4),
                "fac_share": None if fac is None else round(fac / ms, 4)}, out

    res = {"workload": "EVFIAutoEx(train_ours.yml) eval, random init, Frame 1x3x720x1280 + Event 1x16x2x720x1280 (synthetic), "
                       "FAC at 360x640 with a (1,1600,360,640) kernel tensor; wall clock per forward incl. the model's own "
                       "host round trip (Frame2Lap via OpenCV) unless stated",
           "params": sum(p.numel() for p in model.parameters())}
    res["fp32_ours"], out_fp32 = measure(False)
    if refmodel.use_reference_cuda_fac(True):
        res["fp32_reference_fac_cuda_kernels_sm100a"], out_ref = measure(False)
        res["fp32_final_max_abs_diff_ours_vs_reference_kernels"] = float((out_fp32[1] - out_ref[1]).abs().max())
    refmodel.use_reference_cuda_fac(False)
    res["bf16_ours"], out_bf16 = measure(True)
    num = (out_bf16[1].float() - out_fp32[1]).norm()
    res["bf16_vs_fp32_final_rel_l2"] = float(num / out_fp32[1].norm())
    # widenings: frame map on the GPU, KernelConv -> FAC fused (INTEGRATION.md shows both edits)
    mod.Frame2Lap = frame_ops.Frame2Lap
    res["bf16_ours_gpu_frame_map"], _ = measure(True)
    mod.Modification.forward = fused_modification_forward
    try:
        r, out_fused = measure(True)
        r["fac_ms"] = r["fac_share"] = None       # the KPN module is bypassed
        res["bf16_ours_gpu_frame_map_fused_kernelconv_fac"] = r
        res["bf16_fused_vs_fp32_final_rel_l2"] = float((out_fused[1].float() - out_fp32[1]).norm() / out_fp32[1].norm())
    finally:
        mod.Frame2Lap, mod.Modification.forward = orig_lap, orig_mod_fwd
    res["bf16_tolerance"] = "stated: rel-L2 of the final frame vs the fp32 run <= 2e-2 (autocast bf16 convolutions dominate)"
    del model
    torch.cuda.empty_cache()
    return res


def cfg5_train_step(torch, dev, world, steps=8, warm=3, batch=8, size=256, ab=True):
    """BASELINE configs[4]: one optimizer step of the reference model (train_ours.py:250-277) on B=8 256x256 crops per
    GPU, Adam lr 1e-4 (train_ours.yml:59-65), loss = Lap + census on both outputs (train_ours.py:261-262,762-764),
    DistributedDataParallel with the gradient all-reduce ENABLED (the reference wraps its step in no_sync(), :250,
    and therefore never reduces). Timed with CUDA events, max over ranks."""
    if not refmodel.available():
        return {"unavailable": "baseline/_ref/ebfi_be not staged"}
    import torch.distributed as dist
    mod = refmodel.load()
    from ebfi_be_b200 import frame_ops
    import loss as ref_loss
    rank = dist.get_rank() if world > 1 else 0
    model = refmodel.build_model(seed=0).to(dev).train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index]) if world > 1 else model
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.999), amsgrad=False)
    lap, census = ref_loss.LaplacianLoss().to(dev), ref_loss.Ternary(7)
    if not torch.is_tensor(census.w) or not census.w.is_cuda:
        census.w = torch.as_tensor(census.w).float().to(dev)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    frame = torch.rand(batch, 3, size, size, generator=g).to(dev)
    event = torch.round(2 * torch.rand(batch, 16, 2, size, size, generator=g)).to(dev)
    t = torch.rand(batch, 1, generator=g).to(dev)
    gt = torch.rand(batch, 3, size, size, generator=g).to(dev)

    def step():
        opt.zero_grad()
        pre, fin = net(Frame=frame, Event=event, T=t, GTEx=None)
        loss = 0.1 * (lap(fin, gt) + census(fin, gt)) + (lap(pre, gt) + census(pre, gt))
        loss.backward()
        opt.step()
        return loss

    def measure():
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            loss = step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt)
        return ms, float(loss)

    grad_bytes = 4 * sum(p.numel() for p in model.parameters() if p.requires_grad)
    ms, loss = measure()
    res = {"workload": f"EVFIAutoEx(train_ours.yml) train, fp32 (cuDNN TF32 convolutions, torch default), B={batch}/GPU "
                       f"{size}x{size}, Adam, Lap+census loss, FAC at {size // 2}x{size // 2} "
                       f"({batch},1600,{size // 2},{size // 2}) kernel tensor; "
                       + (f"DDP over NCCL, {grad_bytes / 1e6:.1f} MB gradient all-reduce per step (enabled)" if world > 1 else "single GPU"),
           "ms_per_step": round(ms, 3), "samples_s": round(world * batch / (ms * 1e-3), 2),
           "Mpix_s": round(world * batch * size * size / 1e6 / (ms * 1e-3), 3), "n_gpus": world, "loss": round(loss, 4)}
    # the same step with the model's host round trip (Frame2Lap through OpenCV, one sync per forward) moved to the GPU
    orig = mod.Frame2Lap
    mod.Frame2Lap = frame_ops.Frame2Lap
    try:
        ms2, _ = measure()
    finally:
        mod.Frame2Lap = orig
    res["gpu_frame_map"] = {"ms_per_step": round(ms2, 3), "samples_s": round(world * batch / (ms2 * 1e-3), 2),
                            "Mpix_s": round(world * batch * size * size / 1e6 / (ms2 * 1e-3), 3)}
    if ab and world == 1 and refmodel.use_reference_cuda_fac(True):
        try:
            ms3, _ = measure()
            res["reference_fac_cuda_kernels_sm100a"] = {"ms_per_step": round(ms3, 3),
                                                        "samples_s": round(batch / (ms3 * 1e-3), 2)}
        finally:
            refmodel.use_reference_cuda_fac(False)
    del net, model, opt
    torch.cuda.empty_cache()
    return res
