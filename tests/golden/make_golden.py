"""Generate the golden vectors in tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference and oracle/_ref/dcn_cpu,
see oracle/build_ref.py):   python tests/golden/make_golden.py

What each file pins, and with what:
  dcn_*.npz     backward: the reference's own CPU build (dcn_v2_cpu_backward compiled
                unmodified from models/DCNv2/src/cpu/*.cpp). forward: the reference CPU
                forward returns uninitialised memory (dcn_v2_cpu.cpp:65,127 never writes
                `output`), so the forward golden is torchvision.ops.deform_conv2d, which the
                survey probe found bit-identical to the CUDA semantics (dcn_v2_cuda.cu:69).
  dcn_zero_offset.npz   the reference's known-answer test, testcuda.py:32-67
                (identity centre-tap weights, zero offsets, mask 0.5  =>  2*out == input).
  fac_*.npz     the reference FAC op has no CPU path (KernelConv2D.py:38-39): golden =
                F.unfold restatement + torch autograd, shapes from the reference's
                gradient_check recipe (KernelConv2D.py:61-74). On the GPU box the tests also
                compare against the reference's own CUDA kernels (oracle/_ref/fac_cuda).
  events_*.npz  dataloader/encodings.py imported unchanged from /root/reference.
"""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "dcn_cpu"))
import _ext as ref_dcn  # noqa: E402  (reference CPU build)

spec = importlib.util.spec_from_file_location("ref_encodings", os.path.join(REF, "dataloader/encodings.py"))
enc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(enc)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrs.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


def dcn_case(name, seed, B, C, Co, H, W, k, stride, pad, dil, dg, off_scale=2.0, off_bias=0.0):
    g = torch.Generator().manual_seed(seed)
    (kh, kw), (sh, sw), (ph, pw), (dh, dw) = k, stride, pad, dil
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    x = torch.randn(B, C, H, W, generator=g)
    w = (torch.rand(Co, C, kh, kw, generator=g) * 2 - 1) / (C * kh * kw) ** 0.5
    b = torch.randn(Co, generator=g)
    off = torch.randn(B, 2 * dg * kh * kw, Ho, Wo, generator=g) * off_scale + off_bias
    msk = torch.sigmoid(torch.randn(B, dg * kh * kw, Ho, Wo, generator=g))
    go = torch.randn(B, Co, Ho, Wo, generator=g)
    gi, goff, gmsk, gw, gb = ref_dcn.dcn_v2_backward(x, w, b, off, msk, go, kh, kw, sh, sw, ph, pw, dh, dw, dg)
    out = torchvision.ops.deform_conv2d(x, off, w, b, stride=(sh, sw), padding=(ph, pw),
                                        dilation=(dh, dw), mask=msk)
    save(name, input=x, weight=w, bias=b, offset=off, mask=msk, grad_output=go,
         geom=np.array([kh, kw, sh, sw, ph, pw, dh, dw, dg]), output=out, grad_input=gi,
         grad_offset=goff, grad_mask=gmsk, grad_weight=gw, grad_bias=gb)


def dcn_zero_offset():
    # testcuda.py:14-67: N=2, inC=outC=2, 4x4, 3x3, dg=1
    g = torch.Generator().manual_seed(7)
    N, C, H, W = 2, 2, 4, 4
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.zeros(C, C, 3, 3)
    for p in range(C):
        w[p, p, 1, 1] = 1.0
    save("dcn_zero_offset", input=x, weight=w, bias=torch.zeros(C), offset=torch.zeros(N, 18, H, W),
         mask=torch.full((N, 9, H, W), 0.5), geom=np.array([3, 3, 1, 1, 1, 1, 1, 1, 1]),
         output=x * 0.5)


def fac_case(name, seed, B, C, K, H, W):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H + K - 1, W + K - 1, generator=g, requires_grad=True)
    ker = torch.randn(B, C * K * K, H, W, generator=g, requires_grad=True)
    go = torch.randn(B, C, H, W, generator=g)
    out = (F.unfold(x, K).view(B, C, K * K, H, W) * ker.view(B, C, K * K, H, W)).sum(2)
    out.backward(go)
    save(name, input=x, kernel=ker, grad_output=go, K=np.array(K), output=out,
         grad_input=x.grad, grad_kernel=ker.grad)


def event_cloud(seed, n, H, W, oob=0, dup_ts=False, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    xs = torch.randint(0, W, (n,), generator=g).to(dtype)
    ys = torch.randint(0, H, (n,), generator=g).to(dtype)
    ts = torch.sort(torch.rand(n, generator=g, dtype=torch.float64))[0]
    if dup_ts:  # quantise so that many timestamps repeat and some land on bin boundaries
        ts = torch.round(ts * 32) / 32
    ts = ((ts - ts[0]) / (ts[-1] - ts[0] + 1e-6)).to(dtype)
    ps = (torch.randint(0, 2, (n,), generator=g) * 2 - 1).float()
    for j in range(oob):  # sprinkle out-of-range coordinates (negative, too large, fractional)
        i = int(torch.randint(0, n, (1,), generator=g))
        if j % 4 == 0: xs[i] = W + j
        elif j % 4 == 1: ys[i] = -1.0 - j
        elif j % 4 == 2: xs[i] = -0.5
        else: ys[i] = float(H)
    return xs, ys, ts, ps


def events_cases():
    H, W = 12, 16
    for name, kw in [("events_plain", dict(seed=1, n=500)),
                     ("events_oob", dict(seed=2, n=400, oob=24)),
                     ("events_dupts", dict(seed=3, n=600, oob=8, dup_ts=True)),
                     ("events_f64", dict(seed=4, n=300, oob=6, dup_ts=True, dtype=torch.float64))]:
        xs, ys, ts, ps = event_cloud(H=H, W=W, **kw)
        out = dict(xs=xs, ys=ys, ts=ts, ps=ps, sensor=np.array([H, W]))
        a = [t.clone() for t in (xs, ys, ps)]
        out["image"] = enc.events_to_image(*a, sensor_size=(H, W))
        out["image_xs"], out["image_ys"], out["image_ps"] = a
        a = [t.clone() for t in (xs, ys, ps)]
        out["mask"] = enc.events_to_mask(*a, sensor_size=(H, W))
        a = [t.clone() for t in (xs, ys, ps)]
        out["channels"] = enc.events_to_channels(*a, sensor_size=(H, W))
        if xs.dtype == torch.float32:   # fp64 ts makes index_put_ reject the fp64 values
            a = [t.clone() for t in (xs, ys, ts, ps)]
            out["voxel5"] = enc.events_to_voxel(*a, 5, sensor_size=(H, W))
            out["voxel_xs"], out["voxel_ys"] = a[0], a[1]
        for nb in (4, 16):
            a = [t.clone() for t in (xs, ys, ts, ps)]
            out[f"stack{nb}"] = enc.events_to_stack(*a, nb, sensor_size=(H, W))
            out[f"stack{nb}_xs"], out[f"stack{nb}_ys"] = a[0], a[1]
        save(name, **out)
    # degenerate inputs, encodings.py:319-320
    z = torch.zeros(5)
    save("events_degenerate",
         stack_tssum0=enc.events_to_stack(z.clone(), z.clone(), z.clone(), torch.ones(5), 3, sensor_size=(4, 4)),
         stack_len3=enc.events_to_stack(torch.tensor([1., 2, 3]), torch.tensor([1., 2, 3]),
                                        torch.tensor([0., .5, 1]), torch.ones(3), 3, sensor_size=(4, 4)))


def raw_event_slice(seed, n, H, W, oob=0, quant=None, ps01=False):
    """An HDF5 event slice in the on-disk dtypes (generate_dataset/tools/event_packagers.py:128-131)."""
    g = np.random.default_rng(seed)
    xs = g.integers(0, W, n).astype(np.int16)
    ys = g.integers(0, H, n).astype(np.int16)
    ts = np.sort(g.random(n))
    if quant:                                   # repeated timestamps, some exactly on bin boundaries
        ts = np.round(ts * quant) / quant
    ts = (1234.5 + 0.25 * ts).astype(np.float64)    # seconds since the start of the recording
    ps = (g.integers(0, 2, n) if ps01 else g.integers(0, 2, n) * 2 - 1).astype(np.int8)
    for j in range(oob):
        i = int(g.integers(0, n))
        if j % 4 == 0: xs[i] = W + j
        elif j % 4 == 1: ys[i] = -1 - j
        elif j % 4 == 2: xs[i] = -1
        else: ys[i] = H
    return xs, ys, ts, ps


def dataset_event_stack(xs, ys, ts, ps, bins, sensor):
    """H5Dataset.GetEvents (dataloader/h5dataset.py:327-349) on one slice. h5dataset.py itself needs h5py
    (absent here), so its five array lines are restated; the encoder is the reference's own."""
    if len(xs) == 0 or len(ys) == 0 or len(ts) == 0 or len(ps) == 0:          # :332-333
        xs = ys = ts = ps = np.array([0.])
    ts = (ts - ts[0]) / (ts[-1] - ts[0] + 1e-6)                                  # :335
    ev = torch.from_numpy(np.concatenate((xs[np.newaxis, ...], ys[np.newaxis, ...], ts[np.newaxis, ...],
                                          ps[np.newaxis, ...]), axis=0))        # :336, 4xN float64
    return enc.events_to_stack(xs=ev[0], ys=ev[1], ts=ev[2], ps=ev[3].float(), B=bins,
                               sensor_size=sensor).transpose(0, 1)              # :349, TBx2xHxW


def events_raw_cases():
    H, W = 12, 16
    out = {"sensor": np.array([H, W])}
    for name, kw in [("plain", dict(seed=31, n=700)),
                     ("dupts_oob", dict(seed=32, n=600, oob=24, quant=64)),
                     ("ps01", dict(seed=33, n=300, oob=4, quant=16, ps01=True)),
                     ("len3", dict(seed=34, n=3)),
                     ("len4", dict(seed=35, n=4)),
                     ("empty", dict(seed=36, n=0))]:
        xs, ys, ts, ps = raw_event_slice(H=H, W=W, **kw)
        for k, v in zip(("xs", "ys", "ts", "ps"), (xs, ys, ts, ps)):
            out[f"{name}_{k}"] = v
        for nb in (4, 16):
            out[f"{name}_stack{nb}"] = dataset_event_stack(xs, ys, ts, ps, nb, (H, W)).contiguous()
    xs, ys, ts, ps = raw_event_slice(seed=37, n=50, H=H, W=W)
    ts[:] = 99.0                                                                 # one instant: ts.sum() == 0
    for k, v in zip(("xs", "ys", "ts", "ps"), (xs, ys, ts, ps)):
        out[f"same_ts_{k}"] = v
    out["same_ts_stack4"] = dataset_event_stack(xs, ys, ts, ps, 4, (H, W)).contiguous()
    out["same_ts_stack16"] = dataset_event_stack(xs, ys, ts, ps, 16, (H, W)).contiguous()
    save("events_raw", **out)


def frame_cases():
    """myutils/utils.py:15-49 (Frame2DCP, Frame2Lap) executed with the OpenCV of this image. The module itself imports
    the trainer's data-list code and calls .cuda(); its two function bodies are restated without the device move."""
    import cv2

    def Frame2DCP(ims, sz=35):                                            # utils.py:15-31
        out = []
        for i in range(ims.size(0)):
            im = ims[i].permute(1, 2, 0).cpu().numpy()
            b, g, r = cv2.split(im)
            dc = cv2.min(cv2.min(r, g), b)
            kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (sz, sz))
            out.append(torch.from_numpy(cv2.erode(dc, kernel)).unsqueeze(0))
        return torch.stack(out, dim=0)

    def Frame2Lap(ims):                                                   # utils.py:34-49
        out = []
        for i in range(ims.size(0)):
            im = ims[i].permute(1, 2, 0).cpu().numpy()
            im = (im * 255).astype(np.uint8)
            gray = cv2.cvtColor(im, cv2.COLOR_BGR2GRAY)
            out.append(torch.from_numpy(cv2.Laplacian(gray, cv2.CV_64F).astype('float32')).unsqueeze(0))
        return torch.stack(out, dim=0)

    g = torch.Generator().manual_seed(41)
    frames = torch.rand(2, 3, 45, 52, generator=g)
    frames[0, :, 0, 0] = 1.0                                              # exactly 255 after the cast
    frames[1, :, 10:14, 20:30] = 0.0
    tiny = torch.rand(1, 3, 1, 7, generator=g)                            # single row: reflect-101 degenerates
    save("frames", cv2_version=np.array(cv2.__version__), frames=frames, lap=Frame2Lap(frames), dcp35=Frame2DCP(frames),
         dcp4=Frame2DCP(frames, 4), dcp1=Frame2DCP(frames, 1), tiny=tiny, tiny_lap=Frame2Lap(tiny), tiny_dcp=Frame2DCP(tiny, 35))


if __name__ == "__main__":
    torch.set_num_threads(1)
    dcn_zero_offset()
    # gradcheck-recipe shape of testcuda.py:69-97 (N=2, C=2, 4x4, dg=1), input*0.01 scale not needed here
    dcn_case("dcn_small_dg1", 11, 2, 2, 2, 4, 4, (3, 3), (1, 1), (1, 1), (1, 1), 1)
    dcn_case("dcn_dg4_stride2", 12, 1, 8, 6, 12, 10, (3, 3), (2, 2), (1, 1), (1, 1), 4)
    dcn_case("dcn_dil2_bigoff", 13, 1, 6, 4, 10, 11, (3, 3), (1, 1), (2, 2), (2, 2), 3, off_scale=6.0)
    dcn_case("dcn_pad_h_ne_w", 14, 2, 4, 4, 9, 8, (3, 3), (1, 1), (2, 1), (1, 1), 2)   # pad_h/pad_h quirk
    dcn_case("dcn_k1", 15, 1, 4, 4, 7, 9, (1, 1), (1, 1), (0, 0), (1, 1), 2)
    dcn_case("dcn_border", 16, 1, 4, 4, 8, 8, (3, 3), (1, 1), (1, 1), (1, 1), 2, off_scale=0.3, off_bias=-4.0)
    dcn_case("dcn_c64_dg8", 17, 1, 64, 64, 16, 16, (3, 3), (1, 1), (1, 1), (1, 1), 8)
    fac_case("fac_k5", 21, 2, 3, 5, 9, 12)
    fac_case("fac_k3", 22, 3, 4, 3, 8, 10)
    fac_case("fac_k1", 23, 1, 5, 1, 10, 8)
    fac_case("fac_k5_odd", 24, 1, 2, 5, 7, 13)
    events_cases()
    events_raw_cases()
    frame_cases()
