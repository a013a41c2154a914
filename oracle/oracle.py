"""ctypes/numpy front end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs import this module; nothing under
ebfi-be_b200/ does. The arithmetic lives in ebfi_oracle.c / dcn_fac_oracle.inc,
each function citing the reference lines it restates.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


class DcnGeom(ctypes.Structure):
    """Mirror of `ebfi_dcn_geom` (include/ebfi_b200.h)."""

    _fields_ = [(n, ctypes.c_int) for n in (
        "batch", "channels", "height", "width", "channels_out", "kernel_h", "kernel_w",
        "stride_h", "stride_w", "pad_h", "pad_w", "dilation_h", "dilation_w", "deformable_group")]


def build(force=False):
    """Compile liboracle.so with the Makefile next to this file."""
    srcs = [os.path.join(_HERE, f) for f in ("ebfi_oracle.c", "dcn_fac_oracle.inc", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def dcn_geom(input_shape, weight_shape, stride, padding, dilation, deformable_groups):
    def pair(v):
        return (v, v) if isinstance(v, int) else tuple(v)
    (sh, sw), (ph, pw), (dh, dw) = pair(stride), pair(padding), pair(dilation)
    B, C, H, W = input_shape
    Co, _, kh, kw = weight_shape
    return DcnGeom(B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, deformable_groups)


def dcn_out_size(g):
    ho, wo = ctypes.c_int(), ctypes.c_int()
    if lib().oracle_dcn_out_size(ctypes.byref(g), ctypes.byref(ho), ctypes.byref(wo)):
        raise ValueError("inconsistent DCN geometry")
    return ho.value, wo.value


def dcn_forward(input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups,
                precision="f64"):
    """DCNv2 forward on numpy arrays (argument order of dcn_v2.py:17-21)."""
    input, offset, mask, weight, bias = map(_f32, (input, offset, mask, weight, bias))
    g = dcn_geom(input.shape, weight.shape, stride, padding, dilation, deformable_groups)
    ho, wo = dcn_out_size(g)
    assert offset.shape == (g.batch, 2 * g.deformable_group * g.kernel_h * g.kernel_w, ho, wo)
    assert mask.shape == (g.batch, g.deformable_group * g.kernel_h * g.kernel_w, ho, wo)
    out = np.empty((g.batch, g.channels_out, ho, wo), np.float32)
    fn = getattr(lib(), "oracle_dcn_forward_" + precision)
    rc = fn(ctypes.byref(g), _p(input), _p(weight), _p(bias), _p(offset), _p(mask), _p(out))
    if rc:
        raise RuntimeError("oracle_dcn_forward failed: %d" % rc)
    return out


def dcn_backward(input, offset, mask, weight, bias, grad_output, stride, padding, dilation,
                 deformable_groups, precision="f64"):
    """Returns (grad_input, grad_offset, grad_mask, grad_weight, grad_bias)."""
    input, offset, mask, weight, bias, grad_output = map(
        _f32, (input, offset, mask, weight, bias, grad_output))
    g = dcn_geom(input.shape, weight.shape, stride, padding, dilation, deformable_groups)
    gi, go, gm = np.empty_like(input), np.empty_like(offset), np.empty_like(mask)
    gw, gb = np.empty_like(weight), np.empty_like(bias)
    fn = getattr(lib(), "oracle_dcn_backward_" + precision)
    rc = fn(ctypes.byref(g), _p(input), _p(weight), _p(bias), _p(offset), _p(mask), _p(grad_output),
            _p(gi), _p(go), _p(gm), _p(gw), _p(gb))
    if rc:
        raise RuntimeError("oracle_dcn_backward failed: %d" % rc)
    return gi, go, gm, gw, gb


def fac_forward(input_pad, kernel, K, precision="f64"):
    input_pad, kernel = _f32(input_pad), _f32(kernel)
    B, C, Hi, Wi = input_pad.shape
    H, W = kernel.shape[2:]
    assert kernel.shape[1] == C * K * K and Hi - K == H - 1 and Wi - K == W - 1
    out = np.empty((B, C, H, W), np.float32)
    getattr(lib(), "oracle_fac_forward_" + precision)(
        _p(input_pad), _p(kernel), _p(out), B, C, H, W, K)
    return out


def fac_backward(input_pad, kernel, grad_output, K, precision="f64"):
    """Returns (grad_input, grad_kernel)."""
    input_pad, kernel, grad_output = _f32(input_pad), _f32(kernel), _f32(grad_output)
    B, C, _, _ = input_pad.shape
    H, W = kernel.shape[2:]
    gi, gk = np.empty_like(input_pad), np.empty_like(kernel)
    getattr(lib(), "oracle_fac_backward_" + precision)(
        _p(input_pad), _p(kernel), _p(grad_output), _p(gi), _p(gk), B, C, H, W, K)
    return gi, gk


def _coords(*arrs):
    """Give xs/ys/ts one common float dtype (fp32 stays fp32, anything else -> fp64)."""
    dt = 0 if all(np.asarray(a).dtype == np.float32 for a in arrs) else 1
    ty = np.float32 if dt == 0 else np.float64
    return dt, [np.array(a, dtype=ty, copy=True, order="C") for a in arrs]


def events_to_image(xs, ys, ps, sensor_size, binary=False):
    """Returns (img, xs_after, ys_after, ps_after): the reference mutates its inputs."""
    H, W = sensor_size
    dt, (xs, ys) = _coords(xs, ys)
    ps = np.array(ps, dtype=np.float32, copy=True)
    img = np.zeros((H, W), np.float32)
    fn = lib().oracle_events_to_mask if binary else lib().oracle_events_to_image
    fn(_p(xs), _p(ys), _p(ps), dt, ctypes.c_int64(len(xs)), H, W, _p(img))
    return img, xs, ys, ps


def events_to_voxel(xs, ys, ts, ps, num_bins, sensor_size):
    """Returns (voxel, xs_after, ys_after)."""
    H, W = sensor_size
    dt, (xs, ys, ts) = _coords(xs, ys, ts)
    ps = _f32(ps)
    vox = np.zeros((num_bins, H, W), np.float32)
    lib().oracle_events_to_voxel(_p(xs), _p(ys), _p(ts), _p(ps), dt, ctypes.c_int64(len(xs)),
                                 num_bins, H, W, _p(vox))
    return vox, xs, ys


def events_to_stack(xs, ys, ts, ps, B, sensor_size):
    """Returns (stack(2,B,H,W), xs_after, ys_after, bounds)."""
    H, W = sensor_size
    dt, (xs, ys, ts) = _coords(xs, ys, ts)
    ps = _f32(ps)
    stack = np.zeros((2, B, H, W), np.float32)
    bounds = np.zeros(2 * B, np.int64)
    if float(ts.sum()) == 0 or len(ts) <= 3:          # encodings.py:319-320
        return stack, xs, ys, bounds
    lib().oracle_events_to_stack(_p(xs), _p(ys), _p(ts), _p(ps), dt, ctypes.c_int64(len(xs)),
                                 B, H, W, _p(stack), _p(bounds))
    return stack, xs, ys, bounds


def dataset_event_stack(xs, ys, ts, ps, B, sensor_size):
    """H5Dataset.GetEvents (dataloader/h5dataset.py:327-349) for one slice in the on-disk dtypes (int16,
    int16, float64 seconds, int8; generate_dataset/tools/event_packagers.py:128-131): empty-slice
    substitute (:332-333), float64 normalisation (:335), everything promoted to float64 by the concatenate
    (:336), `ps.float()` and events_to_stack, `.transpose(0, 1)` (:349). Returns (B, 2, H, W)."""
    xs, ys, ts, ps = (np.asarray(a) for a in (xs, ys, ts, ps))
    if len(xs) == 0 or len(ys) == 0 or len(ts) == 0 or len(ps) == 0:
        xs = ys = ts = ps = np.array([0.])
    ts = (ts.astype(np.float64) - ts[0]) / (ts[-1] - ts[0] + 1e-6)
    stack, *_ = events_to_stack(xs.astype(np.float64), ys.astype(np.float64), ts, ps.astype(np.float32), B, sensor_size)
    return np.ascontiguousarray(stack.transpose(1, 0, 2, 3))


def _bf16_round(a):
    """Round-to-nearest-even to bfloat16, returned as float32 (numpy has no bf16)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(a))


def kpn_fused_forward(event_feat, frame_feat, conv_weight, conv_bias, K, negative_slope=0.01, bf16_operands=False):
    """models/Ours/model_singleframe.py:145-146,159-162 (class Modification) restated in float64:
        Kernel = LeakyReLU(conv3x3(cat([Event, Frame], 1)), pad 1)      # ConvLayer: nn.Conv2d + nn.LeakyReLU()
        out    = KernelConv2D(K)(Event, Kernel)                          # ReplicationPad2d((K-1)/2) + FAC
    bf16_operands=True rounds the convolution's inputs and weights to bf16 first (what the fused kernel feeds the
    tensor cores); everything else stays float64. Returns (out float64, Kernel float64)."""
    ev = np.asarray(event_feat, np.float64)
    feat = np.concatenate([np.asarray(event_feat, np.float32), np.asarray(frame_feat, np.float32)], axis=1)
    w = np.asarray(conv_weight, np.float32)
    if bf16_operands:
        feat, w = _bf16_round(feat), _bf16_round(w)
    feat, w = feat.astype(np.float64), w.astype(np.float64)
    B, Cin, H, W = feat.shape
    Cout = w.shape[0]
    fp = np.pad(feat, ((0, 0), (0, 0), (1, 1), (1, 1)))
    cols = np.empty((B, Cin * 9, H * W))
    for c in range(Cin):
        for i in range(3):
            for j in range(3):
                cols[:, c * 9 + i * 3 + j] = fp[:, c, i:i + H, j:j + W].reshape(B, -1)
    ker = np.einsum("ok,bkp->bop", w.reshape(Cout, Cin * 9), cols) + np.asarray(conv_bias, np.float64)[None, :, None]
    ker = np.where(ker > 0, ker, ker * negative_slope).reshape(B, Cout, H, W)
    R = (K - 1) // 2
    evp = np.pad(ev, ((0, 0), (0, 0), (R, R), (R, R)), mode="edge")
    Ce = ev.shape[1]
    out = np.zeros((B, Ce, H, W))
    for c in range(Ce):
        for ky in range(K):
            for kx in range(K):
                out[:, c] += evp[:, c, ky:ky + H, kx:kx + W] * ker[:, c * K * K + ky * K + kx]
    return out, ker


def frame_to_lap(ims):
    """Frame2Lap (myutils/utils.py:34-49) restated with the arithmetic OpenCV >= 4 uses: uint8 cast of im*255,
    15-bit fixed-point BGR2GRAY, 4-neighbour Laplacian with BORDER_REFLECT_101. (B,3,H,W) -> (B,1,H,W) float32."""
    ims = np.asarray(ims, np.float32)
    u8 = (ims * np.float32(255)).astype(np.int64) & 0xFF
    gray = (3735 * u8[:, 0] + 19235 * u8[:, 1] + 9798 * u8[:, 2] + (1 << 14)) >> 15
    H, W = gray.shape[1:]
    ry = np.array([_reflect101(i, H) for i in range(-1, H + 1)])
    rx = np.array([_reflect101(i, W) for i in range(-1, W + 1)])
    p = gray[:, ry][:, :, rx]
    lap = p[:, :-2, 1:-1] + p[:, 2:, 1:-1] + p[:, 1:-1, :-2] + p[:, 1:-1, 2:] - 4 * gray
    return lap[:, None].astype(np.float32)


def _reflect101(i, n):
    if n == 1:
        return 0
    while i < 0 or i >= n:
        i = -i if i < 0 else 2 * n - 2 - i
    return i


def frame_to_dcp(ims, sz=35):
    """Frame2DCP (myutils/utils.py:15-31): channel minimum, then cv2.erode with a sz x sz rectangle = window minimum
    with anchor sz//2, pixels outside the image ignored. (B,3,H,W) -> (B,1,H,W) float32."""
    ims = np.asarray(ims, np.float32)
    dc = ims.min(axis=1)
    a = sz // 2
    pad = np.pad(dc, ((0, 0), (a, sz - 1 - a), (a, sz - 1 - a)), constant_values=np.inf)
    from numpy.lib.stride_tricks import sliding_window_view
    return sliding_window_view(pad, (sz, sz), axis=(1, 2)).min(axis=(3, 4))[:, None].astype(np.float32)
