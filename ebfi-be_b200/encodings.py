"""Host-side mirror of the reference's event encoders (dataloader/encodings.py), running as
ONE scatter kernel per call on the GPU instead of B (or 2B) CPU index_put_ passes.

Same function names, argument order and return shapes as the reference:
    events_to_image   (encodings.py:243-268)     -> (H, W)
    events_to_voxel   (:271-286)                 -> (num_bins, H, W)
    events_to_channels(:289-304)                 -> (2, H, W)
    events_to_stack   (:307-350)                 -> (2, B, H, W)   [datasets transpose to (B,2,H,W)]
    events_to_mask    (:353-377)                 -> (H, W)
Inputs are CUDA tensors (xs, ys, ts: float32 or float64, like h5dataset.py:327-349 builds;
ps: float32). The reference mutates xs / ys / ps of out-of-range events in place (:254-256);
so do these functions when the tensors can be written directly (contiguous, right dtype).
"""
import torch

from . import _lib as L


def _work(t, want):
    """Tensor the kernel can read and write in place: `t` itself when possible, else a converted copy."""
    L.require_cuda(t)
    return t if (t.dtype == want and t.is_contiguous()) else t.to(want).contiguous()


def _prep(xs, ys, ts=None):
    """Common float dtype for xs / ys (/ ts): float32 if all are float32, else float64."""
    group = (xs, ys) if ts is None else (xs, ys, ts)
    want = torch.float32 if all(t.dtype == torch.float32 for t in group) else torch.float64
    return (L.EBFI_F32 if want == torch.float32 else L.EBFI_F64), [_work(t, want) for t in group]


def _sync_back(orig, work):
    """Propagate the kernel's in-place zeroing (encodings.py:254-256) to the caller's tensor
    when a converted copy had to be used."""
    if work is not orig:
        orig.copy_(work)


def _ps(ps):
    return _work(ps, torch.float32)


def events_to_image(xs, ys, ps, sensor_size=(180, 240)):
    """Accumulate events into an image (encodings.py:243-268)."""
    H, W = sensor_size
    dt, (x, y) = _prep(xs, ys)
    p = _ps(ps)
    with torch.cuda.device(x.device):
        img = torch.zeros((H, W), dtype=torch.float32, device=x.device)
        L.check(L.load().ebfi_events_to_image(L.stream_ptr(x.device), L.ptr(x), L.ptr(y), L.ptr(p), dt,
                                              x.numel(), H, W, L.ptr(img), 1), "events_to_image")
    _sync_back(xs, x), _sync_back(ys, y), _sync_back(ps, p)
    return img


def events_to_mask(xs, ys, ps, sensor_size=(180, 240)):
    """Binary event mask, last event of a pixel wins (encodings.py:353-377)."""
    H, W = sensor_size
    dt, (x, y) = _prep(xs, ys)
    p = _ps(ps)
    with torch.cuda.device(x.device):
        img = torch.zeros((H, W), dtype=torch.float32, device=x.device)
        scratch = torch.empty((H, W), dtype=torch.int64, device=x.device)
        L.check(L.load().ebfi_events_to_mask(L.stream_ptr(x.device), L.ptr(x), L.ptr(y), L.ptr(p), dt,
                                             x.numel(), H, W, L.ptr(img), L.ptr(scratch), 1), "events_to_mask")
    _sync_back(xs, x), _sync_back(ys, y), _sync_back(ps, p)
    return img


def events_to_voxel(xs, ys, ts, ps, num_bins, sensor_size=(180, 240)):
    """Temporal-bilinear voxel grid (encodings.py:271-286); ts is expected in [0, 1]."""
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    H, W = sensor_size
    dt, (x, y, t) = _prep(xs, ys, ts)
    p = _ps(ps)
    with torch.cuda.device(x.device):
        vox = torch.zeros((num_bins, H, W), dtype=torch.float32, device=x.device)
        L.check(L.load().ebfi_events_to_voxel(L.stream_ptr(x.device), L.ptr(x), L.ptr(y), L.ptr(t), L.ptr(p),
                                              dt, x.numel(), num_bins, H, W, L.ptr(vox), 1), "events_to_voxel")
    _sync_back(xs, x), _sync_back(ys, y)
    return vox


def events_to_channels(xs, ys, ps, sensor_size=(180, 240)):
    """Two-channel positive / negative event counters (encodings.py:289-304)."""
    assert len(xs) == len(ys) and len(ys) == len(ps)
    p = _ps(ps)
    pos = events_to_image(xs, ys, p * p.clamp(min=0), sensor_size=sensor_size)
    # xs / ys of out-of-range events are now (0, 0), exactly as in the reference's second call
    neg = events_to_image(xs, ys, p * p.clamp(max=0), sensor_size=sensor_size)
    return torch.stack([pos, neg])


def events_to_stack(xs, ys, ts, ps, B, sensor_size=(180, 240)):
    """(2, B, H, W) per-bin positive / negative counts (encodings.py:307-350)."""
    H, W = sensor_size
    L.require_cuda(ts)
    if len(ts) <= 3:                                   # encodings.py:319-320
        return torch.zeros([2, B, H, W], device=ts.device)
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    dt, (x, y, t) = _prep(xs, ys, ts)
    p = _ps(ps)
    with torch.cuda.device(x.device):
        # the other half of the reference's early-out, `ts.sum() == 0`, stays on the device: the
        # kernels read the flag, the host never waits for it
        skip = (t.sum() == 0).to(torch.uint8)
        stack = torch.zeros((2, B, H, W), dtype=torch.float32, device=x.device)
        bounds = torch.empty(2 * B, dtype=torch.int64, device=x.device)
        L.check(L.load().ebfi_events_to_stack(L.stream_ptr(x.device), L.ptr(x), L.ptr(y), L.ptr(t), L.ptr(p),
                                              dt, x.numel(), B, H, W, L.ptr(stack), L.ptr(bounds), 1, L.ptr(skip)),
                "events_to_stack")
    _sync_back(xs, x), _sync_back(ys, y)
    return stack
