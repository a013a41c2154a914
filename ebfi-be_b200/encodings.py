"""Host-side mirror of the reference's event encoders (dataloader/encodings.py), running as
ONE scatter kernel per call on the GPU instead of B (or 2B) CPU index_put_ passes.

Same function names, argument order and return shapes as the reference:
    events_to_image   (encodings.py:243-268)     -> (H, W)
    events_to_voxel   (:271-286)                 -> (num_bins, H, W)
    events_to_channels(:289-304)                 -> (2, H, W)
    events_to_stack   (:307-350)                 -> (2, B, H, W)   [datasets transpose to (B,2,H,W)]
    events_to_mask    (:353-377)                 -> (H, W)
    events_raw_to_stack: H5Dataset.GetEvents (dataloader/h5dataset.py:327-349) on the on-disk dtypes
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
        scratch = torch.empty(4800 + 8, dtype=torch.uint8, device=x.device)      # EBFI_EVENTS_SUM_SCRATCH_BYTES + the flag
        skip = scratch[4800:4801]
        L.check(L.load().ebfi_events_ts_sum_is_zero(L.stream_ptr(x.device), L.ptr(t), dt, t.numel(), L.ptr(scratch), L.ptr(skip)),
                "events_ts_sum_is_zero")
        stack = torch.zeros((2, B, H, W), dtype=torch.float32, device=x.device)
        bounds = torch.empty(2 * B, dtype=torch.int64, device=x.device)
        L.check(L.load().ebfi_events_to_stack(L.stream_ptr(x.device), L.ptr(x), L.ptr(y), L.ptr(t), L.ptr(p),
                                              dt, x.numel(), B, H, W, L.ptr(stack), L.ptr(bounds), 1, L.ptr(skip)),
                "events_to_stack")
    _sync_back(xs, x), _sync_back(ys, y)
    return stack


def events_raw_to_stack(xs, ys, ts, ps, B, sensor_size=(180, 240), bins_major=True, out=None):
    """The datasets' event path (dataloader/h5dataset.py:327-349) on the on-disk dtypes of the HDF5 files
    (generate_dataset/tools/event_packagers.py:128-131): xs, ys int16; ts float64 seconds (non-decreasing);
    ps int8 — CUDA tensors. Normalisation of ts (:335), `ps.float()` and events_to_stack happen inside one
    pass; nothing is converted to float64 first (13 instead of 32 bytes per event). Returns (B, 2, H, W),
    i.e. the `.transpose(0, 1)` the datasets apply (:349), or the reference's (2, B, H, W) with
    bins_major=False. `out`: optional zero-filled result tensor to accumulate into."""
    H, W = sensor_size
    L.require_cuda(xs, ys, ts, ps)
    want = (torch.int16, torch.int16, torch.float64, torch.int8)
    for name, t, dt in zip(("xs", "ys", "ts", "ps"), (xs, ys, ts, ps), want):
        if t.dtype != dt or t.dim() != 1:
            raise RuntimeError(f"events_raw_to_stack: {name} must be a 1-D {dt} tensor, got {t.dtype} {tuple(t.shape)}")
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    xs, ys, ts, ps = (t.contiguous() for t in (xs, ys, ts, ps))
    shape = (B, 2, H, W) if bins_major else (2, B, H, W)
    with torch.cuda.device(xs.device):
        if out is None:
            out = torch.zeros(shape, dtype=torch.float32, device=xs.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
            raise RuntimeError(f"events_raw_to_stack: out must be a contiguous float32 {shape} tensor")
        bounds = torch.empty(2 * B, dtype=torch.int64, device=xs.device)
        L.check(L.load().ebfi_events_raw_to_stack(L.stream_ptr(xs.device), L.ptr(xs), L.ptr(ys), L.ptr(ts), L.ptr(ps),
                                                  xs.numel(), B, H, W, L.ptr(out), L.ptr(bounds), int(bins_major)),
                "events_raw_to_stack")
    return out


class EventSliceFeeder:
    """Host side of `events_raw_to_stack` for a DataLoader-style producer: takes the four numpy arrays of
    one or more HDF5 slices (`h5_file['ori_events/xs'][i0:i1]`, ... — h5dataset.py:328-331), stages them in
    reusable pinned buffers, copies them on a side stream and encodes every slice into one
    (n_slices, B, 2, H, W) batch. The compute stream only waits for the copies; the host never waits for
    the GPU unless a staging buffer still in flight has to be reused."""

    def __init__(self, device, B, sensor_size, max_events=1 << 20):
        self.device, self.B, self.sensor = torch.device(device), B, tuple(sensor_size)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._cap = 0
        self._slot = 0
        self._stage = [None, None]             # two staging sets: fill one while the other is in flight
        self._reserve(max_events)

    def _reserve(self, n):
        if n <= self._cap:
            return
        self._cap = n
        for i in range(2):
            if self._stage[i] is not None and self._stage[i]["done"] is not None:
                self._stage[i]["done"].synchronize()
            self._stage[i] = {
                "host": tuple(torch.empty(n, dtype=dt).pin_memory() for dt in (torch.int16, torch.int16, torch.float64, torch.int8)),
                "dev": tuple(torch.empty(n, dtype=dt, device=self.device) for dt in (torch.int16, torch.int16, torch.float64, torch.int8)),
                "done": None}

    def encode(self, slices):
        """slices: sequence of (xs, ys, ts, ps) numpy arrays / CPU tensors in the on-disk dtypes."""
        total = sum(len(s[0]) for s in slices)
        self._reserve(max(total, 1))
        st = self._stage[self._slot]
        self._slot ^= 1
        if st["done"] is not None:
            st["done"].synchronize()           # the encode that last used this staging set has finished
        spans, pos = [], 0
        for s in slices:
            n = len(s[0])
            for h, a in zip(st["host"], s):
                h[pos:pos + n].copy_(torch.as_tensor(a))
            spans.append((pos, n))
            pos += n
        comp = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            for h, d in zip(st["host"], st["dev"]):
                d[:total].copy_(h[:total], non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self.copy_stream)
        comp.wait_event(copied)
        H, W = self.sensor
        with torch.cuda.device(self.device):
            out = torch.zeros((len(slices), self.B, 2, H, W), dtype=torch.float32, device=self.device)
            for i, (p0, n) in enumerate(spans):
                events_raw_to_stack(*(d[p0:p0 + n] for d in st["dev"]), self.B, self.sensor, True, out[i])
            st["done"] = torch.cuda.Event()
            st["done"].record(comp)
        return out
