"""Data-parallel plumbing for the alignment ops: one process per GPU, torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests).

The three operators shard by batch with no data-path exchange: samples are
independent in DCNv2 (the reference itself loops over the batch, dcn_v2_cuda.cu:150), FAC has no
parameters, and every rank encodes its own event windows like the reference's DataLoader workers.
The only cross-sample reduction on the path is DCNv2's grad_weight / grad_bias (dcn_v2_cuda.cu:203-208);
in data-parallel training that is one small all-reduce per step, bucketed here into a single flat
buffer (147 KB + 256 B at the benchmark shape). The reference never all-reduces gradients at all
(every backward runs under `no_sync`, train_ours.py:250-272) — this is the correct behaviour
BASELINE.json asks for, not parity with that bug.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items, rank=None, world_size=None):
    """Contiguous [start, end) slice of `n_items` batch entries owned by `rank`; the first
    n_items % world_size ranks get one extra entry, so every entry is owned exactly once."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(*tensors, rank=None, world_size=None):
    """Slice every tensor along dim 0 to this rank's shard."""
    out = []
    for t in tensors:
        s, e = shard_range(t.shape[0], rank, world_size)
        out.append(t[s:e])
    return out if len(out) > 1 else out[0]


class _BucketWork:
    """Handle of an in-flight flat-bucket all-reduce; `wait()` orders the current stream after the collective
    and copies the reduced values back into the gradient tensors."""

    def __init__(self, grads, flat, work, scale):
        self.grads, self.flat, self.work, self.scale = grads, flat, work, scale

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
            if self.scale != 1.0:
                self.flat *= self.scale
            o = 0
            for g in self.grads:
                n = g.numel()
                g.copy_(self.flat[o:o + n].view_as(g))
                o += n
        return self.grads


def allreduce_weight_grads(grads, group=None, average=False, async_op=False):
    """Sum (or average) parameter gradients across ranks with ONE collective: the tensors are packed
    into a flat bucket, all-reduced, and copied back in place. Deterministic for a fixed world size
    (NCCL/gloo ring order is fixed). Returns the list it was given — or, with async_op=True, a handle
    whose `wait()` does the copy-back: the collective (147 KB: pure latency, ~0.15 ms on 8 GPUs) then runs on
    NCCL's stream underneath whatever is launched in between (the FAC kernels of the same step)."""
    grads = [g for g in grads if g is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return _BucketWork(grads, None, None, 1.0) if async_op else grads
    flat = torch.cat([g.reshape(-1) for g in grads])
    scale = 1.0 / dist.get_world_size(group) if average else 1.0
    handle = _BucketWork(grads, flat, dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), scale)
    return handle if async_op else handle.wait()


class GradComm:
    """Peer-memory communicator for the fused weight-gradient all-reduce (include/ebfi_b200.h `ebfi_dp_comm`,
    csrc/dp_comm.cuh): one symmetric allocation per rank, mapped into every peer by torch's symmetric memory
    (plumbing only — the exchange itself runs inside this repo's kernels over NVLink, no NCCL call per step).
    `n_floats` = the largest number of values all-reduced by one call (DCNv2: weight.numel() + bias.numel()).
    Collective constructor: every rank of `group` must create it at the same point. CUDA + NCCL group only."""

    def __init__(self, n_floats, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("GradComm: at most 8 ranks (one NVSwitch domain)")
        nbytes = int(L.load().ebfi_dp_comm_bytes(int(n_floats)))
        self.n_floats = int(n_floats)
        self.buf = symm.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)
        self.buf.zero_()                                    # epoch counter and flags start at 0
        self.handle = symm.rendezvous(self.buf, group.group_name if hasattr(group, "group_name") else group)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                 # every rank's zero fill is done before anyone signals
        ptrs = list(self.handle.buffer_ptrs)
        self.struct = L.DpComm(self.world, self.rank, (L.c_void * 8)(*(ptrs + [None] * (8 - len(ptrs)))), nbytes)

    def _call(self, fn, what, a, b):
        from . import _lib as L
        n = a.numel() + (b.numel() if b is not None else 0)
        if n > self.n_floats:
            raise RuntimeError(f"GradComm: {n} values > capacity {self.n_floats}")
        for t in (a, b):
            if t is not None and not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError("GradComm.allreduce_: contiguous float32 CUDA tensors only")
        L.check(getattr(L.load(), fn)(L.stream_ptr(a.device), self.struct, L.ptr(a), a.numel(),
                                      L.ptr(b) if b is not None else None, b.numel() if b is not None else 0), what)
        return a, b

    def allreduce_(self, a, b=None):
        """In place: `a` (and `b`) become the sums over all ranks — one kernel of this repo, on the current stream."""
        return self._call("ebfi_dp_allreduce_sum", "dp_allreduce_sum", a, b)

    def publish(self, a, b=None):
        """First half of the exchange: hand this rank's values to the peers, do not wait."""
        return self._call("ebfi_dp_publish", "dp_publish", a, b)

    def complete(self, a, b=None):
        """Second half: wait for every peer's values of the last publish (this class's, or a deferred
        `_ext.dcn_v2_backward(..., comm=, defer=True)`); `a` / `b` become the rank-ordered sums."""
        return self._call("ebfi_dp_complete", "dp_complete", a, b)


def dcn_backward_data_parallel(backward_fn, input, weight, bias, offset, mask, grad_output, *geom, group=None, comm=None):
    """Run `backward_fn` (signature of `_ext.dcn_v2_backward`) on this rank's batch shard and
    all-reduce grad_weight / grad_bias. Returns the shard's grad_input / grad_offset / grad_mask
    and the GLOBAL grad_weight / grad_bias. With `comm` (a GradComm) the all-reduce happens inside the backward's own
    reduction kernel over NVLink peer memory; without it, one flat-bucket collective of torch.distributed follows."""
    x, off, msk, go = shard_batch(input, offset, mask, grad_output)
    if comm is not None:
        return tuple(backward_fn(x.contiguous(), weight, bias, off.contiguous(), msk.contiguous(), go.contiguous(),
                                 *geom, comm=comm))
    g_in, g_off, g_msk, g_w, g_b = backward_fn(x.contiguous(), weight, bias, off.contiguous(),
                                                msk.contiguous(), go.contiguous(), *geom)
    allreduce_weight_grads([g_w, g_b], group=group)
    return g_in, g_off, g_msk, g_w, g_b


def encode_windows_data_parallel(encode_fn, windows):
    """Each rank encodes the event windows it owns (no collective): `windows` is a list of argument
    tuples for `encode_fn`; returns [(global_index, encoded)] for this rank's share."""
    s, e = shard_range(len(windows))
    return [(i, encode_fn(*windows[i])) for i in range(s, e)]
