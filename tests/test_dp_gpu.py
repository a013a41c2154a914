"""GPU: the fused weight-gradient all-reduce over NVLink peer memory (csrc/dp_comm.cuh).
world = 1 runs on any box (the exchange degenerates to a publish + read-back of the rank's own buffer, same kernels);
world = 2 spawns two processes and needs two GPUs (skipped otherwise) — `gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py`."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(dev, seed, B=1, H=40, W=48):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    x, off, msk = r(B, 64, H, W), 2 * r(B, 144, H, W), torch.sigmoid(r(B, 72, H, W))
    w, b, go = (torch.rand(64, 64, 3, 3, generator=g) * 2 - 1) / 24, r(64), r(B, 64, H, W)
    return [v.to(dev) for v in (x, w, b, off, msk, go)]


GEOM = (3, 3, 1, 1, 1, 1, 1, 1, 8)


class _SoloComm:
    """world = 1 communicator over a plain device buffer (no peers to map)."""
    def __init__(self, n, dev):
        from ebfi_be_b200 import _lib as L
        nbytes = int(L.load().ebfi_dp_comm_bytes(n))
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self.struct = L.DpComm(1, 0, (L.c_void * 8)(self.buf.data_ptr()), nbytes)


def test_world_one_backward_dp_equals_backward():
    from ebfi_be_b200 import _lib as L
    from ebfi_be_b200.shims import _ext
    dev = torch.device("cuda:0")
    x, w, b, off, msk, go = _inputs(dev, 3)
    comm = _SoloComm(w.numel() + b.numel(), dev)
    want = _ext.dcn_v2_backward(x, w, b, off, msk, go, *GEOM)
    for _ in range(3):                                  # the device-side epoch counter advances by itself
        got = _ext.dcn_v2_backward_dp(x, w, b, off, msk, go, *GEOM, comm)
        for a, c in zip(got, want):
            assert torch.equal(a, c)
    got = _ext.dcn_v2_backward_dp(x, w, b, off, msk, go, *GEOM, comm, defer=True)      # publish ... complete
    comm_call = lambda fn, *t: L.check(getattr(L.load(), fn)(L.stream_ptr(dev), comm.struct, L.ptr(t[0]), t[0].numel(),
                                                             L.ptr(t[1]), t[1].numel()), fn)
    comm_call("ebfi_dp_complete", got[3], got[4])
    for a, c in zip(got, want):
        assert torch.equal(a, c)
    a, c = torch.randn(1000, device=dev), torch.randn(77, device=dev)
    a0, c0 = a.clone(), c.clone()
    L.check(L.load().ebfi_dp_allreduce_sum(L.stream_ptr(dev), comm.struct, L.ptr(a), a.numel(), L.ptr(c), c.numel()), "dp")
    torch.cuda.synchronize()
    assert torch.equal(a, a0) and torch.equal(c, c0)
    assert int(comm.buf[:4].view(torch.int32)[0]) == 5     # five publishing launches so far


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from ebfi_be_b200 import parallel
        from ebfi_be_b200.shims import _ext
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", device_id=dev)
        x, w, b, off, msk, go = _inputs(dev, 100 + rank)
        w, b = _inputs(dev, 100)[1:3]                    # replicated parameters, per-rank data
        comm = parallel.GradComm(w.numel() + b.numel(), dev)
        ok = True
        for it in range(4):
            if it % 2 == 0:
                fused = _ext.dcn_v2_backward_dp(x, w, b, off, msk, go, *GEOM, comm)
            else:                                        # two halves with unrelated work in between
                fused = _ext.dcn_v2_backward_dp(x, w, b, off, msk, go, *GEOM, comm, defer=True)
                torch.zeros(32 << 20, device=dev).sum()
                comm.complete(fused[3], fused[4])
            local = _ext.dcn_v2_backward(x, w, b, off, msk, go, *GEOM)
            gw, gb = local[3].clone(), local[4].clone()
            dist.all_reduce(gw); dist.all_reduce(gb)
            # two ranks: a + b == b + a, so NCCL's sum and the rank-ordered sum agree bit for bit
            tol = 0.0 if world == 2 else 1e-6 * float(gw.abs().max())
            ok &= float((fused[3] - gw).abs().max()) <= tol and float((fused[4] - gb).abs().max()) <= tol * 64 + 0.0
            ok &= all(torch.equal(a, c) for a, c in zip(fused[:3], local[:3]))
            # identical bits on every rank
            ref = fused[3].clone(); dist.broadcast(ref, 0)
            ok &= torch.equal(ref, fused[3])
        t = torch.arange(5000, device=dev, dtype=torch.float32) * (rank + 1)
        comm.publish(t)
        comm.complete(t)
        torch.cuda.synchronize()
        ok &= torch.equal(t, torch.arange(5000, device=dev, dtype=torch.float32) * sum(range(1, world + 1)))
        t = torch.arange(5000, device=dev, dtype=torch.float32) * (rank + 1)
        comm.allreduce_(t)
        torch.cuda.synchronize()
        ok &= torch.equal(t, torch.arange(5000, device=dev, dtype=torch.float32) * sum(range(1, world + 1)))
        q.put((rank, bool(ok), ""))
        dist.destroy_process_group()
    except Exception as e:                               # surface the failure instead of a join timeout
        import traceback
        q.put((rank, False, traceback.format_exc()[-1500:]))


@pytest.mark.parametrize("world", [2, 8])
def test_fused_allreduce_matches_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, ok, msg in res:
        assert ok, f"rank {rank}: {msg}"
