"""CPU, world_size = 2, gloo: the N > 1 path shards the batch with no data-path collective and
all-reduces only DCNv2's weight / bias gradients. The per-shard compute is the CPU oracle here
(tests may use it); on the GPU box the same helpers wrap `_ext.dcn_v2_backward`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GRAD_TOL, rel_err


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_backward(x, w, b, off, msk, go, kh, kw, sh, sw, ph, pw, dh, dw, dg):
    from oracle import oracle
    out = oracle.dcn_backward(x.numpy(), off.numpy(), msk.numpy(), w.numpy(), b.numpy(), go.numpy(),
                              (sh, sw), (ph, pw), (dh, dw), dg, "f64")
    return [torch.from_numpy(a) for a in out]


def _make(B=4, C=8, Co=6, H=9, W=10, dg=2):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Co, C, 3, 3, generator=g) * 0.1
    b = torch.randn(Co, generator=g)
    off = torch.randn(B, 2 * dg * 9, H, W, generator=g) * 2
    msk = torch.sigmoid(torch.randn(B, dg * 9, H, W, generator=g))
    go = torch.randn(B, Co, H, W, generator=g)
    return x, w, b, off, msk, go, (3, 3, 1, 1, 1, 1, 1, 1, dg)


def _worker(rank, world_size, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from ebfi_be_b200 import parallel
    x, w, b, off, msk, go, geom = _make()
    res = parallel.dcn_backward_data_parallel(_oracle_backward, x, w, b, off, msk, go, *geom)
    s, e = parallel.shard_range(x.shape[0])
    enc = parallel.encode_windows_data_parallel(lambda a, k: a * k, [(np.arange(3), i) for i in range(5)])
    # asynchronous bucket: other work may be issued between the call and wait(); averaged variant
    a, c = torch.full((3, 2), float(rank + 1)), torch.full((5,), 10.0 * (rank + 1))
    h = parallel.allreduce_weight_grads([a, None, c], average=True, async_op=True)
    busy = torch.ones(4).sum()
    h.wait()
    assert float(busy) == 4 and torch.equal(a, torch.full((3, 2), 1.5)) and torch.equal(c, torch.full((5,), 15.0))
    q.put((rank, s, e, [t.numpy() for t in res], [(i, v.tolist()) for i, v in enc]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_every_item_once():
    from ebfi_be_b200 import parallel
    for n in (0, 1, 4, 7, 8, 9):
        for ws in (1, 2, 3, 8):
            owned = [i for r in range(ws) for i in range(*parallel.shard_range(n, r, ws))]
            assert owned == list(range(n))


@pytest.mark.timeout(180)
def test_dp2_gloo_weight_grad_allreduce_matches_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted([q.get(timeout=150) for _ in procs], key=lambda t: t[0])
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    x, w, b, off, msk, go, geom = _make()
    full = [t.numpy() for t in _oracle_backward(x, w, b, off, msk, go, *geom)]
    # sharded data gradients concatenate to the full-batch ones; no exchange was needed
    for i in range(3):
        cat = np.concatenate([g[3][i] for g in got], 0)
        assert rel_err(cat, full[i]) < 1e-12
    # weight / bias gradients: every rank holds the all-reduced (global) value
    for g in got:
        assert rel_err(g[3][3], full[3]) < GRAD_TOL and rel_err(g[3][4], full[4]) < GRAD_TOL
    assert np.array_equal(got[0][3][3], got[1][3][3])
    assert [(g[1], g[2]) for g in got] == [(0, 2), (2, 4)]
    # event windows: 5 windows over 2 ranks -> 3 + 2, each encoded exactly once, no collective
    assert sorted(i for g in got for i, _ in g[4]) == [0, 1, 2, 3, 4]
