"""Seeded random sweeps over operator geometries (shapes the hand-picked cases do not list): DCNv2 with non-square
kernels / strides / dilations / paddings / group counts, FAC with K in {1,3,5,7} and ragged sizes, event stacks with
random bin counts — each against the CPU oracle at the gates of BASELINE.json's north_star."""
import numpy as np
import pytest
import torch

from conftest import FWD_TOL, GRAD_TOL, rel_err

pytestmark = pytest.mark.gpu
GRADS = ["grad_input", "grad_offset", "grad_mask", "grad_weight", "grad_bias"]


def _dcn_configs(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        dg = int(rng.choice([1, 2, 3, 4, 8]))
        cpg = int(rng.choice([1, 2, 4, 8, 8, 16]))
        C, Co = dg * cpg, int(rng.choice([3, 8, 16, 64, 64, 80]))
        kh, kw = int(rng.choice([1, 2, 3, 3])), int(rng.choice([1, 3, 3, 4]))
        sh, sw = int(rng.choice([1, 1, 2])), int(rng.choice([1, 1, 2]))
        dh, dw = int(rng.choice([1, 1, 2])), int(rng.choice([1, 1, 3]))
        ph, pw = int(rng.integers(0, 3)), int(rng.integers(0, 3))
        B, H, W = int(rng.integers(1, 3)), int(rng.integers(6, 30)), int(rng.integers(6, 30))
        Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
        Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
        if Ho > 0 and Wo > 0:
            out.append((B, C, Co, H, W, kh, kw, sh, sw, ph, pw, dh, dw, dg, float(rng.choice([0.5, 2.0, 6.0]))))
    return out


@pytest.mark.parametrize("cfg", _dcn_configs(24, 2024))
def test_dcn_random_geometry(oracle, cfg):
    from ebfi_be_b200 import dcn_v2
    from gpu_util import n, t
    B, C, Co, H, W, kh, kw, sh, sw, ph, pw, dh, dw, dg, osc = cfg
    rng = np.random.default_rng(abs(hash(cfg)) % (2 ** 32))
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    x = rng.standard_normal((B, C, H, W), dtype=np.float32)
    w = ((rng.random((Co, C, kh, kw), dtype=np.float32) * 2 - 1) / np.sqrt(C * kh * kw)).astype(np.float32)
    b = rng.standard_normal(Co, dtype=np.float32)
    off = (rng.standard_normal((B, 2 * dg * kh * kw, Ho, Wo)) * osc).astype(np.float32)
    msk = rng.random((B, dg * kh * kw, Ho, Wo), dtype=np.float32)
    go = rng.standard_normal((B, Co, Ho, Wo), dtype=np.float32)
    ts = [t(a).requires_grad_() for a in (x, off, msk, w, b)]
    out = dcn_v2.dcn_v2_conv(*ts, (sh, sw), (ph, pw), (dh, dw), dg)
    out.backward(t(go))
    assert rel_err(n(out), oracle.dcn_forward(x, off, msk, w, b, (sh, sw), (ph, pw), (dh, dw), dg)) < FWD_TOL
    want = oracle.dcn_backward(x, off, msk, w, b, go, (sh, sw), (ph, pw), (dh, dw), dg)
    for name, v, ref in zip(GRADS, ts, want):
        assert rel_err(n(v.grad), ref) < GRAD_TOL, name


def _fac_configs(n, seed):
    rng = np.random.default_rng(seed)
    return [(int(rng.integers(1, 3)), int(rng.integers(1, 6)), int(rng.choice([1, 3, 5, 5, 7])),
             int(rng.integers(1, 70)), int(rng.integers(1, 300))) for _ in range(n)]


@pytest.mark.parametrize("cfg", _fac_configs(20, 77))
def test_fac_random_shapes(oracle, cfg):
    from ebfi_be_b200.kernelconv2d import KernelConv2DFunction
    from gpu_util import n, t
    B, C, K, H, W = cfg
    rng = np.random.default_rng(abs(hash(cfg)) % (2 ** 32))
    xi = rng.standard_normal((B, C, H + K - 1, W + K - 1), dtype=np.float32)
    ker = rng.standard_normal((B, C * K * K, H, W), dtype=np.float32)
    go = rng.standard_normal((B, C, H, W), dtype=np.float32)
    xg, kg = t(xi).requires_grad_(), t(ker).requires_grad_()
    out = KernelConv2DFunction.apply(xg, kg, K)
    out.backward(t(go))
    assert rel_err(n(out), oracle.fac_forward(xi, ker, K)) < FWD_TOL
    gi, gk = oracle.fac_backward(xi, ker, go, K)
    assert rel_err(n(xg.grad), gi) < GRAD_TOL and rel_err(n(kg.grad), gk) < GRAD_TOL


@pytest.mark.parametrize("seed", range(8))
def test_event_stack_random_bins_and_sizes(oracle, seed):
    from ebfi_be_b200 import encodings
    from gpu_util import dev, n, t
    rng = np.random.default_rng(100 + seed)
    N, H, W, bins = int(rng.integers(4, 50_000)), int(rng.integers(1, 90)), int(rng.integers(1, 120)), int(rng.integers(1, 40))
    xs = rng.integers(-1, W + 1, N).astype(np.int16)
    ys = rng.integers(-1, H + 1, N).astype(np.int16)
    ts = 3.0 + np.round(np.sort(rng.random(N)) * 2000) / 1e4            # coarse stamps: many duplicates and boundary hits
    ps = (rng.integers(0, 2, N) * 2 - 1).astype(np.int8)
    got = encodings.events_raw_to_stack(*(torch.from_numpy(a).to(dev()) for a in (xs, ys, ts, ps)), bins, (H, W))
    assert np.array_equal(n(got), oracle.dataset_event_stack(xs, ys, ts, ps, bins, (H, W)))
    tn = ((ts - ts[0]) / (ts[-1] - ts[0] + 1e-6)).astype(np.float32)
    st = encodings.events_to_stack(t(xs.astype(np.float32)), t(ys.astype(np.float32)), t(tn), t(ps.astype(np.float32)), bins, (H, W))
    assert np.array_equal(n(st), oracle.events_to_stack(xs.astype(np.float32), ys.astype(np.float32), tn, ps.astype(np.float32), bins, (H, W))[0])


def _shifted(a, dtype=torch.float32):
    """Contiguous CUDA copy of `a` whose storage starts one element off the allocation (not 16-byte aligned)."""
    from gpu_util import dev
    src = torch.as_tensor(a).to(dtype)
    buf = torch.empty(src.numel() + 1, dtype=dtype, device=dev())
    v = buf[1:].view(src.shape)
    v.copy_(src)
    assert v.data_ptr() % 16 != 0
    return v


def test_misaligned_views_fac_kpn_events(oracle):
    """Vector loads / TMA need 16-byte aligned bases; contiguous views that are not aligned must take the scalar
    paths and still match the oracle (FAC forward + backward, fused KernelConv->FAC, raw event stack)."""
    from ebfi_be_b200 import encodings, modification
    from ebfi_be_b200.shims import kernelconv2d_cuda as kc
    from gpu_util import n
    rng = np.random.default_rng(5)
    B, C, K, H, W = 1, 3, 5, 20, 256
    xi = rng.standard_normal((B, C, H + 4, W + 4), dtype=np.float32)
    ker = rng.standard_normal((B, C * 25, H, W), dtype=np.float32)
    go = rng.standard_normal((B, C, H, W), dtype=np.float32)
    out, gi, gk = _shifted(np.zeros_like(go)), _shifted(np.zeros_like(xi)), _shifted(np.zeros_like(ker))
    kc.forward(_shifted(xi), _shifted(ker), K, out)
    kc.backward(_shifted(xi), _shifted(ker), K, _shifted(go), gi, gk)
    assert rel_err(n(out), oracle.fac_forward(xi, ker, K)) < FWD_TOL
    wi, wk = oracle.fac_backward(xi, ker, go, K)
    assert rel_err(n(gi), wi) < GRAD_TOL and rel_err(n(gk), wk) < GRAD_TOL

    ev = rng.standard_normal((1, 32, 19, 12), dtype=np.float32)
    fr = rng.standard_normal((1, 32, 19, 12), dtype=np.float32)
    w = (0.05 * rng.standard_normal((800, 64, 3, 3))).astype(np.float32)
    b = (0.1 * rng.standard_normal(800)).astype(np.float32)
    with torch.no_grad():
        got = modification.kernelconv_fac_fused(_shifted(ev), _shifted(fr), _shifted(w), _shifted(b), 5, 0.01)
    assert rel_err(n(got), oracle.kpn_fused_forward(ev, fr, w, b, 5, 0.01, bf16_operands=True)[0]) < 1e-5

    N = 5000
    xs, ys = rng.integers(0, 40, N).astype(np.int16), rng.integers(0, 30, N).astype(np.int16)
    ts = 1.0 + np.sort(rng.random(N))
    ps = (rng.integers(0, 2, N) * 2 - 1).astype(np.int8)
    st = encodings.events_raw_to_stack(_shifted(xs, torch.int16), _shifted(ys, torch.int16), _shifted(ts, torch.float64),
                                       _shifted(ps, torch.int8), 8, (30, 40))
    assert np.array_equal(n(st), oracle.dataset_event_stack(xs, ys, ts, ps, 8, (30, 40)))
