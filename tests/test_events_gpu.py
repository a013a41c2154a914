"""GPU parity of the event encoders vs the golden vectors produced by the reference's own
dataloader/encodings.py, and vs the CPU oracle at larger sizes."""
import numpy as np
import pytest
import torch

from conftest import EVENT_CASES, load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def enc():
    from ebfi_be_b200 import encodings
    return encodings


def _ev(g):
    from gpu_util import t
    dt = torch.float64 if g["xs"].dtype == np.float64 else torch.float32
    return t(g["xs"], dt), t(g["ys"], dt), t(g["ts"], dt), t(g["ps"])


@pytest.mark.parametrize("case", EVENT_CASES)
def test_golden_vectors(enc, case):
    from gpu_util import n
    g = load_golden(case)
    H, W = (int(v) for v in g["sensor"])
    xs, ys, ts, ps = _ev(g)
    img = enc.events_to_image(xs, ys, ps, sensor_size=(H, W))
    assert np.array_equal(n(img), g["image"])                     # +-1 sums: exact
    # in-place zeroing of out-of-range events, encodings.py:254-256
    assert np.array_equal(n(xs), g["image_xs"]) and np.array_equal(n(ys), g["image_ys"])
    assert np.array_equal(n(ps), g["image_ps"])
    xs, ys, ts, ps = _ev(g)
    assert np.array_equal(n(enc.events_to_mask(xs, ys, ps, sensor_size=(H, W))), g["mask"])
    xs, ys, ts, ps = _ev(g)
    assert np.array_equal(n(enc.events_to_channels(xs, ys, ps, sensor_size=(H, W))), g["channels"])
    if "voxel5" in g:
        xs, ys, ts, ps = _ev(g)
        vox = enc.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(H, W))
        assert rel_err(n(vox), g["voxel5"]) < 1e-6               # fp32 sums in a different order
        assert np.array_equal(n(xs), g["voxel_xs"]) and np.array_equal(n(ys), g["voxel_ys"])
    for nb in (4, 16):
        xs, ys, ts, ps = _ev(g)
        st = enc.events_to_stack(xs, ys, ts, ps, nb, sensor_size=(H, W))
        assert np.array_equal(n(st), g[f"stack{nb}"]), nb          # counts: exact
        assert np.array_equal(n(xs), g[f"stack{nb}_xs"]) and np.array_equal(n(ys), g[f"stack{nb}_ys"])


def test_degenerate_inputs(enc):
    from gpu_util import dev, n
    g = load_golden("events_degenerate")
    z = torch.zeros(5, device=dev())
    assert np.array_equal(n(enc.events_to_stack(z, z.clone(), z.clone(), torch.ones(5, device=dev()), 3, (4, 4))),
                          g["stack_tssum0"])
    a = torch.tensor([1., 2, 3], device=dev())
    st = enc.events_to_stack(a, a.clone(), torch.tensor([0., .5, 1], device=dev()), torch.ones(3, device=dev()), 3, (4, 4))
    assert np.array_equal(n(st), g["stack_len3"])
    e = torch.zeros(0, device=dev())
    assert float(enc.events_to_image(e, e, e, (4, 4)).abs().sum()) == 0
    assert float(enc.events_to_voxel(e, e, e, e, 5, (4, 4)).abs().sum()) == 0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_against_oracle_200k_events(enc, oracle, dtype):
    from gpu_util import n, t
    rng = np.random.default_rng(0)
    N, H, W = 200_000, 72, 128
    xs = rng.integers(-2, W + 2, N).astype(dtype)                 # a few out-of-range on both sides
    ys = rng.integers(-1, H + 1, N).astype(dtype)
    ts = np.sort(rng.random(N))
    ts = ((ts - ts[0]) / (ts[-1] - ts[0] + 1e-6)).astype(dtype)
    ps = (rng.integers(0, 2, N) * 2 - 1).astype(np.float32)
    td = torch.float64 if dtype == np.float64 else torch.float32
    st = enc.events_to_stack(t(xs, td), t(ys, td), t(ts, td), t(ps), 16, sensor_size=(H, W))
    want, *_ = oracle.events_to_stack(xs, ys, ts, ps, 16, (H, W))
    assert np.array_equal(n(st), want)
    if dtype == np.float32:
        vox = enc.events_to_voxel(t(xs, td), t(ys, td), t(ts, td), t(ps), 5, sensor_size=(H, W))
        want, *_ = oracle.events_to_voxel(xs, ys, ts, ps, 5, (H, W))
        assert rel_err(n(vox), want) < 1e-5
        assert abs(float(vox.double().sum()) - float(want.astype(np.float64).sum())) < 1e-2


def test_full_size_properties(enc):
    """BASELINE config 3: 10 M events, 5 bins, 1280x720. In-range events: the voxel grid sums to
    sum(ps) (bilinear weights of the two touched bins add to 1), the stack counts every event
    once per polarity (no timestamp sits on a bin boundary), channels == stack summed over bins."""
    from gpu_util import dev
    torch.manual_seed(0)
    N, H, W = 10_000_000, 720, 1280
    xs = torch.randint(0, W, (N,), device=dev()).float()
    ys = torch.randint(0, H, (N,), device=dev()).float()
    ts = torch.sort(torch.rand(N, device=dev(), dtype=torch.float64))[0]
    ts = ((ts - ts[0]) / (ts[-1] - ts[0] + 1e-6))
    ps = (torch.randint(0, 2, (N,), device=dev()) * 2 - 1).float()
    vox = enc.events_to_voxel(xs, ys, ts.float(), ps, 5, sensor_size=(H, W))
    assert abs(float(vox.double().sum()) - float(ps.double().sum())) < 1.0
    st = enc.events_to_stack(xs.double(), ys.double(), ts, ps, 16, sensor_size=(H, W))
    assert float(st[0].sum()) == float((ps > 0).sum()) and float(st[1].sum()) == float((ps < 0).sum())
    ch = enc.events_to_channels(xs, ys, ps, sensor_size=(H, W))
    assert torch.equal(ch, st.sum(1))
