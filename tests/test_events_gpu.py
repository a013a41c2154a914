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


# ---- the datasets' event path on the on-disk dtypes (h5dataset.py:327-349) ----
RAW_CASES = ["plain", "dupts_oob", "ps01", "len3", "len4", "empty", "same_ts"]


def _raw(g, name):
    from gpu_util import dev
    return [torch.from_numpy(np.ascontiguousarray(g[f"{name}_{k}"])).to(dev()) for k in ("xs", "ys", "ts", "ps")]


@pytest.mark.parametrize("name", RAW_CASES)
def test_raw_slices_golden(enc, name):
    from gpu_util import n
    g = load_golden("events_raw")
    H, W = (int(v) for v in g["sensor"])
    for nb in (4, 16):
        got = enc.events_raw_to_stack(*_raw(g, name), nb, (H, W))
        assert np.array_equal(n(got), g[f"{name}_stack{nb}"]), (name, nb)
        ref_layout = enc.events_raw_to_stack(*_raw(g, name), nb, (H, W), bins_major=False)
        assert np.array_equal(n(ref_layout), g[f"{name}_stack{nb}"].transpose(1, 0, 2, 3))


def _raw_cloud(rng, N, H, W, oob=True):
    xs = rng.integers(-2 if oob else 0, W + (2 if oob else 0), N).astype(np.int16)
    ys = rng.integers(-1 if oob else 0, H + (1 if oob else 0), N).astype(np.int16)
    ts = 17.0 + np.round(np.sort(rng.random(N)) * 4e5) / 1e6        # microsecond stamps over 0.4 s: duplicates
    ps = (rng.integers(0, 2, N) * 2 - 1).astype(np.int8)
    return xs, ys, ts.astype(np.float64), ps


def test_raw_against_oracle_300k_events(enc, oracle):
    from gpu_util import dev, n
    rng = np.random.default_rng(11)
    H, W = 72, 128
    arrs = _raw_cloud(rng, 300_000, H, W)
    got = enc.events_raw_to_stack(*(torch.from_numpy(a).to(dev()) for a in arrs), 16, (H, W))
    assert np.array_equal(n(got), oracle.dataset_event_stack(*arrs, 16, (H, W)))


def test_raw_equals_float64_path(enc):
    """Same result as doing what the dataset does (float64 normalisation, concatenate, ps.float()) and
    calling events_to_stack on the converted tensors."""
    from gpu_util import dev, n
    rng = np.random.default_rng(12)
    H, W = 180, 240
    xs, ys, ts, ps = _raw_cloud(rng, 500_000, H, W)
    tn = (ts - ts[0]) / (ts[-1] - ts[0] + 1e-6)
    d = dev()
    want = enc.events_to_stack(torch.from_numpy(xs.astype(np.float64)).to(d), torch.from_numpy(ys.astype(np.float64)).to(d),
                               torch.from_numpy(tn).to(d), torch.from_numpy(ps.astype(np.float32)).to(d), 16, (H, W))
    got = enc.events_raw_to_stack(*(torch.from_numpy(a).to(d) for a in (xs, ys, ts, ps)), 16, (H, W))
    assert np.array_equal(n(got), n(want.transpose(0, 1)))


def test_event_slice_feeder_batches(enc, oracle):
    from gpu_util import dev, n
    rng = np.random.default_rng(13)
    H, W, B = 40, 56, 16
    feeder = enc.EventSliceFeeder(dev(), B, (H, W), max_events=1000)          # forces a staging regrow
    for rep in range(3):                                                      # staging sets are reused
        slices = [_raw_cloud(rng, int(k), H, W) for k in (5000, 0, 3, 12345)]
        out = feeder.encode(slices)
        assert out.shape == (4, B, 2, H, W)
        for i, s in enumerate(slices):
            assert np.array_equal(n(out[i]), oracle.dataset_event_stack(*s, B, (H, W))), (rep, i)


def test_raw_full_size_counts(enc):
    """10 M events on 1280x720 (BASELINE config 3 size): every in-range event is counted once per containing
    bin; with distinct bin interiors the grand total is N + (events shared by two bins)."""
    from gpu_util import dev
    torch.manual_seed(0)
    N, H, W, B = 10_000_000, 720, 1280, 16
    d = dev()
    xs = torch.randint(0, W, (N,), device=d, dtype=torch.int16)
    ys = torch.randint(0, H, (N,), device=d, dtype=torch.int16)
    ts = torch.sort(torch.rand(N, device=d, dtype=torch.float64))[0] + 5.0
    ps = (torch.randint(0, 2, (N,), device=d, dtype=torch.int8) * 2 - 1)
    st = enc.events_raw_to_stack(xs, ys, ts, ps, B, (H, W))
    total = float(st.double().sum())
    assert N <= total <= N + B                       # random float64 stamps: at most one shared event per boundary
    assert float(st[:, 0].double().sum()) + float(st[:, 1].double().sum()) == total
    pos = float((ps > 0).sum())
    assert abs(float(st[:, 0].double().sum()) - pos) <= B


def test_raw_rejects_wrong_dtypes(enc):
    from gpu_util import dev
    z = torch.zeros(8, device=dev())
    with pytest.raises(RuntimeError, match="int16"):
        enc.events_raw_to_stack(z, z, z.double(), z.to(torch.int8), 4, (4, 4))


def test_clustered_events_exact_under_contention(enc, oracle):
    """90 % of the events on Gaussian blobs covering ~1 % of the pixels (heavy atomic contention):
    the polarity counts stay exact, the voxel grid within fp32 summation-order noise."""
    from gpu_util import n, t
    rng = np.random.default_rng(21)
    N, H, W = 400_000, 180, 240
    nb = int(0.9 * N)
    c = rng.integers(0, 3, nb)
    xs = np.concatenate([np.clip(np.array([40, 120, 200])[c] + 3 * rng.standard_normal(nb), 0, W - 1), rng.integers(0, W, N - nb)])
    ys = np.concatenate([np.clip(np.array([50, 90, 140])[c] + 3 * rng.standard_normal(nb), 0, H - 1), rng.integers(0, H, N - nb)])
    perm = rng.permutation(N)
    xs, ys = np.floor(xs[perm]).astype(np.float32), np.floor(ys[perm]).astype(np.float32)
    ts = np.sort(rng.random(N)).astype(np.float32)
    ps = (rng.integers(0, 2, N) * 2 - 1).astype(np.float32)
    st = enc.events_to_stack(t(xs), t(ys), t(ts), t(ps), 16, sensor_size=(H, W))
    assert np.array_equal(n(st), oracle.events_to_stack(xs, ys, ts, ps, 16, (H, W))[0])
    assert float(st.max()) > 50                                   # hot pixels: thousands of events over 16 bins x 2 polarities
    vox = enc.events_to_voxel(t(xs), t(ys), t(ts), t(ps), 5, sensor_size=(H, W))
    want = oracle.events_to_voxel(xs, ys, ts, ps, 5, (H, W))[0]
    assert np.abs(n(vox) - want).max() < 1e-3 * max(1.0, np.abs(want).max())
