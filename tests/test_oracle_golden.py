"""CPU: pin the oracle (oracle/liboracle.so) against the reference's golden vectors.

The vectors in tests/golden/ were produced by the reference itself (see
tests/golden/make_golden.py): its CPU DCNv2 build for the five gradients,
torchvision for the DCN forward (the reference CPU forward returns uninitialised
memory), dataloader/encodings.py for the encoders, and an F.unfold/autograd
restatement for FAC (no reference CPU path exists).
"""
import numpy as np
import pytest

from conftest import (DCN_CASES, EVENT_CASES, FAC_CASES, FWD_TOL, GRAD_TOL, load_golden, rel_err)


def _geom(g):
    kh, kw, sh, sw, ph, pw, dh, dw, dg = (int(v) for v in g["geom"])
    return (sh, sw), (ph, pw), (dh, dw), dg


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("case", DCN_CASES + ["dcn_zero_offset"])
def test_dcn_forward_matches_reference(oracle, case, precision):
    g = load_golden(case)
    s, p, d, dg = _geom(g)
    out = oracle.dcn_forward(g["input"], g["offset"], g["mask"], g["weight"], g["bias"], s, p, d, dg,
                             precision)
    assert rel_err(out, g["output"]) < FWD_TOL


def test_dcn_zero_offset_known_answer(oracle):
    # testcuda.py:32-67 — identity weights, zero offsets, mask 0.5  =>  2*out == input
    g = load_golden("dcn_zero_offset")
    s, p, d, dg = _geom(g)
    out = oracle.dcn_forward(g["input"], g["offset"], g["mask"], g["weight"], g["bias"], s, p, d, dg, "f32")
    assert np.abs(2 * out - g["input"]).max() < 1e-10


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("case", DCN_CASES)
def test_dcn_backward_matches_reference_cpu_build(oracle, case, precision):
    g = load_golden(case)
    s, p, d, dg = _geom(g)
    grads = oracle.dcn_backward(g["input"], g["offset"], g["mask"], g["weight"], g["bias"],
                                g["grad_output"], s, p, d, dg, precision)
    for name, got in zip(["grad_input", "grad_offset", "grad_mask", "grad_weight", "grad_bias"], grads):
        assert rel_err(got, g[name]) < GRAD_TOL, name


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("case", FAC_CASES)
def test_fac_matches_unfold_restatement(oracle, case, precision):
    g = load_golden(case)
    K = int(g["K"])
    out = oracle.fac_forward(g["input"], g["kernel"], K, precision)
    assert rel_err(out, g["output"]) < FWD_TOL
    gi, gk = oracle.fac_backward(g["input"], g["kernel"], g["grad_output"], K, precision)
    assert rel_err(gi, g["grad_input"]) < GRAD_TOL
    assert rel_err(gk, g["grad_kernel"]) < GRAD_TOL


@pytest.mark.parametrize("case", EVENT_CASES)
def test_event_encoders_bit_exact(oracle, case):
    g = load_golden(case)
    H, W = (int(v) for v in g["sensor"])
    img, xs, ys, ps = oracle.events_to_image(g["xs"], g["ys"], g["ps"], (H, W))
    assert np.array_equal(img, g["image"])
    # the reference zeroes out-of-range events in its arguments (encodings.py:254-256)
    assert np.array_equal(xs, g["image_xs"]) and np.array_equal(ys, g["image_ys"])
    assert np.array_equal(ps, g["image_ps"])
    msk, *_ = oracle.events_to_image(g["xs"], g["ys"], g["ps"], (H, W), binary=True)
    assert np.array_equal(msk, g["mask"])
    if "voxel5" in g:
        vox, xs, ys = oracle.events_to_voxel(g["xs"], g["ys"], g["ts"], g["ps"], 5, (H, W))
        assert np.array_equal(vox, g["voxel5"])
        assert np.array_equal(xs, g["voxel_xs"]) and np.array_equal(ys, g["voxel_ys"])
    for nb in (4, 16):
        st, xs, ys, _ = oracle.events_to_stack(g["xs"], g["ys"], g["ts"], g["ps"], nb, (H, W))
        assert np.array_equal(st, g[f"stack{nb}"]), nb
        assert np.array_equal(xs, g[f"stack{nb}_xs"]) and np.array_equal(ys, g[f"stack{nb}_ys"])


def test_event_stack_degenerate(oracle):
    g = load_golden("events_degenerate")
    z = np.zeros(5, np.float32)
    st, *_ = oracle.events_to_stack(z, z, z, np.ones(5, np.float32), 3, (4, 4))
    assert np.array_equal(st, g["stack_tssum0"])
    a = np.array([1, 2, 3], np.float32)
    st, *_ = oracle.events_to_stack(a, a, np.array([0, .5, 1], np.float32), np.ones(3, np.float32), 3, (4, 4))
    assert np.array_equal(st, g["stack_len3"])


RAW_CASES = ["plain", "dupts_oob", "ps01", "len3", "len4", "empty", "same_ts"]


@pytest.mark.parametrize("name", RAW_CASES)
def test_dataset_event_path_on_disk_dtypes(oracle, name):
    """GetEvents (h5dataset.py:327-349) restated around the oracle's events_to_stack vs the reference
    encoder run on the same int16 / int16 / float64 / int8 slices (tests/golden/make_golden.py)."""
    g = load_golden("events_raw")
    H, W = (int(v) for v in g["sensor"])
    for nb in (4, 16):
        got = oracle.dataset_event_stack(g[f"{name}_xs"], g[f"{name}_ys"], g[f"{name}_ts"], g[f"{name}_ps"], nb, (H, W))
        assert got.shape == (nb, 2, H, W)
        assert np.array_equal(got, g[f"{name}_stack{nb}"]), (name, nb)


def test_kpn_oracle_matches_torch_op_sequence(oracle):
    """oracle.kpn_fused_forward vs the reference's modules applied literally with torch (nn.Conv2d -> nn.LeakyReLU
    -> ReplicationPad2d -> unfold contraction), model_singleframe.py:145-146,159-162 + KernelConv2D.py:82-87."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(0)
    B, Ce, Cf, H, W, K = 2, 8, 8, 9, 7, 5
    ev = rng.standard_normal((B, Ce, H, W), dtype=np.float32)
    fr = rng.standard_normal((B, Cf, H, W), dtype=np.float32)
    w = (0.1 * rng.standard_normal((Ce * K * K, Ce + Cf, 3, 3))).astype(np.float32)
    b = rng.standard_normal(Ce * K * K).astype(np.float32)
    out, ker = oracle.kpn_fused_forward(ev, fr, w, b, K)
    feat = torch.cat([torch.from_numpy(ev), torch.from_numpy(fr)], 1).double()
    kt = torch.nn.LeakyReLU()(F.conv2d(feat, torch.from_numpy(w).double(), torch.from_numpy(b).double(), padding=1))
    evp = torch.nn.ReplicationPad2d(2)(torch.from_numpy(ev).double())
    want = (F.unfold(evp, K).view(B, Ce, K * K, H, W) * kt.view(B, Ce, K * K, H, W)).sum(2)
    assert np.abs(ker - kt.numpy()).max() < 1e-12 and np.abs(out - want.numpy()).max() < 1e-12
    # the bf16 rounding helper is torch's round-to-nearest-even
    x = rng.standard_normal(1000).astype(np.float32) * 3
    assert np.array_equal(oracle._bf16_round(x), torch.from_numpy(x).bfloat16().float().numpy())


def test_frame_maps_match_opencv_golden(oracle):
    """oracle.frame_to_lap / frame_to_dcp vs Frame2Lap / Frame2DCP (myutils/utils.py:15-49) run with OpenCV
    (tests/golden/make_golden.py::frame_cases) — integer / min arithmetic: bit-exact."""
    g = load_golden("frames")
    assert np.array_equal(oracle.frame_to_lap(g["frames"]), g["lap"])
    assert np.array_equal(oracle.frame_to_lap(g["tiny"]), g["tiny_lap"])
    for sz, key in ((35, "dcp35"), (4, "dcp4"), (1, "dcp1")):
        assert np.array_equal(oracle.frame_to_dcp(g["frames"], sz), g[key]), sz
    assert np.array_equal(oracle.frame_to_dcp(g["tiny"], 35), g["tiny_dcp"])
