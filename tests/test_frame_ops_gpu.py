"""Frame2Lap / Frame2DCP on the GPU (csrc/frame_ops.cu) vs OpenCV-generated golden vectors, the CPU oracle and,
where cv2 is importable, OpenCV itself at 720p. Integer / min arithmetic: bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fo():
    from ebfi_be_b200 import frame_ops
    return frame_ops


def test_golden_vectors(fo):
    from gpu_util import n, t
    g = load_golden("frames")
    assert np.array_equal(n(fo.Frame2Lap(t(g["frames"]))), g["lap"])
    assert np.array_equal(n(fo.Frame2Lap(t(g["tiny"]))), g["tiny_lap"])
    for sz, key in ((35, "dcp35"), (4, "dcp4"), (1, "dcp1")):
        assert np.array_equal(n(fo.Frame2DCP(t(g["frames"]), sz)), g[key]), sz
    assert np.array_equal(n(fo.Frame2DCP(t(g["tiny"]), 35)), g["tiny_dcp"])


def test_720p_against_oracle_and_opencv(fo, oracle):
    from gpu_util import dev, n
    torch.manual_seed(0)
    frames = torch.rand(2, 3, 720, 1280, device=dev())
    lap, dark = n(fo.Frame2Lap(frames)), n(fo.Frame2DCP(frames))
    f = frames.cpu().numpy()
    assert np.array_equal(lap, oracle.frame_to_lap(f))
    assert np.array_equal(dark, oracle.frame_to_dcp(f, 35))
    cv2 = pytest.importorskip("cv2")
    im = (f[0].transpose(1, 2, 0) * 255).astype(np.uint8)                # myutils/utils.py:43-46
    assert np.array_equal(lap[0, 0], cv2.Laplacian(cv2.cvtColor(im, cv2.COLOR_BGR2GRAY), cv2.CV_64F).astype("float32"))
    dc = f[0].min(axis=0)
    assert np.array_equal(dark[0, 0], cv2.erode(dc, cv2.getStructuringElement(cv2.MORPH_RECT, (35, 35))))


def test_rejects_wrong_shape(fo):
    from gpu_util import dev
    with pytest.raises(RuntimeError, match="B, 3, H, W"):
        fo.Frame2Lap(torch.rand(1, 1, 8, 8, device=dev()))
