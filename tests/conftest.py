import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def rel_err(got, want):
    """max-abs-err / max-abs-ref per tensor — the parity metric of SURVEY.md §8(d)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    if want.size == 0:
        return 0.0
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


DCN_CASES = ["dcn_small_dg1", "dcn_dg4_stride2", "dcn_dil2_bigoff", "dcn_pad_h_ne_w", "dcn_k1",
             "dcn_border", "dcn_c64_dg8"]
FAC_CASES = ["fac_k5", "fac_k3", "fac_k1", "fac_k5_odd"]
EVENT_CASES = ["events_plain", "events_oob", "events_dupts", "events_f64"]

# Parity gates of BASELINE.json's north_star (fp32): forward 1e-5, gradients 1e-4, relative.
FWD_TOL = 1e-5
GRAD_TOL = 1e-4


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    return o
