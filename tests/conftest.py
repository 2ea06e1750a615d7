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
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def assert_parity(got, ref, scale, rtol=1e-12, what=""):
    """north_star tolerance: |got - ref| <= rtol * max(|ref|, scale); NaN pattern must be identical.
    ``scale`` is the field scale S = max|component| (divided by h for gradients), SURVEY 8d."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert np.array_equal(nan_g, nan_r), f"{what}: NaN masks differ at rows {np.where(nan_g != nan_r)[0][:10]}"
    scale = np.broadcast_to(np.asarray(scale, dtype=np.float64), ref.shape)
    ok = ~nan_r
    err = np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), scale[ok])
    worst = float(err.max()) if err.size else 0.0
    assert worst <= rtol, f"{what}: scaled error {worst:.3e} > {rtol:.1e}"
    return worst


@pytest.fixture(scope="session")
def cuda_lib():
    from arbinterp_b200 import _lib
    return _lib.load()
