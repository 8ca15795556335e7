import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    with np.load(os.path.join(GOLDEN, 'reference_numpy.npz')) as d:
        return {k: d[k] for k in d.files}


@pytest.fixture(scope='session')
def schematic_patch():
    from PIL import Image
    return np.asarray(Image.open(os.path.join(GOLDEN, 'schematic_crop_256.png')).convert('RGB'), dtype=np.float32) / 255.0


def rel_err(a, b):
    """max |a-b| / max |b| (scale-relative error, the tolerance unit of BASELINE.json's '1e-5 relative')."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


def assert_parity(got, truth64, ref32=None, tol=1e-5, slack=4.0, what=''):
    """got (CUDA fp32) must match the float64 oracle within `tol` relative, or — for deep chains where float32
    arithmetic itself drifts — within `slack` x the float32 oracle's own distance from the float64 oracle."""
    e = rel_err(got, truth64)
    bound = tol
    if ref32 is not None:
        bound = max(tol, slack * rel_err(ref32, truth64))
    assert e <= bound, '{}: relative error {:.3e} > bound {:.3e}'.format(what, e, bound)
    return e
