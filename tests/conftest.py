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


def pytest_sessionfinish(session, exitstatus):
    """Achieved parity errors against the executed-reference fixtures (tests/golden/tfgraph_common.REPORT, filled by Golden.check):
    written where a GPU run can bring them back (gpurun_out/) — the committed copy lives in profiles/parity_report.json."""
    import json
    try:
        from golden import tfgraph_common as C
    except Exception:
        return
    if not C.REPORT:
        return
    try:
        import torch
        if not torch.cuda.is_available():
            return
    except Exception:
        return
    rows = {}
    for key, err in sorted(C.REPORT.items()):
        case = key.split('/')[0]
        r = rows.setdefault(case, {'tensors': 0, 'max_err': 0.0, 'worst_tensor': ''})
        r['tensors'] += 1
        if err >= r['max_err']:
            r['max_err'], r['worst_tensor'] = err, key[len(case) + 1:]
    out = {'what': 'scale-relative error max|got - ref| / max|ref| of the CUDA path against tests/golden/tf_graph_golden.npz (the reference\'s own '
                   'code executed in float64); for piecewise-continuous tensors the (1 - allowed outlier fraction) quantile',
           'cases': rows}
    for d in (os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')):
        try:
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, 'parity_report.json'), 'w') as f:
                json.dump(out, f, indent=1, sort_keys=True)
        except OSError:
            pass
