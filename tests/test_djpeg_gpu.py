"""GPU parity of the fused dJPEG kernels (through the C-ABI) against the CPU oracle (models/jpeg.py:91-159 restated)."""
import numpy as np
import pytest
import torch

from conftest import assert_parity, rel_err
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def _tables(q):
    return R.jpeg_qtable(q, 0), R.jpeg_qtable(q, 1)


def _oracle(x, q, mode, dtype):
    ql, qc = _tables(q)
    y, X = R.djpeg(torch.tensor(x, dtype=dtype), ql, qc, mode)
    return y.numpy(), X.numpy()


@pytest.mark.parametrize('mode', ['sin', 'harmonic'])
@pytest.mark.parametrize('shape,q', [((1, 8, 8, 3), 50), ((3, 16, 24, 3), 80), ((2, 128, 128, 3), 50), ((1, 256, 256, 3), 90),
                                     ((5, 40, 8, 3), 10), ((2, 8, 72, 3), 100),
                                     # rectangular 32-block tiles of the TMA path: 8 x 4 blocks with a partial last tile across images,
                                     # 4 x 8 blocks with three tiles per block row and a single valid block row, 16 x 2, 32 x 1
                                     ((3, 24, 64, 3), 60), ((1, 8, 96, 3), 40), ((7, 16, 128, 3), 70), ((3, 8, 256, 3), 35),
                                     ((2, 64, 32, 3), 25)])
def test_forward_continuous_modes(mode, shape, q):
    """Continuous quantisers: strict 1e-5 (scale-relative) against the float64 oracle, y and coefficients."""
    from neural_imaging_b200 import ops
    from neural_imaging_b200.tensor import as_device
    x = np.random.RandomState(7).uniform(size=shape).astype(np.float32)
    y, X = ops.djpeg_fwd(as_device(x), *_tables(q), mode, want_coeffs=True)
    y64, X64 = _oracle(x, q, mode, torch.float64)
    y32, X32 = _oracle(x, q, mode, torch.float32)
    assert_parity(y.cpu().numpy(), y64, y32, tol=1e-5, what='y')
    assert_parity(X.cpu().numpy(), X64, X32, tol=1e-5, what='X (block order)')


@pytest.mark.parametrize('shape,q', [((2, 16, 16, 3), 50), ((4, 128, 128, 3), 50), ((1, 256, 256, 3), 80), ((2, 64, 32, 3), 25)])
def test_forward_soft_tie_aware(shape, q):
    """'soft' = hard rounding forward: identical integers except where X/Q sits within delta of k+0.5."""
    from neural_imaging_b200 import ops
    from neural_imaging_b200.tensor import as_device
    x = np.random.RandomState(11).uniform(size=shape).astype(np.float32)
    ql, qc = _tables(q)
    y, X = ops.djpeg_fwd(as_device(x), ql, qc, 'soft', want_coeffs=True)
    y64, X64 = _oracle(x, q, 'soft', torch.float64)
    # pre-rounding values from the continuous 'identity' path of the oracle
    Zpre = R.djpeg(torch.tensor(x, dtype=torch.float64), ql, qc, 'identity')[1].numpy()
    n, h, w, _ = shape
    nb = (h // 8) * (w // 8)
    Q = np.concatenate([np.tile(ql[None], (nb, 1, 1)), np.tile(qc[None], (2 * nb, 1, 1))], 0)
    Q = np.tile(Q, (n, 1, 1))
    frac = np.abs(np.abs(Zpre / Q - np.round(Zpre / Q)) - 0.5)
    near_tie = frac < 1e-4
    Xg = X.cpu().numpy()
    mism = (np.abs(Xg - X64) > 1e-3 * np.maximum(1, np.abs(X64)))
    assert not np.any(mism & ~near_tie), 'rounded coefficients differ away from ties: {}'.format(int(np.sum(mism & ~near_tie)))
    assert np.mean(mism) < 1e-4
    if not np.any(mism):
        assert_parity(y.cpu().numpy(), y64, tol=1e-5, what='parity')
    else:
        blocks_bad = np.any(mism, axis=(1, 2))
        assert np.mean(blocks_bad) < 1e-3


def test_golden_image(schematic_patch):
    from neural_imaging_b200.models import jpeg
    x = schematic_patch[None]
    for q in (50, 80, 90):
        y = jpeg.JPEG(q, 'sin').process(x).numpy()
        y64 = _oracle(x, q, 'sin', torch.float64)[0]
        assert_parity(y, y64, tol=1e-5, what='parity')
    y = jpeg.JPEG(50, 'soft').process(x).numpy()[0]
    psnr = 10 * np.log10(1.0 / np.mean((y.astype(np.float64) - schematic_patch) ** 2))
    assert abs(psnr - 36.12) < 0.1          # SURVEY.md section 6 anchor


@pytest.mark.parametrize('mode', ['soft', 'sin', 'harmonic'])
@pytest.mark.parametrize('shape,q', [((2, 16, 24, 3), 50), ((3, 64, 64, 3), 80), ((1, 128, 128, 3), 30)])
def test_backward(mode, shape, q):
    from neural_imaging_b200 import ops
    from neural_imaging_b200.tensor import as_device
    rs = np.random.RandomState(5)
    x = rs.uniform(-0.05, 1.05, size=shape).astype(np.float32)       # exercise the clip mask
    dy = rs.normal(size=shape).astype(np.float32)
    ql, qc = _tables(q)
    dx = ops.djpeg_bwd(as_device(x), as_device(dy), ql, qc, mode).cpu().numpy()
    grads = {}
    for dt in (torch.float64, torch.float32):
        xt = torch.tensor(x, dtype=dt, requires_grad=True)
        y = R.djpeg(xt, ql, qc, mode)[0]
        grads[dt], = torch.autograd.grad(y, xt, torch.tensor(dy, dtype=dt))
    g64, g32 = grads[torch.float64].numpy(), grads[torch.float32].numpy()
    if mode == 'soft':
        # hard rounding in the recomputed forward: compare only blocks whose clip mask / rounding agree
        diff = np.abs(dx - g64).reshape(shape[0], shape[1] // 8, 8, shape[2] // 8, 8, 3).max(axis=(2, 4, 5))
        bad = diff > 1e-4 * np.abs(g64).max()
        assert np.mean(bad) < 5e-3, 'blocks with gradient mismatch: {}'.format(np.mean(bad))
    else:
        assert_parity(dx, g64, g32, tol=2e-5, what='dx')


def test_api_wrapper_and_errors():
    from neural_imaging_b200 import _lib
    from neural_imaging_b200.models import jpeg
    x = np.random.RandomState(3).uniform(size=(2, 32, 32, 3)).astype(np.float32)
    model = jpeg.DifferentiableJPEG(50, 'sin')
    y, X = model(x)
    assert tuple(y.shape) == (2, 32, 32, 3) and tuple(X.shape) == (3 * 2 * 16, 8, 8)
    codec = jpeg.JPEG(50, 'soft')
    a, ent = codec.process(x, return_entropy=True)
    assert np.isnan(ent)
    b = codec.process(x, quality=80).numpy()                       # temporary table swap, then restored
    assert not np.allclose(a.numpy(), b) and np.array_equal(codec._model._q_mtx_luma, R.jpeg_qtable(50, 0))
    c = jpeg.differentiable_jpeg(x, 80).numpy()
    assert np.array_equal(b, c)
    assert codec.process(x[0]).shape[0] == 1                        # 3-D input gets a batch dimension
    with pytest.raises((ValueError, _lib.NIError)):
        codec.process(np.zeros((1, 12, 16, 3), np.float32))
    assert codec.process(np.zeros((0, 16, 16, 3), np.float32)).shape[0] == 0


def test_full_size_properties():
    """BASELINE config size (1280 x 128x128x3): size-independent properties instead of the (slow) oracle."""
    from neural_imaging_b200 import ops
    from neural_imaging_b200.tensor import as_device
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.rand((1280, 128, 128, 3), device='cuda', generator=g)
    ql, qc = _tables(50)
    y = ops.djpeg_fwd(x, ql, qc, 'soft')
    assert float(y.min()) >= 0 and float(y.max()) <= 1
    # batch independence: a sub-batch processed alone is bit-identical
    y_sub = ops.djpeg_fwd(x[100:108].contiguous(), ql, qc, 'soft')
    assert torch.equal(y_sub, y[100:108])
    # block equivariance: shifting the image by one 8x8 block shifts the output
    xs = torch.roll(x[:16], shifts=(8, 16), dims=(1, 2)).contiguous()
    assert torch.equal(ops.djpeg_fwd(xs, ql, qc, 'soft'), torch.roll(y[:16], shifts=(8, 16), dims=(1, 2)))
    # oracle on a bounded sample of the big batch
    y64 = _oracle(x[:2].cpu().numpy(), 50, 'sin', torch.float64)[0]
    assert_parity(ops.djpeg_fwd(x[:2].contiguous(), ql, qc, 'sin').cpu().numpy(), y64, tol=1e-5, what='parity')
    # backward is linear in dy
    d1, d2 = torch.randn_like(x[:8]), torch.randn_like(x[:8])
    xb = x[:8].contiguous()
    g1, g2 = ops.djpeg_bwd(xb, d1, ql, qc, 'soft'), ops.djpeg_bwd(xb, d2, ql, qc, 'soft')
    g12 = ops.djpeg_bwd(xb, d1 + 2 * d2, ql, qc, 'soft')
    assert float((g12 - (g1 + 2 * g2)).abs().max()) < 1e-4 * float(g12.abs().max())
