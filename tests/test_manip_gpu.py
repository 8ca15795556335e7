"""GPU parity of the manipulation / down-sampling kernels against the CPU oracle (helpers/tf_helpers.py:68-184)."""
import numpy as np
import pytest
import torch

from conftest import assert_parity, rel_err
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def _x(shape, seed=0, lo=0.0, hi=1.0):
    return np.random.RandomState(seed).uniform(lo, hi, size=shape).astype(np.float32)


def _both(fn, x):
    return fn(torch.tensor(x, dtype=torch.float64)).numpy(), fn(torch.tensor(x)).numpy()


@pytest.mark.parametrize('shape', [(2, 16, 16, 3), (1, 33, 20, 3), (3, 64, 64, 3)])
@pytest.mark.parametrize('strength', [1, 0.25, 1.5])
def test_sharpen(shape, strength):
    from neural_imaging_b200.helpers import tf_helpers
    if shape[1] != shape[2]:
        shape = (shape[0], shape[1], shape[1], 3)
    x = _x(shape, 1)
    y = tf_helpers.manipulation_sharpen(x, strength).numpy()
    y64, y32 = _both(lambda t: R.manipulation_sharpen(t, strength), x)
    # hue is discontinuous (wrap-around and max-channel switches): compare where the float32 oracle itself is stable
    stable = np.abs(y32 - y64) < 1e-4
    assert np.mean(stable) > 0.99
    assert np.max(np.abs(y - y64)[stable]) < 2e-4
    assert np.mean(np.abs(y - y64) < 1e-4) > 0.99


@pytest.mark.parametrize('shape,factor', [((2, 16, 16, 3), 50), ((1, 64, 64, 3), 50), ((2, 40, 40, 3), 75), ((1, 30, 30, 3), 40),
                                          ((1, 32, 32, 3), 0.5)])
def test_resample(shape, factor):
    from neural_imaging_b200.helpers import tf_helpers
    x = _x(shape, 2)
    y = tf_helpers.manipulation_resample(x, factor).numpy()
    y64, y32 = _both(lambda t: R.manipulation_resample(t, factor), x)
    assert_parity(y, y64, y32, tol=1e-5, what='resample')


@pytest.mark.parametrize('shape,k,std', [((2, 16, 16, 3), 5, 0.83), ((1, 21, 37, 3), 5, 2.0), ((2, 8, 8, 3), 3, 0.5), ((1, 32, 32, 3), 7, 1.5)])
def test_gaussian_fwd_bwd(shape, k, std):
    from neural_imaging_b200 import ops
    from neural_imaging_b200.tensor import as_device, zeros
    x = _x(shape, 3, -0.2, 1.2)
    dy = np.random.RandomState(4).normal(size=shape).astype(np.float32)
    op = ops.GaussianOp(k)
    xd = as_device(x)
    y = op.forward(xd, torch.empty_like(xd), std, training=True).cpu().numpy()
    dx = op.backward(xd, as_device(dy), zeros(shape), std).cpu().numpy()
    res = {}
    for dt in (torch.float64, torch.float32):
        xt = torch.tensor(x, dtype=dt, requires_grad=True)
        yt = R.manipulation_gaussian(xt, k, std)
        g, = torch.autograd.grad(yt, xt, torch.tensor(dy, dtype=dt))
        res[dt] = (yt.detach().numpy(), g.numpy())
    assert_parity(y, res[torch.float64][0], res[torch.float32][0], tol=1e-5, what='gaussian y')
    # clip mask may flip where the un-clipped value is within rounding of 0 or 1
    bad = np.abs(dx - res[torch.float64][1]) > 1e-5 * np.abs(res[torch.float64][1]).max()
    assert np.mean(bad) < 1e-3


def test_awgn_gamma_median():
    from neural_imaging_b200 import ops
    from neural_imaging_b200.helpers import tf_helpers
    from neural_imaging_b200.tensor import as_device, zeros
    shape = (2, 16, 16, 3)
    x = _x(shape, 5, 0.02, 0.98)
    noise = np.random.RandomState(6).normal(size=shape).astype(np.float32)
    # awgn with injected noise; rounding to 1/255 levels: compare away from ties
    y = tf_helpers.manipulation_awgn(x, 5.1 / 255, noise=noise).numpy()
    y64 = R.manipulation_awgn(torch.tensor(x, dtype=torch.float64), 5.1 / 255, torch.tensor(noise, dtype=torch.float64)).numpy()
    assert np.mean(np.abs(y - y64) > 1e-6) < 1e-3
    # on-device Philox noise: statistics only
    big = _x((4, 64, 64, 3), 7, 0.3, 0.7)
    yn = tf_helpers.manipulation_awgn(big, 5.0 / 255).numpy()
    resid = (yn - big) * 255 / 5.0
    assert abs(resid.mean()) < 0.02 and abs(resid.std() - 1.0) < 0.05
    # gamma
    for s in (2.0, 3.0, 0.7):
        yg = tf_helpers.manipulation_gamma(x, s).numpy()
        g64 = R.manipulation_gamma(torch.tensor(x, dtype=torch.float64), s).numpy()
        assert np.mean(np.abs(yg - g64) > 1e-5) < 2e-3
    # median: exact selection
    for k in (3, 5, 4):
        ym = tf_helpers.manipulation_median(x, k).numpy()
        m64 = R.manipulation_median(torch.tensor(x), k).numpy()
        assert np.array_equal(ym, m64)
    # backward of gamma / awgn / median against autograd
    dy = np.random.RandomState(8).normal(size=shape).astype(np.float32)
    xd, dyd = as_device(x), as_device(dy)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    g_ref, = torch.autograd.grad(R.manipulation_gamma(xt, 2.0), xt, torch.tensor(dy, dtype=torch.float64))
    op = ops.GammaOp()
    g = op.backward(xd, dyd, zeros(shape), 2.0).cpu().numpy()
    bad = np.abs(g - g_ref.numpy()) > 1e-4 * np.abs(g_ref.numpy()).max()
    assert np.mean(bad) < 5e-3
    op = ops.AwgnOp(); op.noise = as_device(noise)
    op.forward(xd, torch.empty_like(xd), 5.1)
    g = op.backward(xd, dyd, zeros(shape), 5.1).cpu().numpy()
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    g_ref, = torch.autograd.grad(R.manipulation_awgn(xt, 5.1 / 255, torch.tensor(noise, dtype=torch.float64)), xt, torch.tensor(dy, dtype=torch.float64))
    bad = np.abs(g - g_ref.numpy()) > 1e-4 * np.abs(g_ref.numpy()).max()
    assert np.mean(bad) < 5e-3
    op = ops.MedianOp()
    g = op.backward(xd, dyd, zeros(shape), 3).cpu().numpy()
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    g_ref, = torch.autograd.grad(R.manipulation_median(xt, 3), xt, torch.tensor(dy, dtype=torch.float64))
    assert_parity(g, g_ref.numpy(), tol=1e-6, what='parity')


@pytest.mark.parametrize('shape,factor', [((2, 16, 16, 3), 50), ((1, 40, 40, 3), 75)])
def test_resample_backward(shape, factor):
    from neural_imaging_b200 import ops
    from neural_imaging_b200.tensor import as_device, zeros
    x = _x(shape, 9)
    dy = np.random.RandomState(10).normal(size=shape).astype(np.float32)
    g = ops.ResampleOp().backward(as_device(x), as_device(dy), zeros(shape), factor).cpu().numpy()
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    g_ref, = torch.autograd.grad(R.manipulation_resample(xt, factor), xt, torch.tensor(dy, dtype=torch.float64))
    assert_parity(g, g_ref.numpy(), tol=1e-5, what='parity')


@pytest.mark.parametrize('shape,k', [((2, 16, 16, 3), 2), ((1, 15, 21, 3), 2), ((2, 12, 12, 3), 3), ((1, 8, 8, 3), 1)])
def test_avgpool(shape, k):
    from neural_imaging_b200 import ops
    from neural_imaging_b200.tensor import as_device
    x = _x(shape, 11)
    y = ops.avgpool_fwd(as_device(x), k).cpu().numpy()
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    yt = R.avg_pool(xt, k)
    assert_parity(y, yt.detach().numpy(), tol=1e-6, what='parity')
    dy = np.random.RandomState(12).normal(size=y.shape).astype(np.float32)
    g_ref, = torch.autograd.grad(yt, xt, torch.tensor(dy, dtype=torch.float64))
    g = ops.avgpool_bwd(as_device(dy), shape, k).cpu().numpy()
    assert_parity(g, g_ref.numpy(), tol=1e-6, what='parity')


@pytest.mark.parametrize('b,hw,names', [(3, 64, ('sharpen', 'resample', 'gaussian', 'jpeg')), (2, 96, ('sharpen', 'resample', 'gaussian')),
                                         (1, 256, ('gaussian', 'resample')), (2, 40, ('resample', 'gamma', 'sharpen', 'median')),
                                         (2, 72, ('gaussian3', 'resample70'))])
def test_fused_pooled_stack_equals_operator_by_operator(b, hw, names):
    """ni_manip_stack_pool2_fwd / _bwd (manipulations with the 2x2 average pooling folded into the store) against the
    stand-alone kernels followed by ni_avgpool: run_manipulations + run_downsampling of workflows/manipulation_classification.py:199-245.
    Sizes cover whole tiles, partial tiles (96, 40, 72 are not multiples of the 32-pixel tile) and image borders inside one tile."""
    from collections import OrderedDict
    from neural_imaging_b200 import _lib, ops
    from neural_imaging_b200.tensor import as_device, empty, ptr, stream, zeros
    makers = {'sharpen': (ops.SharpenOp, 1.0), 'resample': (ops.ResampleOp, 50), 'gaussian': (lambda: ops.GaussianOp(5), 0.83),
              'jpeg': (ops.JpegOp, 80), 'gamma': (ops.GammaOp, 3.0), 'median': (ops.MedianOp, 3), 'gaussian3': (lambda: ops.GaussianOp(3), 1.2),
              'resample70': (ops.ResampleOp, 70)}
    opsd, strengths = OrderedDict(), {}
    for nme in names:
        opsd[nme] = makers[nme][0]()
        strengths[nme] = makers[nme][1]
    rs = np.random.RandomState(b * 1000 + hw)
    Y = as_device(rs.uniform(-0.1, 1.1, size=(b, hw, hw, 3)).astype(np.float32))
    k = len(names) + 1
    dc = as_device(rs.normal(size=(k * b, hw // 2, hw // 2, 3)).astype(np.float32))
    dY0 = rs.normal(size=(b, hw, hw, 3)).astype(np.float32)
    L = _lib.lib()
    # operator by operator
    m = empty((k * b, hw, hw, 3))
    ops.copy_into(m[:b], Y)
    for i, (nme, op) in enumerate(opsd.items()):
        op.forward(Y, m[(i + 1) * b:(i + 2) * b], strengths[nme], training=True)
    c_ref = ops.avgpool_fwd(m, 2)
    dm = ops.avgpool_bwd(dc, m.shape, 2)
    dY_ref = as_device(dY0.copy())
    L.ni_axpy(ptr(dY_ref), ptr(dm[:b]), 1.0, dY_ref.numel(), stream())
    for i, (nme, op) in enumerate(opsd.items()):
        if op.has_grad:
            op.backward(Y, dm[(i + 1) * b:(i + 2) * b], dY_ref, strengths[nme])
    # fused
    assert ops.PooledStack.applicable('pool:2', hw, hw) and not ops.PooledStack.applicable('bilinear', hw, hw)
    stack = ops.PooledStack(opsd)
    plan = stack.plan(strengths, hw)
    assert plan['slots'][0] == 0 and len(plan['rest']) == sum(n in ('jpeg', 'gamma', 'median', 'resample70') for n in names)
    scratch = empty((b, hw, hw, 3))
    c = stack.forward(Y, zeros((k * b, hw // 2, hw // 2, 3)), plan, strengths, scratch, training=True)
    dY = stack.backward(Y, dc, as_device(dY0.copy()), plan, strengths, scratch)
    torch.cuda.synchronize()
    c, c_ref, dY, dY_ref = c.cpu().numpy(), c_ref.cpu().numpy(), dY.cpu().numpy(), dY_ref.cpu().numpy()
    # forward: same arithmetic operation for operation -> float32 rounding only (FMA contraction may differ between the two kernels)
    assert rel_err(c, c_ref) < 2e-6, rel_err(c, c_ref)
    for i in range(k):
        assert rel_err(c[i * b:(i + 1) * b], c_ref[i * b:(i + 1) * b]) < 2e-6, (i, names)
    # backward: the reference chain scatters with atomics (order varies), the fused kernel gathers
    assert rel_err(dY, dY_ref) < 5e-6, rel_err(dY, dY_ref)
