"""SSIM kernel (csrc/metrics.cu) against the oracle restatements of tf.image.ssim and skimage structural_similarity."""
import numpy as np
import pytest

from conftest import assert_parity
from oracle import ref_ops as R


def _pair(n, h, w, seed):
    rs = np.random.RandomState(seed)
    a = rs.uniform(size=(n, h, w, 3)).astype(np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    a = (0.5 * a + 0.5 * (np.sin(yy / 7.0)[None, :, :, None] * np.cos(xx / 5.0)[None, :, :, None] * 0.5 + 0.5)).astype(np.float32)
    b = np.clip(a + rs.normal(scale=0.05, size=a.shape), 0, 1).astype(np.float32)
    return a, b


def test_oracle_ssim_sanity():
    a, b = _pair(2, 40, 48, 0)
    assert np.allclose(R.ssim_tf(a, a), 1.0) and abs(R.ssim_skimage(a[0], a[0]) - 1.0) < 1e-12
    v = R.ssim_tf(a, b)
    assert ((v > 0.3) & (v < 0.999)).all()
    # constant images: luminance term only
    c0, c1 = np.full((1, 32, 32, 3), 0.2), np.full((1, 32, 32, 3), 0.6)
    expect = (2 * 0.2 * 0.6 + 1e-4) / (0.04 + 0.36 + 1e-4)
    assert abs(R.ssim_tf(c0, c1)[0] - expect) < 1e-9 and abs(R.ssim_skimage(c0[0], c1[0]) - expect) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(2, 128, 128), (3, 37, 53), (1, 11, 11), (1, 256, 256)])
def test_ssim_tf_matches_oracle(shape):
    from neural_imaging_b200.helpers import metrics
    n, h, w = shape
    a, b = _pair(n, h, w, 1)
    got = metrics.ssim_tf(a, b).cpu().numpy()
    ref = R.ssim_tf(a, b)
    assert np.max(np.abs(got - ref)) < 1e-5, (got, ref)          # 1e-5 absolute on values in (0, 1]


@pytest.mark.gpu
def test_ssim_skimage_flavour_and_wrappers():
    from neural_imaging_b200.helpers import metrics
    a, b = _pair(3, 64, 80, 2)
    got = metrics.ssim(a, b)
    ref = np.array([R.ssim_skimage(a[i], b[i]) for i in range(3)])
    assert got.shape == (3,) and np.max(np.abs(got - ref)) < 1e-5
    assert abs(metrics.ssim(a[:1], b[:1]) - ref[0]) < 1e-5          # 1 x H x W x C is squeezed to one float
    assert abs(metrics.batch(a, b, metrics.ssim) - ref.mean()) < 1e-5
    m = np.mean((a[0].astype(np.float64) - b[0]) ** 2)
    assert abs(metrics.psnr(a[0], b[0]) - 10 * np.log10(1 / m)) < 1e-9 and abs(metrics.mse(a, b)[0] - m) < 1e-12
    with pytest.raises(ValueError):
        metrics.ssim(a[0, 0], b[0, 0])
    from neural_imaging_b200 import _lib
    with pytest.raises(_lib.NIError):
        metrics.ssim_tf(a[:, :8, :8], b[:, :8, :8])                 # window larger than the image


@pytest.mark.gpu
def test_dcn_training_step_reports_tf_ssim():
    from neural_imaging_b200.models.compression import TwitterDCN
    a, b = _pair(2, 64, 64, 3)
    v = float(TwitterDCN.ssim(a, b).numpy())
    assert abs(v - R.ssim_tf(a, b).mean()) < 1e-5


# ---------------------------------------------------------------------------------------------- SSIM / MS-SSIM as ISP losses (a3)
def _loss_oracle(name, a, b, dt):
    import torch as T
    ta = T.tensor(a, dtype=dt, requires_grad=True)
    loss = getattr(R, name)(ta, T.tensor(b, dtype=dt))
    g, = T.autograd.grad(loss, ta)
    return float(loss.detach()), g.numpy()


@pytest.mark.gpu
@pytest.mark.parametrize('name,shape', [('ssim_loss', (2, 50, 45, 3)), ('ssim_loss', (3, 11, 11, 1)), ('ssim_loss', (1, 64, 64, 4)),
                                        ('msssim_loss', (2, 176, 192, 3)), ('msssim_loss', (1, 256, 256, 3))])
def test_structural_losses_forward_backward(name, shape):
    """helpers/tf_helpers.py:39-44 — value and gradient w.r.t. the first image against float64 autograd of the restatement."""
    import torch as T
    from neural_imaging_b200 import ops
    from neural_imaging_b200.helpers import tf_helpers
    from neural_imaging_b200.tensor import as_device, zeros
    rs = np.random.RandomState(5)
    a = rs.uniform(size=shape).astype(np.float32)
    b = np.clip(a + 0.15 * rs.normal(size=shape), 0, 1).astype(np.float32)
    l64, g64 = _loss_oracle(name, a, b, T.float64)
    l32, g32 = _loss_oracle(name, a, b, T.float32)
    got = float(getattr(tf_helpers, name)(a, b).numpy())
    assert abs(got - l64) <= max(1e-5 * abs(l64), 4 * abs(l32 - l64)), (got, l64, l32)
    op = ops.StructuralLoss(name == 'msssim_loss')
    da, db_ = as_device(a), as_device(b)
    acc = zeros((1,))
    op.forward(da, db_, acc, loss_scale=2.0, grad_scale=0.5)
    g = T.full(shape, 3.0, dtype=T.float32, device=da.device)
    op.backward(g, accumulate=True)                      # accumulate on top of a constant
    assert abs(float(acc.item()) - 2.0 * l64) <= 1e-4 * abs(l64)
    assert_parity(g.cpu().numpy() - 3.0, 0.5 * g64, 0.5 * g32, tol=2e-5, slack=6.0, what=name + ' grad (accumulate)')
    g2 = T.full(shape, 7.0, dtype=T.float32, device=da.device)
    op.backward(g2, accumulate=False)                    # overwrite
    assert_parity(g2.cpu().numpy(), 0.5 * g64, 0.5 * g32, tol=2e-5, slack=6.0, what=name + ' grad')
    assert float(getattr(tf_helpers, name)(a, a).numpy()) < 1e-3          # identical images: zero loss


@pytest.mark.gpu
def test_structural_loss_errors_and_nip_training():
    from neural_imaging_b200.helpers import tf_helpers
    from neural_imaging_b200.models import pipelines
    rs = np.random.RandomState(11)
    with pytest.raises(ValueError):
        tf_helpers.ssim_loss(np.zeros((1, 8, 8, 3), np.float32), np.zeros((1, 8, 8, 3), np.float32))       # window does not fit
    with pytest.raises(ValueError):
        tf_helpers.msssim_loss(np.zeros((1, 64, 64, 3), np.float32), np.zeros((1, 64, 64, 3), np.float32))  # coarsest scale 4 x 4
    with pytest.raises(ValueError):
        pipelines.ONet(loss_metric='bogus', patch_size=16)
    # an ISP trained with the SSIM loss: the step's loss equals the oracle's on the model output and goes down
    import torch as T
    from oracle import ref_models as M
    model = pipelines.UNet(loss_metric='SSIM', patch_size=32, seed=3)
    x = rs.uniform(size=(2, 32, 32, 4)).astype(np.float32)
    t = rs.uniform(size=(2, 64, 64, 3)).astype(np.float32)
    P = M.to_params(model._store.state_dict(), T.float64)
    yt = M.unet_forward(P, T.tensor(x, dtype=T.float64))
    ref = R.ssim_loss(yt, T.tensor(t, dtype=T.float64))
    gref = T.autograd.grad(ref, list(P.values()))
    first = float(model.training_step(x, t, learning_rate=1e-3).numpy())
    assert abs(first - float(ref)) < 1e-4 * float(ref)
    g = {p.name: p.grad.detach().cpu().numpy() for p in model._store.trainable}
    for (k, _), gr in zip(P.items(), gref):
        assert_parity(g[k], gr.numpy(), tol=5e-4, what='UNet SSIM-loss grad ' + k)
    for _ in range(5):
        last = float(model.training_step(x, t, learning_rate=1e-3).numpy())
    assert last < first
    v = float(model.loss(model.process(x), t).numpy())
    assert np.isfinite(v) and v < first
