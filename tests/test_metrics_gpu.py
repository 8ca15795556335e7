"""SSIM kernel (csrc/metrics.cu) against the oracle restatements of tf.image.ssim and skimage structural_similarity."""
import numpy as np
import pytest

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
