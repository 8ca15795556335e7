"""Image-quality metrics of the B200 path — mirror of the reference's `helpers/metrics.py:9-94` (ssim / psnr / mse / mae / batch,
which wrap scikit-image) plus `tf.image.ssim` as the learned codec's training step reports it (models/compression.py:89).

SSIM runs in one fused kernel (`ni_ssim`, csrc/metrics.cu) for both flavours; the arrays may be numpy or device tensors, results are
numpy floats / arrays like the reference's.
"""
import numpy as np
import torch

from .. import _lib
from ..tensor import as_device, device, ptr, stream

_K1, _K2 = 0.01, 0.03


def _gauss_window(size=11, sigma=1.5):
    x = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    g = np.exp(-0.5 * x * x / (sigma * sigma))
    return (g / g.sum()).astype(np.float32)


def _ssim_device(a, b, win, cov_norm, data_range=1.0):
    a, b = as_device(a), as_device(b)
    if a.dim() == 3:
        a, b = a.unsqueeze(0), b.unsqueeze(0)
    if a.shape != b.shape or a.dim() != 4:
        raise ValueError('ssim: expected two (N)HWC arrays of the same shape')
    n, h, w, c = a.shape
    out = torch.empty((n,), dtype=torch.float32, device=device())
    win = np.ascontiguousarray(win, dtype=np.float32)
    _lib.lib().ni_ssim(ptr(a), ptr(b), ptr(out), n, h, w, c, win.ctypes.data, len(win), float(cov_norm), float((_K1 * data_range) ** 2),
                       float((_K2 * data_range) ** 2), stream())
    return out


def ssim_tf(a, b, max_val=1.0):
    """tf.image.ssim(a, b, max_val): per-image values as a device tensor (11 x 11 Gaussian, sigma 1.5, VALID)."""
    return _ssim_device(a, b, _gauss_window(), 1.0, max_val)


def ssim(a, b):
    """skimage structural_similarity(a, b, multichannel=True, data_range=1) (reference helpers/metrics.py:9-26): 7 x 7 uniform window,
    sample covariance, 3-pixel border cropped. 3-D inputs -> float, 4-D -> array of per-image values (1 x H x W x C is squeezed)."""
    a, b = (t if isinstance(t, torch.Tensor) else np.asarray(t) for t in (a, b))
    if a.ndim == 4 and a.shape[0] == 1:
        a = a[0]
    if b.ndim == 4 and b.shape[0] == 1:
        b = b[0]
    if a.ndim not in (3, 4) or a.ndim != b.ndim:
        raise ValueError('Incompatible tensor shapes! Expected 3- or 4-dim arrays.')
    v = _ssim_device(a, b, np.full((7,), 1.0 / 7.0), 49.0 / 48.0).cpu().numpy().astype(np.float64)
    return float(v[0]) if a.ndim == 3 else v


def _per_image(a, b, fn):
    a, b = np.asarray(a), np.asarray(b)
    if a.ndim == 4 and a.shape[0] == 1:
        a = a[0]
    if b.ndim == 4 and b.shape[0] == 1:
        b = b[0]
    if a.ndim == 3 and b.ndim == 3:
        return fn(a.astype(np.float64), b.astype(np.float64))
    if a.ndim == 4 and b.ndim == 4:
        return np.array([fn(a[i].astype(np.float64), b[i].astype(np.float64)) for i in range(a.shape[0])])
    raise ValueError('Incompatible tensor shapes! Expected 3- or 4-dim arrays.')


def mse(a, b):
    return _per_image(a, b, lambda x, y: float(np.mean((x - y) ** 2)))


def mae(a, b):
    return _per_image(a, b, lambda x, y: float(np.mean(np.abs(x - y))))


def psnr(a, b):
    """skimage peak_signal_noise_ratio(a, b, data_range=1) = 10 log10(1 / mse)."""
    def f(x, y):
        m = float(np.mean((x - y) ** 2))
        return float(10.0 * np.log10(1.0 / m)) if m > 0 else float('inf')
    return _per_image(a, b, f)


def batch(a, b, metric=ssim):
    a, b = np.asarray(a), np.asarray(b)
    assert a.ndim == 4 and b.ndim == 4, 'Input arrays need to be 4-dim: batch, height, width, channels'
    assert len(a) == len(b), 'Image batches must be of the same length'
    return np.mean([metric(a[r], b[r]) for r in range(len(a))])
