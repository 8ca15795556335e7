"""Device versions of the reference's helpers/tf_helpers.py operators (same names, argument meaning and error
behaviour; inputs may be NumPy arrays or device tensors, outputs are device tensors with ``.numpy()``).

manipulation_* / soft_quantization / losses: helpers/tf_helpers.py:31-44,68-184,271-287.
"""
import numpy as np

from .. import _lib, ops
from ..tensor import as_device, empty, ptr, stream, wrap, zeros

activation_mapping = {'leaky_relu': 'leaky_relu', 'relu': 'relu', 'tanh': 'tanh', 'sigmoid': 'sigmoid'}


def _run(op, x, strength):
    x = as_device(x)
    if x.dim() == 3:
        x = x.unsqueeze(0)
    return wrap(op.forward(x, empty(x.shape), strength))


def manipulation_sharpen(x, strength=1, hsv=True):
    if not hsv:
        raise NotImplementedError('only the hsv=True variant used by the workflow is implemented')
    return _run(ops.SharpenOp(), x, strength)


def manipulation_resample(x, factor=50, method='bilinear'):
    if method != 'bilinear':
        raise NotImplementedError('only bilinear resampling is implemented')
    return _run(ops.ResampleOp(), x, factor)


def manipulation_gaussian(x, kernel, std, skip_clip=False):
    x = as_device(x)
    op = ops.GaussianOp(kernel)
    if skip_clip:
        from ..helpers import kernels
        n, h, w = ops._nhw3(x)
        f, pf = ops._f32(kernels.gkern(int(kernel), std))
        y = empty(x.shape)
        _lib.lib().ni_manip_gaussian_fwd(ptr(x), ptr(y), None, n, h, w, pf, int(kernel), 0, stream())
        return wrap(y)
    return _run(op, x, std)


def manipulation_awgn(x, strength=0.025, noise=None):
    op = ops.AwgnOp()
    op.noise = None if noise is None else as_device(noise)
    return _run(op, x, strength * 255.0)


def manipulation_gamma(x, strength=2.0):
    return _run(ops.GammaOp(), x, strength)


def manipulation_median(x, kernel=3):
    return _run(ops.MedianOp(), x, kernel)


def mse(a, b):
    a, b = as_device(a), as_device(b)
    return wrap(ops.image_loss(a, b, 'L2').reshape(()))


def mae(a, b):
    a, b = as_device(a), as_device(b)
    return wrap(ops.image_loss(a, b, 'L1').reshape(()))


def _structural(a, b, multiscale):
    a, b = as_device(a), as_device(b)
    if a.dim() == 3:
        a, b = a.unsqueeze(0), b.unsqueeze(0)
    acc = zeros((1,))
    ops.StructuralLoss(multiscale).forward(a, b, acc)
    return wrap(acc.reshape(()))


def ssim_loss(a, b):
    """mean(255 (1 - tf.image.ssim(a, b, 1.0))) (helpers/tf_helpers.py:39-40)."""
    return _structural(a, b, False)


def msssim_loss(a, b):
    """mean(255 (1 - tf.image.ssim_multiscale(a, b, 1.0))) (helpers/tf_helpers.py:43-44)."""
    return _structural(a, b, True)


def quantize_and_clip(x):
    """clip(soft_quantization(x), 0, 1) == awgn with zero noise strength."""
    return manipulation_awgn(x, 0.0)
