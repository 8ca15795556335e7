"""Thin Python wrappers over the C-ABI kernels (device tensors in / out). Each *Op class is a manipulation with an
explicit forward and backward, used by the workflow's hand-ordered training step."""
import ctypes

import numpy as np
import torch

from . import _lib
from .helpers import kernels
from .tensor import as_device, empty, ptr, stream, zeros


def _f32(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _nhw3(x):
    if x.dim() != 4 or x.shape[-1] != 3:
        raise ValueError('expected an (N,H,W,3) tensor, got {}'.format(tuple(x.shape)))
    return int(x.shape[0]), int(x.shape[1]), int(x.shape[2])


QUANT_MODES = {'soft': 0, 'sin': 1, 'harmonic': 2}


# ------------------------------------------------------------------------------------------------ dJPEG
def djpeg_fwd(x, q_luma, q_chroma, mode='soft', out=None, want_coeffs=False):
    n, h, w = _nhw3(x)
    if h % 8 or w % 8:
        raise ValueError('dJPEG needs H and W to be multiples of 8, got {}x{}'.format(h, w))
    y = empty(x.shape) if out is None else out
    X = empty((3 * n * (h // 8) * (w // 8), 8, 8)) if want_coeffs else None
    ql, pl = _f32(q_luma)
    qc, pc = _f32(q_chroma)
    _lib.lib().ni_djpeg_fwd(ptr(x), ptr(y), ptr(X), n, h, w, pl, pc, QUANT_MODES[mode], stream())
    return (y, X) if want_coeffs else y


def djpeg_bwd(x, dy, q_luma, q_chroma, mode='soft', out=None):
    n, h, w = _nhw3(x)
    dx = empty(x.shape) if out is None else out
    ql, pl = _f32(q_luma)
    qc, pc = _f32(q_chroma)
    _lib.lib().ni_djpeg_bwd(ptr(x), ptr(dy), ptr(dx), n, h, w, pl, pc, QUANT_MODES[mode], stream())
    return dx


# ------------------------------------------------------------------------------------------------ manipulations
def sharpen_filter(strength):
    """3x3 H/V filter of manipulation_sharpen (reference helpers/tf_helpers.py:156-160)."""
    gk = np.array([[-0.0833, -0.1667, -0.0833], [-0.1667, 0, -0.1667], [-0.0833, -0.1667, -0.0833]])
    gk = strength * gk / np.abs(gk.sum())
    gk[1, 1] = strength + 1
    return gk.astype(np.float32)


class ManipOp:
    """A manipulation y = f(x, strength) on (N,H,W,3) with forward into a caller-provided slot and a backward that
    ACCUMULATES into dx. `has_grad` False means the reference's TF version passes no gradient (see SharpenOp)."""
    has_grad = True

    def forward(self, x, y, strength, training=False):
        raise NotImplementedError

    def backward(self, x, dy, dx, strength):
        raise NotImplementedError


class SharpenOp(ManipOp):
    """manipulation_sharpen(hsv=True). tf.image.rgb_to_hsv / hsv_to_rgb are registered NotDifferentiable in the
    reference's pinned TensorFlow 2.1 (python/ops/image_ops_impl.py), so this branch contributes NO gradient to the
    ISP in the reference training step; we reproduce that."""
    has_grad = False

    def forward(self, x, y, strength, training=False):
        n, h, w = _nhw3(x)
        f, pf = _f32(sharpen_filter(strength))
        _lib.lib().ni_manip_sharpen_fwd(ptr(x), ptr(y), n, h, w, pf, stream())
        return y

    def backward(self, x, dy, dx, strength):
        return dx


class ResampleOp(ManipOp):
    """manipulation_resample: bilinear down to H*int(f)//100 then back up (helpers/tf_helpers.py:68-76)."""

    def __init__(self):
        self._tmp = None

    @staticmethod
    def _small(h, factor):
        if 0 < factor <= 1:
            factor = 100 * factor
        return h * int(factor) // 100

    def forward(self, x, y, strength, training=False):
        n, h, w = _nhw3(x)
        s = self._small(h, strength)       # reference uses shape[1] for both dims
        tmp = empty((n, s, s, 3))
        L = _lib.lib()
        L.ni_resize_bilinear_fwd(ptr(x), ptr(tmp), n, h, w, s, s, 0, stream())
        L.ni_resize_bilinear_fwd(ptr(tmp), ptr(y), n, s, s, h, h, 0, stream())
        return y

    def backward(self, x, dy, dx, strength):
        n, h, w = _nhw3(x)
        s = self._small(h, strength)
        L = _lib.lib()
        dtmp = empty((n, s, s, 3))
        L.ni_fill(ptr(dtmp), 0.0, dtmp.numel(), stream())
        L.ni_resize_bilinear_bwd(ptr(dy), ptr(dtmp), n, s, s, h, h, 1.0, stream())
        L.ni_resize_bilinear_bwd(ptr(dtmp), ptr(dx), n, h, w, s, s, 1.0, stream())   # scatter-add == accumulate
        return dx


class GaussianOp(ManipOp):
    """manipulation_gaussian(x, 5, std) (helpers/tf_helpers.py:113-125)."""

    def __init__(self, kernel=5):
        self.kernel = int(kernel)
        self._mask = None

    def forward(self, x, y, strength, training=False):
        n, h, w = _nhw3(x)
        f, pf = _f32(kernels.gkern(self.kernel, strength))
        self._mask = empty((n, h, w), torch.uint8) if training else None
        _lib.lib().ni_manip_gaussian_fwd(ptr(x), ptr(y), ptr(self._mask), n, h, w, pf, self.kernel, 1, stream())
        return y

    def backward(self, x, dy, dx, strength):
        n, h, w = _nhw3(x)
        f, pf = _f32(kernels.gkern(self.kernel, strength))
        _lib.lib().ni_manip_gaussian_bwd(ptr(dy), ptr(self._mask), ptr(dx), n, h, w, pf, self.kernel, 1.0, 1, stream())
        return dx


class JpegOp(ManipOp):
    """models.jpeg.differentiable_jpeg used as a manipulation (workflows/manipulation_classification.py:118)."""

    def forward(self, x, y, strength, training=False):
        from .compression.jpeg_helpers import jpeg_qtable
        q = int(strength)
        return djpeg_fwd(x, jpeg_qtable(q, 0), jpeg_qtable(q, 1), 'soft', out=y)

    def backward(self, x, dy, dx, strength):
        from .compression.jpeg_helpers import jpeg_qtable
        q = int(strength)
        g = djpeg_bwd(x, dy, jpeg_qtable(q, 0), jpeg_qtable(q, 1), 'soft')
        _lib.lib().ni_axpy(ptr(dx), ptr(g), 1.0, g.numel(), stream())
        return dx


class AwgnOp(ManipOp):
    """manipulation_awgn(x, strength/255) (helpers/tf_helpers.py:79-82); noise from on-device Philox unless injected."""

    def __init__(self):
        self.noise = None          # inject a (N,H,W,3) N(0,1) tensor here for parity tests
        self.seed = 0x5EED
        self._last_seed = None

    def forward(self, x, y, strength, training=False):
        n, h, w = _nhw3(x)
        self._last_seed = self.seed
        self.seed += 1
        _lib.lib().ni_manip_awgn_fwd(ptr(x), ptr(self.noise), ptr(y), n, h, w, float(strength) / 255.0, self._last_seed, stream())
        return y

    def backward(self, x, dy, dx, strength):
        n, h, w = _nhw3(x)
        _lib.lib().ni_manip_awgn_bwd(ptr(x), ptr(self.noise), ptr(dy), ptr(dx), n, h, w, float(strength) / 255.0,
                                     self._last_seed, 1, stream())
        return dx


class GammaOp(ManipOp):
    def forward(self, x, y, strength, training=False):
        n, h, w = _nhw3(x)
        _lib.lib().ni_manip_gamma_fwd(ptr(x), ptr(y), n, h, w, float(strength), stream())
        return y

    def backward(self, x, dy, dx, strength):
        n, h, w = _nhw3(x)
        _lib.lib().ni_manip_gamma_bwd(ptr(x), ptr(dy), ptr(dx), n, h, w, float(strength), 1, stream())
        return dx


class MedianOp(ManipOp):
    @staticmethod
    def _k(strength):
        k = int(strength)
        if k % 2 == 0:
            k += 1
        return max(k, 1)

    def forward(self, x, y, strength, training=False):
        n, h, w = _nhw3(x)
        _lib.lib().ni_manip_median_fwd(ptr(x), ptr(y), n, h, w, self._k(strength), stream())
        return y

    def backward(self, x, dy, dx, strength):
        n, h, w = _nhw3(x)
        _lib.lib().ni_manip_median_bwd(ptr(x), ptr(dy), ptr(dx), n, h, w, self._k(strength), stream())
        return dx


def avgpool_fwd(x, k, out=None):
    n, h, w = _nhw3(x)
    y = empty((n, -(-h // k), -(-w // k), 3)) if out is None else out
    _lib.lib().ni_avgpool_fwd(ptr(x), ptr(y), n, h, w, k, stream())
    return y


def avgpool_bwd(dy, shape, k, out=None):
    n, h, w = int(shape[0]), int(shape[1]), int(shape[2])
    dx = empty((n, h, w, 3)) if out is None else out
    _lib.lib().ni_avgpool_bwd(ptr(dy), ptr(dx), n, h, w, k, stream())
    return dx


def copy_into(dst, src):
    """dst <- src with our own kernel (keeps torch's elementwise kernels off the hot path)."""
    _lib.lib().ni_affine(ptr(src), ptr(dst), 1.0, 0.0, 0, src.numel(), stream())
    return dst


def image_loss(a, b, kind='L2'):
    """tf_helpers.mse / mae: scalar device tensor."""
    acc = zeros((1,))
    _lib.lib().ni_image_loss(ptr(a), ptr(b), ptr(acc), a.numel(), 0 if kind == 'L2' else 1, stream())
    return acc / float(a.numel())
