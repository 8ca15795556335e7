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


class PooledStack:
    """run_manipulations + run_downsampling('pool:2') without the full-resolution stack (ni_manip_stack_pool2_fwd / _bwd).

    plan() decides per step which class slots the fused kernels cover — native, sharpen, resample at factor 50 (any strength that
    `ResampleOp._small` maps to H/2), gaussian 3x3 / 5x5 — and which operators (jpeg, awgn, gamma, median, other factors) go through
    their stand-alone kernels into one image-sized scratch buffer followed by a per-slot pooling."""

    def __init__(self, ops_by_name):
        self.ops = ops_by_name            # OrderedDict name -> ManipOp, class slot = position + 1
        self._mask = None

    @staticmethod
    def applicable(downsampling, h, w):
        return downsampling in ('pool', 'pool:2') and h % 2 == 0 and w % 2 == 0 and h >= 8 and w >= 8 and h == w

    def plan(self, strengths, h):
        slots = [0, -1, -1, -1]           # native, sharpen, resample, gaussian
        sharp = gauss = None
        gk = 0
        rest = []
        for i, (name, op) in enumerate(self.ops.items()):
            st = strengths[name]
            if isinstance(op, SharpenOp):
                slots[1], sharp = i + 1, sharpen_filter(st)
            elif isinstance(op, ResampleOp) and 2 * ResampleOp._small(h, st) == h:
                slots[2] = i + 1
            elif isinstance(op, GaussianOp) and op.kernel in (3, 5):
                slots[3], gauss, gk = i + 1, kernels.gkern(op.kernel, st), op.kernel
            else:
                rest.append((i + 1, name, op))
        return {'slots': np.asarray(slots, dtype=np.int32), 'sharp': sharp, 'gauss': gauss, 'gk': gk, 'rest': rest}

    def forward(self, Y, c, plan, strengths, scratch, training=False, mask=None):
        """c (n_classes * B, H/2, W/2, 3) <- pooled stack of Y (B, H, W, 3); scratch: (B, H, W, 3) buffer for the stand-alone operators;
        mask: (B, H, W) uint8 buffer that receives the gaussian slot's clip mask (needed by backward)."""
        b, h, w = _nhw3(Y)
        L = _lib.lib()
        _, ps = _f32(plan['sharp']) if plan['sharp'] is not None else (None, None)
        _, pg = _f32(plan['gauss']) if plan['gauss'] is not None else (None, None)
        self._mask = mask if (training and plan['slots'][3] >= 0) else None
        if training and plan['slots'][3] >= 0 and mask is None:
            self._mask = empty((b, h, w), torch.uint8)
        L.ni_manip_stack_pool2_fwd(ptr(Y), ptr(c), ptr(self._mask), b, h, w, len(self.ops) + 1, plan['slots'].ctypes.data, ps, pg, plan['gk'], stream())
        for slot, name, op in plan['rest']:
            op.forward(Y, scratch, strengths[name], training=training)
            L.ni_avgpool_fwd(ptr(scratch), ptr(c[slot * b:(slot + 1) * b]), b, h, w, 2, stream())
        return c

    def backward(self, Y, dc, dY, plan, strengths, scratch):
        """dY += d(pooled stack)/dY applied to dc."""
        b, h, w = _nhw3(Y)
        L = _lib.lib()
        _, pg = _f32(plan['gauss']) if plan['gauss'] is not None else (None, None)
        L.ni_manip_stack_pool2_bwd(ptr(self._mask), ptr(dc), ptr(dY), b, h, w, len(self.ops) + 1, plan['slots'].ctypes.data, pg, plan['gk'], 1, stream())
        for slot, name, op in plan['rest']:
            if op.has_grad:
                L.ni_avgpool_bwd(ptr(dc[slot * b:(slot + 1) * b]), ptr(scratch), b, h, w, 2, stream())
                op.backward(Y, scratch, dY, strengths[name])
        return dY


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


class StructuralLoss:
    """SSIM / MS-SSIM image losses of the ISPs with an explicit backward (reference helpers/tf_helpers.py:39-44, selected by
    NIPModel.construct_loss, models/pipelines.py:53-63): loss = mean_n 255 (1 - tf.image.ssim[_multiscale](a, b, 1.0)).

    forward() adds `loss * loss_scale` to a device accumulator and keeps the image pyramid and the per-(image, channel) gradient
    coefficients; backward() writes (or accumulates) `grad_scale * d loss / d a`. Every buffer lives in the given Workspace, so the
    pair is CUDA-graph safe. Kernels: ni_ssim_stats, ni_msssim_combine, ni_ssim_bwd (csrc/metrics.cu) + the average-pool kernels."""
    WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)     # TF 2.1 _MSSSIM_WEIGHTS
    K1, K2, SIZE, SIGMA = 0.01, 0.03, 11, 1.5

    def __init__(self, multiscale=False, workspace=None, tag='sloss'):
        from .tensor import Workspace
        self.levels = len(self.WEIGHTS) if multiscale else 1
        self._ws = workspace if workspace is not None else Workspace()
        self._tag = tag
        x = np.arange(self.SIZE, dtype=np.float64) - (self.SIZE - 1) / 2.0
        g = np.exp(-0.5 * x * x / self.SIGMA ** 2)
        self._win, self._pwin = _f32(g / g.sum())
        self._wts, self._pwts = _f32(self.WEIGHTS)
        self._pyr = None

    def _check(self, a, b):
        if a.shape != b.shape or a.dim() != 4:
            raise ValueError('SSIM loss: expected two (N,H,W,C) tensors of the same shape')
        n, h, w, c = (int(v) for v in a.shape)
        div = 2 ** (self.levels - 1)
        if self.levels > 1 and (h % div or w % div or c != 3):
            raise ValueError('MS-SSIM loss on the B200 path needs (N,H,W,3) images with H and W divisible by {}'.format(div))
        if h // div < self.SIZE or w // div < self.SIZE:
            raise ValueError('SSIM loss: the {0} x {0} window does not fit the coarsest scale ({1} x {2})'.format(self.SIZE, h // div, w // div))
        return n, h, w, c

    def forward(self, a, b, acc, loss_scale=1.0, grad_scale=1.0):
        n, h, w, c = self._check(a, b)
        L, s, ws, t = _lib.lib(), stream(), self._ws, self._tag
        stats = ws.get(t + '_stats', (self.levels, n * c, 2))
        self._coef = ws.get(t + '_coef', (self.levels, n * c, 2))
        self._pyr = [(a, b)]
        c1, c2 = self.K1 ** 2, self.K2 ** 2
        for l in range(self.levels):
            if l > 0:
                pa, pb = self._pyr[-1]
                shape = (n, pa.shape[1] // 2, pa.shape[2] // 2, c)
                self._pyr.append((avgpool_fwd(pa, 2, out=ws.get('{}_a{}'.format(t, l), shape)),
                                  avgpool_fwd(pb, 2, out=ws.get('{}_b{}'.format(t, l), shape))))
            la, lb = self._pyr[-1]
            L.ni_ssim_stats(ptr(la), ptr(lb), ptr(stats[l]), n, int(la.shape[1]), int(la.shape[2]), c, self._pwin, self.SIZE, 1.0, c1, c2, s)
        L.ni_msssim_combine(ptr(stats), ptr(self._coef), ptr(acc), n, c, self.levels, self._pwts, float(loss_scale), float(grad_scale), s)
        return acc

    def backward(self, da, accumulate=False):
        """da (+)= grad_scale * d loss / d a for the (a, b) of the last forward()."""
        if self._pyr is None:
            raise RuntimeError('StructuralLoss.backward() before forward()')
        L, s, ws, t = _lib.lib(), stream(), self._ws, self._tag
        c1, c2 = self.K1 ** 2, self.K2 ** 2
        n, c = int(da.shape[0]), int(da.shape[3])
        g_coarse = None
        for l in range(self.levels - 1, -1, -1):
            la, lb = self._pyr[l]
            if l == 0:
                g, acc_flag = da, accumulate
            else:
                g, acc_flag = ws.get('{}_g{}'.format(t, l), la.shape), False
            if g_coarse is not None:            # chain rule through the 2 x 2 average pooling of the next coarser scale
                if l == 0 and accumulate:
                    tmp = avgpool_bwd(g_coarse, la.shape, 2, out=ws.get(t + '_g0', la.shape))
                    L.ni_axpy(ptr(g), ptr(tmp), 1.0, g.numel(), s)
                else:
                    avgpool_bwd(g_coarse, la.shape, 2, out=g)
                acc_flag = True
            L.ni_ssim_bwd(ptr(la), ptr(lb), ptr(self._coef[l]), ptr(g), int(acc_flag), n, int(la.shape[1]), int(la.shape[2]), c, self._pwin,
                          self.SIZE, 1.0, c1, c2, s)
            g_coarse = g
        return da
