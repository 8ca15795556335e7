"""Stand-alone layers with the reference's names and constructor arguments (`models/layers.py:12-258`), forward path on the device.

Inside the B200 models these layers are fused into their consumers (the constrained filter normalisation + mirrored padding inside
`FAN`, the soft-codebook quantiser + soft histogram inside `DCN`, the scalar rounding modes inside the differentiable-JPEG kernel, the
demosaicing layer inside `ClassicISP`); the classes here serve callers that use a layer on its own (notebooks, tests). They are
inference-only: gradients live in the models' hand-ordered backward passes.
"""
import numpy as np
import torch

from .. import _lib, nn
from ..helpers import kernels
from ..tensor import as_device, device, empty, ptr, stream, wrap, zeros

_SCALAR_MODES = {'round': 0, 'sin': 1, 'soft': 2, 'harmonic': 3, 'identity': 4}


class ConstrainedConv2D(object):
    """Residual filter of the forensic network (reference `models/layers.py:12-57`): raw (5,5,3,3) kernel initialised to the channel
    diagonal of [[0,0,0,0,0],[0,-1,-2,-1,0],[0,-2,12,-2,0],...]; on every call the centre taps are zeroed, each output channel is
    normalised to sum `filter_strength` and the centres set to -filter_strength; SYMMETRIC pad 2 + VALID convolution, no bias."""

    def __init__(self, filter_strength=100, trainable=True):
        self.filter_strength = float(filter_strength)
        self.trainable = trainable
        f = np.array([[0, 0, 0, 0, 0], [0, -1, -2, -1, 0], [0, -2, 12, -2, 0], [0, -1, -2, -1, 0], [0, 0, 0, 0, 0]])
        self._store = nn.ParamStore()
        self._conv = nn.Conv2D(self._store, 'constrained_conv2d', 5, 3, 3, padding='VALID', use_bias=False, pad_mode=_lib.PAD_SYMMETRIC,
                               explicit_pad=2, kernel_init=kernels.repeat_2dfilter(f, 3))
        self._store.finalize()
        self._nf = empty((5, 5, 3, 3))

    @property
    def kernel(self):
        return wrap(self._conv.w.value)

    def normalized_kernel(self):
        _lib.lib().ni_constrained_filter_fwd(ptr(self._conv.w.value), ptr(self._nf), 5, 3, self.filter_strength, stream())
        return wrap(self._nf)

    def __call__(self, x):
        x = as_device(x)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        n, h, w, c = x.shape
        if c != 3:
            raise ValueError('ConstrainedConv2D expects 3-channel images')
        self.normalized_kernel()
        y = empty((n, h, w, 3))
        _lib.lib().ni_cconv5_fwd(ptr(x.contiguous()), ptr(self._nf), ptr(y), n, h, w, stream())
        return wrap(y)

    call = __call__


class Quantization(object):
    """Reference `models/layers.py:60-172`. rounding: 'round' | 'sin' | 'soft' | 'harmonic' | 'identity' | 'soft-codebook'
    (t-Student weights for v > 0, Gaussian for v <= 0; float64 inside, straight-through hard values out)."""

    def __init__(self, rounding='soft', v=50, gamma=25, latent_bpf=4, trainable=False, taylor_terms=1):
        if rounding not in _SCALAR_MODES and rounding != 'soft-codebook':
            raise ValueError('Unsupported quantization: {}'.format(rounding))
        if trainable:
            raise NotImplementedError('trainable codebooks are not implemented on the B200 path')
        self.rounding, self.v, self.gamma, self.latent_bpf, self.taylor_terms, self.trainable = rounding, v, gamma, latent_bpf, taylor_terms, trainable
        qmin, qmax = -2 ** (latent_bpf - 1) + 1, 2 ** (latent_bpf - 1)
        self.codebook = np.arange(qmin, qmax + 1, dtype=np.float32).reshape((1, -1))
        self._codebook_dev = None

    def _cb(self):
        if self._codebook_dev is None:
            self._codebook_dev = torch.from_numpy(self.codebook.reshape(-1)).to(device())
        return self._codebook_dev

    def __call__(self, x, scale=None, hist=None):
        x = as_device(x)
        y = torch.empty_like(x)
        L = _lib.lib()
        if self.rounding == 'soft-codebook':
            cb = self._cb()
            L.ni_latent_softcodebook_fwd(ptr(x), ptr(scale), ptr(cb), ptr(y), ptr(hist), x.numel(), cb.numel(), float(self.v), float(self.gamma), stream())
        else:
            L.ni_quantize_scalar(ptr(x), ptr(y), x.numel(), _SCALAR_MODES[self.rounding], int(self.taylor_terms), stream())
        return wrap(y)

    call = __call__


class DiscreteLatent(object):
    """Quantisation layer with an entropy estimate (reference `models/layers.py:175-203`): latent * scaling_factor -> Quantization ->
    soft-histogram entropy (`helpers/tf_helpers.py:290-333`) of the quantised values. Returns (latent, entropy)."""

    def __init__(self, rounding='soft', v=50, gamma=25, latent_bpf=4, trainable_codebook=False, trainable_scale=True):
        if rounding not in {'round', 'sin', 'soft', 'identity', 'harmonic', 'soft-codebook'}:
            raise ValueError('Unsupported quantization: {}'.format(rounding))
        self.trainable_scale, self.rounding, self.v, self.gamma, self.latent_bpf = trainable_scale, rounding, v, gamma, latent_bpf
        self.trainable_codebook = trainable_codebook
        self.quantization = Quantization(rounding, v, gamma, latent_bpf, trainable_codebook)
        self.scaling_factor = torch.ones((1,), dtype=torch.float32, device=device()) if trainable_scale else None

    def __call__(self, inputs):
        x = as_device(inputs)
        L, st = _lib.lib(), stream()
        cb = self.quantization._cb()
        n = x.numel()
        fused = {'soft-codebook': 0, 'sin': 1, 'soft': 2, 'identity': 3}
        if self.rounding == 'soft-codebook':
            hist = zeros((cb.numel(),), torch.float64)
            latent = self.quantization(x, scale=self.scaling_factor, hist=hist)
        elif self.rounding in fused:       # scaling, scalar rounding and the soft histogram of the quantised values in one kernel
            hist = zeros((cb.numel(),), torch.float64)
            latent = torch.empty_like(x)
            L.ni_latent_quantise_fwd(ptr(x), ptr(self.scaling_factor) if self.scaling_factor is not None else None, ptr(cb), ptr(latent), ptr(hist),
                                     n, cb.numel(), float(self.v), float(self.gamma), fused[self.rounding], st)
            latent = wrap(latent)
        else:
            latent = self.quantization(x * self.scaling_factor if self.scaling_factor is not None else x)
            # the entropy estimate always uses the soft histogram of the (quantised) values
            hist = zeros((cb.numel(),), torch.float64)
            tmp = torch.empty_like(latent)
            L.ni_latent_softcodebook_fwd(ptr(latent), None, ptr(cb), ptr(tmp), ptr(hist), n, cb.numel(), float(self.v), float(self.gamma), st)
        ent = zeros((1,))
        L.ni_entropy_from_hist(ptr(hist), n, cb.numel(), 0.0, ptr(ent), None, st)
        return latent, wrap(ent.reshape(()))

    call = __call__
