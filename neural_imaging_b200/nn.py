"""Layer engine for the B200 path: flat parameter storage, convolution layers with explicit forward / backward
(no autograd tape: the reference's tf.GradientTape chain is replaced by hand-ordered kernel launches), Keras-Adam.

Layout conventions (TensorFlow's): activations NHWC, Conv2D kernels HWIO (kh, kw, cin, cout), Dense kernels
(in, out) stored as (1, 1, in, out). Conv2DTranspose(2x2, stride 2) is stored as the equivalent 1x1 convolution
(1, 1, cin, 4*cout) whose output goes through depth_to_space(2): W1x1[ci, (a*2+b)*cout + f] = K_keras[a, b, f, ci].
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import (ACT_CLIP01, ACT_LEAKY_RELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, MODE_BLOCK2, MODE_PLAIN, PAD_REFLECT,
                   PAD_SYMMETRIC, PAD_ZERO, ConvDesc)
from .tensor import device, empty, ptr, stream, zeros

ACTIVATIONS = {None: ACT_NONE, 'none': ACT_NONE, 'leaky_relu': ACT_LEAKY_RELU, 'relu': ACT_RELU, 'tanh': ACT_TANH,
               'sigmoid': ACT_SIGMOID, 'clip01': ACT_CLIP01}


def same_padding(size, k, stride):
    """TF SAME rule: (out, pad_before)."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return out, total // 2


def glorot_uniform(rng, shape):
    """Keras default kernel initializer; fans follow keras.initializers._compute_fans."""
    receptive = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    fan_in, fan_out = shape[-2] * receptive, shape[-1] * receptive
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def variance_scaling(rng, shape, scale=1.0):
    """tf.keras.initializers.VarianceScaling() defaults (fan_in, truncated normal at 2 sigma; models/pipelines.py:313):
    stddev = sqrt(scale / fan_in) / 0.87962566103423978 (the truncation's variance correction)."""
    receptive = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    std = math.sqrt(scale / max(1.0, shape[-2] * receptive)) / 0.87962566103423978
    out = rng.normal(size=shape)
    bad = np.abs(out) > 2.0
    while bad.any():                       # resample the tails (truncated normal)
        out[bad] = rng.normal(size=int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * std).astype(np.float32)


# bench.py's CPU-baseline leg sets this to build the models' initial weights without a GPU (specs only, no device buffers)
HOST_ONLY = False

# The filter gradient of a layer runs on a SIDE stream next to the same layer's input gradient (both only read dy; they write disjoint
# buffers) and is joined before bprop returns, so nothing outside the pair can observe the concurrency. The two persistent kernels cannot
# share an SM (shared memory), so the later one takes the SMs the earlier one leaves: the tail of one fills with the head of the other and
# grids smaller than the machine (deep layers, small per-GPU batches) run side by side. Off while bench.py's per-entry event profiler is
# installed (events on the main stream would mis-attribute the time). The fork / join are stream waits, so they capture into CUDA graphs.
WGRAD_OVERLAP = os.environ.get('NI_WGRAD_OVERLAP', '1') != '0'
_SIDE_STREAMS = {}


# Deferred join (`with deferred_wgrad_join():`, used by the U-Net and FAN backward passes; NI_WGRAD_DEFER=0 turns it off): inside the block bprop does not join the side stream; the filter gradients queue
# up on it and run beside EVERYTHING the main stream does until the block ends. Only for backward passes whose gradient and activation
# buffers are all distinct (U-Net, FAN: one workspace buffer per layer), because a filter gradient still reads its dy and x then.
WGRAD_DEFER = os.environ.get('NI_WGRAD_DEFER', '1') != '0'
_DEFER = {'depth': 0, 'pending': None}


class deferred_wgrad_join(object):
    def __enter__(self):
        if WGRAD_DEFER:
            _DEFER['depth'] += 1
        return self

    def __exit__(self, *exc):
        if WGRAD_DEFER:
            _DEFER['depth'] -= 1
            if _DEFER['depth'] == 0 and _DEFER['pending'] is not None:
                torch.cuda.current_stream().wait_stream(_DEFER['pending'])
                _DEFER['pending'] = None
        return False


def _side_stream():
    if not WGRAD_OVERLAP or _lib.PROFILER is not None:
        return None
    dev = torch.cuda.current_device()
    s = _SIDE_STREAMS.get(dev)
    if s is None:
        s = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return s


class Param:
    __slots__ = ('name', 'shape', 'init', 'trainable', 'offset', 'size', 'value', 'grad', 'keras')

    def __init__(self, name, shape, init, trainable):
        self.name, self.shape, self.trainable = name, tuple(int(s) for s in shape), trainable
        # storage rule of the Keras variable this parameter stands for (Keras .h5 interchange, models/tfmodel.py): None = same shape,
        # 'dense' = (in, out), 'conv2d_transpose' = (2, 2, cout, cin), 'internal' = no Keras counterpart (not part of a weight file)
        self.keras = None
        self.init = np.ascontiguousarray(np.broadcast_to(np.asarray(init, dtype=np.float32), self.shape))
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.offset, self.value, self.grad = None, None, None


class ParamStore:
    """All parameters of a model in ONE flat device buffer (+ one flat gradient buffer): Adam and the data-parallel
    all-reduce are each a single launch over the flat buffers. Offsets are padded to 4 floats (16-byte alignment)."""

    def __init__(self):
        self.params = []
        self.flat = self.gflat = self.frozen = self.arena = None

    def add(self, name, shape, init, trainable=True):
        p = Param(name, shape, init, trainable)
        self.params.append(p)
        return p

    def finalize(self):
        if HOST_ONLY:
            return self
        for trainable in (True, False):
            ps = [p for p in self.params if p.trainable == trainable]
            off = 0
            for p in ps:
                p.offset = off
                off += (p.size + 3) // 4 * 4
            host = np.zeros((max(off, 4),), np.float32)
            for p in ps:
                host[p.offset:p.offset + p.size] = p.init.reshape(-1)
            buf = torch.from_numpy(host).to(device())
            if trainable:
                self.flat, self.gflat = buf, torch.zeros_like(buf)
            else:
                self.frozen = buf
            for p in ps:
                p.value = buf[p.offset:p.offset + p.size].view(p.shape if p.shape else ())
                p.grad = self.gflat[p.offset:p.offset + p.size].view(p.shape if p.shape else ()) if trainable else None
        return self

    @property
    def trainable(self):
        return [p for p in self.params if p.trainable]

    def rebind_gradients(self, gflat):
        """Move the gradient buffer into `gflat` (a slice of a shared arena, same length): every parameter's .grad view follows."""
        assert gflat.numel() == self.flat.numel() and gflat.dtype == self.flat.dtype
        self.gflat = gflat
        for p in self.trainable:
            p.grad = gflat[p.offset:p.offset + p.size].view(p.shape if p.shape else ())

    def count(self):
        return int(sum(p.size for p in self.trainable))

    def state_dict(self):
        return {p.name: p.value.detach().cpu().numpy().copy() for p in self.params}

    def load_state_dict(self, state, strict=True):
        for p in self.params:
            if p.name not in state:
                if strict:
                    raise KeyError('missing parameter {}'.format(p.name))
                continue
            a = np.asarray(state[p.name], dtype=np.float32)
            if tuple(a.shape) != p.shape:
                raise ValueError('shape mismatch for {}: {} vs {}'.format(p.name, a.shape, p.shape))
            if p.value is None:                        # HOST_ONLY stores (no device buffers): the host copy is the state
                p.init = np.ascontiguousarray(a).copy()
                continue
            p.value.copy_(torch.from_numpy(np.ascontiguousarray(a)).reshape(p.shape))


class Conv2D:
    """Keras Conv2D / Dense / Conv2DTranspose(2,2) with explicit fprop / bprop through libni_b200.so."""

    def __init__(self, store, name, k, cin, cout, stride=1, padding='SAME', activation=None, use_bias=True, rng=None,
                 kernel_init=None, bias_init=None, trainable=True, pad_mode=PAD_ZERO, alpha=0.2, bias_mod=0, explicit_pad=None,
                 keras=None):
        self.name, self.k, self.cin, self.cout, self.stride = name, int(k), int(cin), int(cout), int(stride)
        self.padding, self.act, self.alpha = padding, ACTIVATIONS[activation], float(alpha)
        self.pad_mode, self.bias_mod, self.explicit_pad = pad_mode, int(bias_mod), explicit_pad
        shape = (self.k, self.k, self.cin, self.cout)
        if kernel_init is None:
            kernel_init = glorot_uniform(rng, shape)
        self.w = store.add(name + '/kernel', shape, kernel_init, trainable)
        self.w.keras = keras if keras is not None else ('conv2d_transpose' if self.bias_mod else None)
        nb = self.bias_mod if self.bias_mod else self.cout
        self.b = store.add(name + '/bias', (nb,), np.zeros((nb,), np.float32) if bias_init is None else bias_init,
                           trainable) if use_bias else None
        if self.b is not None and keras == 'internal':
            self.b.keras = 'internal'
        self._wt = None          # (kh,kw,cout,cin) copy for dgrad
        self._bexp = None        # bias tiled to cout when bias_mod (transposed conv)

    # ---- geometry
    def out_hw(self, h, w):
        if self.explicit_pad is not None:        # mirrored pad p then VALID conv: same spatial size for p = k//2
            p = self.explicit_pad
            return (h + 2 * p - self.k) // self.stride + 1, (w + 2 * p - self.k) // self.stride + 1
        if self.padding == 'SAME':
            return same_padding(h, self.k, self.stride)[0], same_padding(w, self.k, self.stride)[0]
        return (h - self.k) // self.stride + 1, (w - self.k) // self.stride + 1

    def desc(self, n, h, w, in_pitch=None, in_coff=0, out_pitch=None, out_coff=0, in_mode=MODE_PLAIN, out_mode=MODE_PLAIN,
             accumulate=False, act=None):
        oh, ow = self.out_hw(h, w)
        if self.explicit_pad is not None:
            pt = pl = self.explicit_pad
        elif self.padding == 'SAME':
            pt, pl = same_padding(h, self.k, self.stride)[1], same_padding(w, self.k, self.stride)[1]
        else:
            pt = pl = 0
        cin_phys = self.cin // 4 if in_mode == MODE_BLOCK2 else self.cin
        cout_phys = self.cout // 4 if out_mode == MODE_BLOCK2 else self.cout
        d = ConvDesc()
        d.n, d.h, d.w, d.cin, d.cout, d.kh, d.kw, d.stride = n, h, w, self.cin, self.cout, self.k, self.k, self.stride
        d.pad_t, d.pad_l, d.oh, d.ow = pt, pl, oh, ow
        d.in_pitch, d.in_coff = (cin_phys if in_pitch is None else in_pitch), in_coff
        d.out_pitch, d.out_coff = (cout_phys if out_pitch is None else out_pitch), out_coff
        d.in_mode, d.out_mode = in_mode, out_mode
        d.act, d.act_alpha = (self.act if act is None else act), self.alpha
        d.accumulate, d.pad_mode, d.bias_mod = int(accumulate), self.pad_mode, self.bias_mod
        return d

    def _bias_ptr(self):
        return None if self.b is None else ptr(self.b.value)

    def _scratch(self, tag, shape):
        """Per-layer persistent buffer keyed by shape (never re-pointed: captured CUDA graphs keep its address)."""
        bufs = self.__dict__.setdefault('_scratch_bufs', {})
        key = (tag, tuple(int(v) for v in shape))
        if key not in bufs:
            bufs[key] = empty(key[1])
        return bufs[key]

    # ---- compute
    def _s2d_desc(self, d):
        """The 3x3 stride-1 convolution over space_to_depth(2) of the input that equals this 5x5 stride-2 SAME convolution (see StridedConv5)."""
        d3 = ConvDesc()
        d3.n, d3.h, d3.w, d3.cin, d3.cout, d3.kh, d3.kw, d3.stride = d.n, d.h // 2, d.w // 2, 4 * self.cin, self.cout, 3, 3, 1
        d3.pad_t, d3.pad_l, d3.oh, d3.ow = 1, 1, d.h // 2, d.w // 2
        d3.in_pitch, d3.in_coff, d3.in_mode = 4 * self.cin, 0, MODE_PLAIN
        d3.out_pitch, d3.out_coff, d3.out_mode = d.out_pitch, d.out_coff, d.out_mode
        d3.act, d3.act_alpha, d3.accumulate, d3.pad_mode, d3.bias_mod = d.act, d.act_alpha, d.accumulate, PAD_ZERO, d.bias_mod
        return d3

    def _image_end_s2(self, d):
        """5x5 stride-2 SAME convolution on a 3 / 4-channel image with >= 32-multiple outputs (DCN encoder input layer)."""
        return (d.stride == 2 and d.kh == 5 and d.kw == 5 and self.padding == 'SAME' and d.h % 2 == 0 and d.w % 2 == 0 and 2 <= self.cin <= 4
                and self.cout % 32 == 0 and d.in_mode == MODE_PLAIN and d.in_pitch == self.cin and d.in_coff == 0 and d.out_mode == MODE_PLAIN)

    def fprop(self, x, y, d, weight=None):
        L = _lib.lib()
        if weight is None and self._image_end_s2(d):
            # as a 3x3 convolution over space_to_depth(2): 4 * cin = 12 / 16 contraction channels on zero-padded tensor-core tiles instead of
            # the FP32 strided fallback (3.7 -> 1.1 ms for 1280 x 128 x 128 x 3 -> 64)
            st = stream()
            xs = self._scratch('xs', (d.n, d.h // 2, d.w // 2, 4 * self.cin))
            w3 = self._scratch('w3', (3, 3, 4 * self.cin, self.cout))
            L.ni_space_to_depth2(ptr(x), ptr(xs), d.n, d.h // 2, d.w // 2, self.cin, 0, 0, st)
            L.ni_s2conv_weights(ptr(self.w.value), ptr(w3), self.cin, self.cout, 0, st)
            L.ni_conv2d_fprop(ctypes.byref(self._s2d_desc(d)), ptr(xs), ptr(w3), self._bias_ptr(), ptr(y), st)
            return y
        L.ni_conv2d_fprop(ctypes.byref(d), ptr(x), ptr(self.w.value if weight is None else weight), self._bias_ptr(), ptr(y), stream())
        return y

    def bias_grad_ptr(self, need_dw=True):
        return ptr(self.b.grad) if (self.b is not None and self.b.trainable and need_dw) else None

    def fuse_info(self, y, pitch=None, coff=0, need_dw=True):
        """What the layer ABOVE needs to fold this layer's activation backward + bias gradient into its dgrad epilogue
        (`bprop(fuse_prev=...)`): forward output y (plain NHWC, channel pitch / offset), activation, bias-gradient pointer."""
        return (y, self.cout if pitch is None else int(pitch), int(coff), self.act, self.alpha, self.bias_grad_ptr(need_dw), self.bias_mod)

    def bprop(self, x, y, dy, dx, d, **kw):
        """Backward of fprop(x -> y): see _bprop. The filter gradient is forked onto the side stream (when an input gradient follows it)
        and joined here, before the caller can touch any of the buffers."""
        self._forked = None
        try:
            return self._bprop(x, y, dy, dx, d, **kw)
        finally:
            if self._forked is not None:
                if _DEFER['depth'] > 0:
                    _DEFER['pending'] = self._forked
                else:
                    torch.cuda.current_stream().wait_stream(self._forked)
                self._forked = None

    def _wgrad(self, dd, x, dy, dw, overlap):
        """ni_conv2d_wgrad on the side stream (behind everything queued on the main stream so far) when `overlap`, else in order."""
        L = _lib.lib()
        side = _side_stream() if (overlap or _DEFER['depth'] > 0) else None
        if side is None:
            L.ni_conv2d_wgrad(ctypes.byref(dd), ptr(x), ptr(dy), ptr(dw), stream())
            return
        side.wait_stream(torch.cuda.current_stream())
        L.ni_conv2d_wgrad(ctypes.byref(dd), ptr(x), ptr(dy), ptr(dw), side.cuda_stream)
        self._forked = side

    def _bprop(self, x, y, dy, dx, d, weight=None, dweight=None, need_dx=True, dy_addr=None, dx_addr=None,
               dx_accumulate=False, need_dw=True, dpad=None, act_bias_done=False, fuse_prev=None):
        """Backward of fprop(x -> y) described by the forward descriptor `d`.

        For mirrored padding (REFLECT / SYMMETRIC + VALID conv) the input gradient is computed on the padded domain into
        `dpad` (n, h+2p, w+2p, cin) and folded back into the plain tensor dx (the transpose of tf.pad).

        dy is modified in place (multiplied by the activation derivative). dW / db go to the flat gradient buffer
        (or `dweight`). dy_addr / dx_addr = (pitch, coff, mode) of the gradient buffers when they are laid out
        differently from y / x (default: same addressing as the forward tensors). act_bias_done: dy already carries the
        activation derivative and the bias gradient has been written (ni_maxpool2_act_bwd_bias, or the fused dgrad of the layer above).
        fuse_prev = producer.fuse_info(...): when the tcgen05 path takes this dgrad, its epilogue multiplies dx by the producing layer's
        act'(y) and accumulates that layer's bias gradient; `self.fused_prev` then tells the caller to pass act_bias_done=True below."""
        self.fused_prev = False
        L = _lib.lib()
        st = stream()
        dyp, dyo, dym = dy_addr if dy_addr is not None else (d.out_pitch, d.out_coff, d.out_mode)
        db = ptr(self.b.grad) if (self.b is not None and self.b.trainable and need_dw) else None
        f4 = d.cout // 4
        if (dym == MODE_BLOCK2 and d.kh * d.kw > 1 and d.stride == 1 and d.cin % 32 == 0 and d.cout % 4 == 0 and dyp == f4
                and (d.cout % 32 == 0 or 8 <= d.cout < 32) and (f4 % 4 == 0 or d.act in (ACT_NONE, ACT_CLIP01))
                and dyo == 0 and d.out_mode == MODE_BLOCK2 and d.out_pitch == f4 and d.out_coff == 0 and self.bias_mod == 0):
            # k x k convolution with a depth_to_space(2) epilogue (the sub-pixel up-sampling layers of the DCN decoder, the U-Net output
            # layer 32 -> 12 whose input gradient runs on zero-padded tensor-core tiles): the tensor-map
            # path reads depth_to_space-addressed gradients for 1x1 filters only, and the scalar fallback ran these layers at 14 - 30 TFLOP/s.
            # Activation backward on the physical (n, 2h, 2w, F) layout (elementwise, y and dy share it), then ONE space_to_depth copy of dy
            # into the logical (n, h, w, 4F) layout -- block-major, the order depth_to_space wrote -- and everything else on the dense path.
            if not act_bias_done and d.act not in (ACT_NONE, ACT_CLIP01):
                L.ni_act_bwd_bias(ptr(y), ptr(dy), None, d.n, 2 * d.oh, 2 * d.ow, f4, f4, 0, MODE_PLAIN, f4, 0, MODE_PLAIN, d.act, d.act_alpha, 0, st)
            deep = self._scratch('dy_deep', (d.n, d.oh, d.ow, d.cout))
            L.ni_space_to_depth2(ptr(dy), ptr(deep), d.n, d.oh, d.ow, f4, 0, 0, st)
            if not act_bias_done and db is not None:
                L.ni_act_bwd_bias(None, ptr(deep), db, d.n, d.oh, d.ow, d.cout, d.cout, 0, MODE_PLAIN, d.cout, 0, MODE_PLAIN, ACT_NONE, 0.0, 0, st)
            dy, dyp, dyo, dym, act_bias_done = deep, d.cout, 0, MODE_PLAIN, True
        if not act_bias_done and (db is not None or d.act not in (ACT_NONE, ACT_CLIP01)):
            L.ni_act_bwd_bias(ptr(y), ptr(dy), db, d.n, d.oh, d.ow, d.cout, d.out_pitch, d.out_coff, d.out_mode,
                              dyp, dyo, dym, d.act, d.act_alpha, self.bias_mod, st)
        dd = ConvDesc()
        ctypes.memmove(ctypes.byref(dd), ctypes.byref(d), ctypes.sizeof(ConvDesc))
        dd.out_pitch, dd.out_coff, dd.out_mode = dyp, dyo, dym
        dd.accumulate = 0
        if need_dw and (self.w.trainable or dweight is not None):
            self._wgrad(dd, x, dy, self.w.grad if dweight is None else dweight, overlap=need_dx)
        if need_dx:
            wv = self.w.value if weight is None else weight
            if d.pad_mode != PAD_ZERO:
                pad = d.pad_t
                if dpad is None or tuple(dpad.shape) != (d.n, d.h + 2 * pad, d.w + 2 * pad, self.cin):
                    raise ValueError('mirrored padding: bprop needs a dpad buffer of shape (n, h+2p, w+2p, cin)')
                if dx_addr is not None:
                    raise ValueError('mirrored padding: dx must be a plain tensor')
                dd.h, dd.w, dd.pad_t, dd.pad_l, dd.pad_mode = d.h + 2 * pad, d.w + 2 * pad, 0, 0, PAD_ZERO
                dd.in_pitch, dd.in_coff, dd.in_mode, dd.accumulate = self.cin, 0, MODE_PLAIN, 0
                L.ni_conv2d_dgrad(ctypes.byref(dd), ptr(dy), ptr(wv), ptr(dpad), st)
                L.ni_pad_fold(ptr(dpad), ptr(dx), d.n, d.h, d.w, self.cin, pad, d.pad_mode, int(dx_accumulate), st)
                return dx
            if (d.stride == 2 and d.kh == 5 and d.kw == 5 and self.padding == 'SAME' and d.h % 2 == 0 and d.w % 2 == 0 and self.cin <= 4
                    and dx_addr is None and d.in_mode == MODE_PLAIN and d.in_pitch == self.cin and d.in_coff == 0 and dym == MODE_PLAIN
                    and weight is None):
                # Input gradient of the 5x5 stride-2 image-end convolution (DCN encoder, 3 -> 64; needed when the gradient flows on into the
                # ISP): computed in the space_to_depth(2) domain, where it is the input gradient of a 3x3 stride-1 convolution over 4 * cin
                # channels (see StridedConv5), then scattered back. The generic strided fallback took 126 ms for 1280 x 128 x 128.
                w3 = self._scratch('w3', (3, 3, 4 * self.cin, self.cout))
                L.ni_s2conv_weights(ptr(wv), ptr(w3), self.cin, self.cout, 0, st)
                d3 = ConvDesc()
                d3.n, d3.h, d3.w, d3.cin, d3.cout, d3.kh, d3.kw, d3.stride = d.n, d.h // 2, d.w // 2, 4 * self.cin, self.cout, 3, 3, 1
                d3.pad_t, d3.pad_l, d3.oh, d3.ow = 1, 1, d.h // 2, d.w // 2
                d3.in_pitch, d3.in_coff, d3.in_mode = 4 * self.cin, 0, MODE_PLAIN
                d3.out_pitch, d3.out_coff, d3.out_mode = dyp, dyo, MODE_PLAIN
                d3.act, d3.act_alpha, d3.accumulate, d3.pad_mode, d3.bias_mod = ACT_NONE, 0.0, 0, PAD_ZERO, 0
                dxs = self._scratch('dxs', (d.n, d.h // 2, d.w // 2, 4 * self.cin))
                L.ni_conv2d_dgrad(ctypes.byref(d3), ptr(dy), ptr(w3), ptr(dxs), st)
                L.ni_space_to_depth2(ptr(dx), ptr(dxs), d.n, d.h // 2, d.w // 2, self.cin, 1, int(dx_accumulate), st)
                return dx
            if dx_addr is not None:
                dd.in_pitch, dd.in_coff, dd.in_mode = dx_addr
            dd.accumulate = int(dx_accumulate)
            dd.pad_mode = PAD_ZERO
            if fuse_prev is not None and not dx_accumulate:
                yp, ypitch, ycoff, actp, alphap, dbp, bmodp = fuse_prev
                if actp in (ACT_NONE, ACT_LEAKY_RELU, ACT_RELU, ACT_TANH, ACT_SIGMOID) and \
                        L.ni_conv2d_dgrad_act_supported(ctypes.byref(dd), ypitch, ycoff):
                    L.ni_conv2d_dgrad_act_tc(ctypes.byref(dd), ptr(dy), ptr(wv), ptr(dx), ptr(yp), ypitch, ycoff, actp, alphap, dbp, bmodp, st)
                    self.fused_prev = True
                    return dx
            L.ni_conv2d_dgrad(ctypes.byref(dd), ptr(dy), ptr(wv), ptr(dx), st)
        return dx


class StridedConv5(Conv2D):
    """Conv2D(k = 5, stride 2, SAME) on even-sized inputs, evaluated as a 3x3 stride-1 convolution over space_to_depth(2) of the input
    (see ni_space_to_depth2 / ni_s2conv_weights in csrc/nn_misc.cu): the parameters keep Keras' (5, 5, cin, cout) layout, the
    computation runs on the tcgen05 implicit-GEMM path (cin % 8 == 0, cout % 32 == 0) instead of the FP32 SIMT fallback."""

    def __init__(self, store, name, cin, cout, **kw):
        super().__init__(store, name, 5, cin, cout, stride=2, **kw)
        if cin % 8 or cout % 32:
            raise ValueError('StridedConv5 needs cin % 8 == 0 and cout % 32 == 0')
        self._inner = None           # geometry helper: a 3x3 stride-1 convolution over 4 * cin channels
        self._bufs = {}

    def _buf(self, tag, shape):
        b = self._bufs.get((tag, tuple(shape)))
        if b is None:
            b = self._bufs[(tag, tuple(shape))] = empty(tuple(shape))
        return b

    def desc(self, n, h, w, **kw):
        if h % 2 or w % 2:
            raise ValueError('StridedConv5 needs even input sides, got {}x{}'.format(h, w))
        d = ConvDesc()
        d.n, d.h, d.w, d.cin, d.cout, d.kh, d.kw, d.stride = n, h // 2, w // 2, 4 * self.cin, self.cout, 3, 3, 1
        d.pad_t, d.pad_l, d.oh, d.ow = 1, 1, h // 2, w // 2
        d.in_pitch, d.in_coff, d.out_pitch, d.out_coff = 4 * self.cin, 0, self.cout, 0
        d.in_mode, d.out_mode = MODE_PLAIN, MODE_PLAIN
        d.act, d.act_alpha = (self.act if kw.get('act') is None else kw['act']), self.alpha
        d.accumulate, d.pad_mode, d.bias_mod = int(kw.get('accumulate', False)), PAD_ZERO, 0
        return d

    def _w3(self):
        L = _lib.lib()
        w3 = self._buf('w3', (3, 3, 4 * self.cin, self.cout))
        L.ni_s2conv_weights(ptr(self.w.value), ptr(w3), self.cin, self.cout, 0, stream())
        return w3

    def fprop(self, x, y, d, weight=None):
        L = _lib.lib()
        xs = self._buf('xs', (d.n, d.h, d.w, 4 * self.cin))
        L.ni_space_to_depth2(ptr(x), ptr(xs), d.n, d.h, d.w, self.cin, 0, 0, stream())
        L.ni_conv2d_fprop(ctypes.byref(d), ptr(xs), ptr(self._w3()), self._bias_ptr(), ptr(y), stream())
        return y

    def bprop(self, x, y, dy, dx, d, need_dx=True, need_dw=True, act_bias_done=False, dx_accumulate=False, **unused):
        """x: the layer's plain input (its space_to_depth copy from the forward pass is still in the layer's buffer)."""
        L, st = _lib.lib(), stream()
        self.fused_prev = False
        db = ptr(self.b.grad) if (self.b is not None and self.b.trainable and need_dw) else None
        if not act_bias_done and (db is not None or d.act not in (ACT_NONE, ACT_CLIP01)):
            L.ni_act_bwd_bias(ptr(y), ptr(dy), db, d.n, d.oh, d.ow, d.cout, d.out_pitch, d.out_coff, d.out_mode,
                              d.out_pitch, d.out_coff, d.out_mode, d.act, d.act_alpha, 0, st)
        dd = ConvDesc()
        ctypes.memmove(ctypes.byref(dd), ctypes.byref(d), ctypes.sizeof(ConvDesc))
        dd.accumulate = 0
        xs = self._buf('xs', (d.n, d.h, d.w, 4 * self.cin))
        if need_dw and self.w.trainable:
            dw3 = self._buf('dw3', (3, 3, 4 * self.cin, self.cout))
            L.ni_conv2d_wgrad(ctypes.byref(dd), ptr(xs), ptr(dy), ptr(dw3), st)
            L.ni_s2conv_weights(ptr(self.w.grad), ptr(dw3), self.cin, self.cout, 1, st)
        if need_dx:
            dxs = self._buf('dxs', (d.n, d.h, d.w, 4 * self.cin))
            L.ni_conv2d_dgrad(ctypes.byref(dd), ptr(dy), ptr(self._w3()), ptr(dxs), st)
            L.ni_space_to_depth2(ptr(dx), ptr(dxs), d.n, d.h, d.w, self.cin, 1, int(dx_accumulate), st)
        return dx


def unify_gradients(stores):
    """Lay the flat gradient buffers of several stores out back to back in ONE arena, so that the data-parallel exchange is a single
    all-reduce over one bucket instead of one collective per model. Returns the arena. Call before the first step (nothing
    may have captured the old gradient pointers yet)."""
    total = sum(s.gflat.numel() for s in stores)
    arena = torch.zeros((total,), dtype=torch.float32, device=stores[0].gflat.device)
    off = 0
    for s in stores:
        n = s.gflat.numel()
        s.rebind_gradients(arena[off:off + n])
        s.arena = arena
        off += n
    return arena


class AdamKeras:
    """tf.keras.optimizers.Adam semantics (eps=1e-7 outside the bias correction) over flat parameter buffers
    (reference workflows/manipulation_classification.py:156,279-283)."""

    def __init__(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
        self.lr, self.beta1, self.beta2, self.eps = float(lr), beta1, beta2, eps
        self.iterations = 0
        self._state = {}
        self._flag = None

    def step_size(self, iterations=None):
        """Bias-corrected step size of iteration t (Keras: lr * sqrt(1 - beta2^t) / (1 - beta1^t))."""
        t = self.iterations if iterations is None else iterations
        return self.lr * np.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1 ** t)

    def apply(self, stores, gscale=1.0, check_finite=True, lr_t_dev=None):
        """One update of every flat buffer. lr_t_dev: device scalar holding step_size() (CUDA-graph capture; the caller
        advances `iterations` and refreshes the scalar before each replay), else the step size is a kernel argument."""
        L = _lib.lib()
        if lr_t_dev is None:
            self.iterations += 1
        if self._flag is None:
            self._flag = zeros((1,), torch.int32)
        for s in stores:
            key = s.flat.data_ptr()
            if key not in self._state:
                self._state[key] = (torch.zeros_like(s.flat), torch.zeros_like(s.flat))
            m, v = self._state[key]
            if lr_t_dev is None:
                L.ni_adam_keras(ptr(s.flat), ptr(s.gflat), ptr(m), ptr(v), s.flat.numel(), self.lr, self.beta1, self.beta2, self.eps,
                                self.iterations, gscale, ptr(self._flag), stream())
            else:
                L.ni_adam_keras_dev(ptr(s.flat), ptr(s.gflat), ptr(m), ptr(v), s.flat.numel(), ptr(lr_t_dev), self.beta1, self.beta2,
                                    self.eps, gscale, ptr(self._flag), stream())
        return self._flag

    def nonfinite(self):
        """Device->host read of the NaN/Inf flag raised by the fused Adam kernel (lazy check)."""
        if self._flag is None:
            return False
        bad = bool(self._flag.item())
        if bad:
            self._flag.zero_()
        return bad
