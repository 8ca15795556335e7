"""Forensic analysis network (FAN) on the B200 path — API mirror of reference models/forensics.py:12-133.

Graph (models/forensics.py:61-92): ConstrainedConv2D (models/layers.py:36-57) -> n x [Conv2D kxk SAME + act +
MaxPool2D 2x2 VALID] -> Conv2D 1x1 + act -> GAP | Flatten -> n_dense x Dense(act) -> Dense(n_classes, softmax).
Forward and backward are explicit kernel sequences (no tape); all parameters live in one flat buffer.
"""
import ctypes

import numpy as np
import torch

from .. import _lib, nn
from .._lib import ACT_NONE, MODE_PLAIN, PAD_SYMMETRIC, PAD_ZERO
from ..helpers import kernels, paramspec
from ..tensor import Workspace, as_device, empty, ptr, stream, wrap, zeros
from .tfmodel import Placeholder, TFModel

_ACTIVATIONS = {'leaky_relu', 'relu', 'tanh', 'sigmoid'}


class FAN(TFModel):

    def __init__(self, n_classes, patch_size=None, n_filters=32, n_fscale=2, n_convolutions=4, kernel=5, dropout=0.0,
                 use_gap=True, n_dense=0, activation='leaky_relu', seed=None):
        super().__init__()
        self._h = paramspec.ParamSpec({
            'n_classes': (7, int, (2, 256)),
            'n_filters': (32, int, (4, 128)),
            'n_fscale': (2, float, (0.25, 4)),
            'n_convolutions': (4, int, (1, 32)),
            'kernel': (5, int, (3, 11)),
            'dropout': (0, float, (0, 1)),
            'use_gap': (False, bool, None),
            'n_dense': (2, int, (0, 16)),
            'activation': ('leaky_relu', str, _ACTIVATIONS),
        })
        self._h.update(n_classes=n_classes, n_filters=n_filters, n_fscale=n_fscale, n_convolutions=n_convolutions,
                       kernel=kernel, dropout=dropout, use_gap=use_gap, n_dense=n_dense, activation=activation)
        self.patch_size = patch_size
        self.n_classes = int(n_classes)
        self.x = Placeholder((patch_size, patch_size, 3))
        self.y = Placeholder((self.n_classes,))
        if not use_gap and patch_size is None:
            raise ValueError('Flatten (use_gap=False) needs a fixed patch_size')

        rng = np.random.RandomState(seed)
        self._seed, self._dropout_calls = int(seed or 0), 0
        self._fuse_pool = True          # tests switch it off to compare with the conv -> max-pool kernel pair
        st = self._store = nn.ParamStore()
        act = self._h.activation
        # constrained residual filter: trainable raw kernel (5,5,3,3), normalised on every call (models/layers.py:36-53)
        f = np.array([[0, 0, 0, 0, 0], [0, -1, -2, -1, 0], [0, -2, 12, -2, 0], [0, -1, -2, -1, 0], [0, 0, 0, 0, 0]])
        self.filter_strength = 100.0
        self._cconv = nn.Conv2D(st, 'constrained_conv2d', 5, 3, 3, padding='VALID', use_bias=False, pad_mode=PAD_SYMMETRIC,
                                explicit_pad=2, kernel_init=kernels.repeat_2dfilter(f, 3))
        self._convs = []
        nf, cin = n_filters, 3
        for i in range(self._h.n_convolutions):
            self._convs.append(nn.Conv2D(st, 'conv2d_{}'.format(i), self._h.kernel, cin, nf, activation=act, rng=rng))
            cin, nf = nf, int(nf * self._h.n_fscale)
        nf = nf // n_fscale
        self._conv1x1 = nn.Conv2D(st, 'conv2d_1x1', 1, cin, int(nf), padding='VALID', activation=act, rng=rng)
        feat = int(nf)
        if not use_gap:
            s = patch_size // (2 ** self._h.n_convolutions)
            feat = s * s * int(nf)
        self._dense = []
        for i in range(self._h.n_dense):
            nf = nf // n_fscale
            self._dense.append(nn.Conv2D(st, 'dense_{}'.format(i), 1, feat, int(nf), padding='VALID', activation=act, rng=rng, keras='dense'))
            feat = int(nf)
        self._out = nn.Conv2D(st, 'dense_out', 1, feat, self.n_classes, padding='VALID', rng=rng, keras='dense')
        st.finalize()
        self._ws = Workspace()
        self._nf = None if nn.HOST_ONLY else empty((5, 5, 3, 3))        # normalised constrained filter
        self._dnf = None if nn.HOST_ONLY else empty((5, 5, 3, 3))
        self._saved = None
        self.optimizer = nn.AdamKeras()

    def reset_performance_stats(self):
        self.performance = {'loss': {'training': [], 'validation': []}, 'accuracy': {'validation': []}, 'confusion': []}

    # ------------------------------------------------------------------------------------------------ forward
    def _forward(self, x, save=False, training=False):
        """x: (M,H,W,3) device tensor. Returns logits (M, n_classes); keeps activations when save=True."""
        L, ws, s = _lib.lib(), self._ws, stream()
        m, h, w = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        L.ni_constrained_filter_fwd(ptr(self._cconv.w.value), ptr(self._nf), 5, 3, self.filter_strength, s)
        acts = {'x': x}
        d0 = self._cconv.desc(m, h, w)
        r = ws.get('r', (m, h, w, 3))
        L.ni_cconv5_fwd(ptr(x), ptr(self._nf), ptr(r), m, h, w, s)
        acts['r'], descs = r, {'cconv': d0}
        cur, ch, cw = r, h, w
        for i, conv in enumerate(self._convs):
            d = conv.desc(m, ch, cw)
            p = ws.get('p%d' % i, (m, ch // 2, cw // 2, conv.cout))
            if self._fuse_pool and L.ni_conv2d_pool2_supported(ctypes.byref(d)):
                # conv + bias + activation + max-pool in one kernel: pooled output + one code byte per pooled element (arg-max position,
                # activation slope) instead of the full-resolution activation (first block: 2.7 GB at 1280 images)
                code = ws.get('code%d' % i, p.shape, torch.uint8)
                L.ni_conv2d_pool2_fwd(ctypes.byref(d), ptr(cur), ptr(conv.w.value), conv._bias_ptr(), ptr(p), ptr(code), s)
                c = None
                acts['code%d' % i] = code
            else:
                c = conv.fprop(cur, ws.get('c%d' % i, (m, ch, cw, conv.cout)), d)
                L.ni_maxpool2_fwd(ptr(c), ptr(p), m, ch, cw, conv.cout, 0, conv.cout, 0, conv.cout, 0, s)
            acts['c%d' % i], acts['p%d' % i], descs['c%d' % i] = c, p, d
            cur, ch, cw = p, ch // 2, cw // 2
        d = self._conv1x1.desc(m, ch, cw)
        f = self._conv1x1.fprop(cur, ws.get('f', (m, ch, cw, self._conv1x1.cout)), d)
        acts['f'], descs['f'] = f, d
        if self._h.use_gap:
            g = ws.get('gap', (m, self._conv1x1.cout))
            L.ni_gap_fwd(ptr(f), ptr(g), m, ch * cw, self._conv1x1.cout, s)
        else:
            g = f.view(m, -1)
        acts['g'] = g
        cur = g
        for i, dl in enumerate(self._dense):
            d = dl.desc(m, 1, 1)
            cur = dl.fprop(cur, ws.get('d%d' % i, (m, dl.cout)), d)
            acts['d%d' % i], descs['d%d' % i] = cur, d
            if training and self._h.dropout > 0:       # Dropout after every hidden dense layer (models/forensics.py:88), inference: identity
                if save:
                    raise NotImplementedError('back-propagation through an active Dropout layer (the reference never trains with it active)')
                dropped = ws.get('drop%d' % i, (m, dl.cout))
                self._dropout_calls += 1
                L.ni_dropout(ptr(cur), ptr(dropped), cur.numel(), float(self._h.dropout), (self._seed << 20) + self._dropout_calls, s)
                cur = dropped
        d = self._out.desc(m, 1, 1)
        logits = self._out.fprop(cur, ws.get('logits', (m, self.n_classes)), d)
        descs['out'] = d
        if save:
            self._saved = (acts, descs, (m, h, w, ch, cw))
        return logits

    def process(self, batch_x, training=False):
        """Class probabilities (M, n_classes) for an image batch (NHWC rgb)."""
        x = as_device(batch_x)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        logits = self._forward(x, training=bool(training))
        probs = empty(logits.shape)
        _lib.lib().ni_softmax_ce(ptr(logits), None, ptr(probs), None, None, logits.shape[0], self.n_classes, 1.0, stream())
        return wrap(probs)

    def process_and_decide(self, batch_x, with_confidence=False):
        probs = self.process(batch_x).numpy()
        if with_confidence:
            return probs.argmax(axis=1), probs.max(axis=1)
        return probs.argmax(axis=1)

    # ------------------------------------------------------------------------------------------------ loss + backward
    def forward_loss(self, x, labels, grad_scale=1.0):
        """Forward with saved activations; returns (probs, loss_sum[1], dlogits). dlogits = d(mean CE)/dlogits * grad_scale."""
        logits = self._forward(x, save=True)
        m = int(logits.shape[0])
        probs, dlogits = self._ws.get('probs', logits.shape), self._ws.get('dlogits', logits.shape)
        loss = self._ws.get('loss', (1,))
        L = _lib.lib()
        L.ni_fill(ptr(loss), 0.0, 1, stream())
        L.ni_softmax_ce(ptr(logits), ptr(labels), ptr(probs), ptr(loss), ptr(dlogits), m, self.n_classes, grad_scale / m, stream())
        return probs, loss, dlogits

    def backward(self, dlogits, need_dx=False):
        """Back-propagate through the saved forward. Parameter gradients -> flat gradient buffer. Returns dx or None."""
        with nn.deferred_wgrad_join():          # every gradient / activation buffer below is a distinct workspace buffer
            return self._backward_impl(dlogits, need_dx)

    def _backward_impl(self, dlogits, need_dx=False):
        L, ws, s = _lib.lib(), self._ws, stream()
        acts, descs, (m, h, w, ch, cw) = self._saved
        cur_in = acts['d%d' % (len(self._dense) - 1)] if self._dense else acts['g']
        feat = self._out.cin
        dcur = ws.get('dfeat_out', (m, feat))
        self._out.bprop(cur_in, None, dlogits, dcur, descs['out'])
        for i in reversed(range(len(self._dense))):
            dl = self._dense[i]
            x_in = acts['d%d' % (i - 1)] if i > 0 else acts['g']
            dprev = ws.get('dfeat_%d' % i, (m, dl.cin))
            dl.bprop(x_in, acts['d%d' % i], dcur, dprev, descs['d%d' % i])
            dcur = dprev
        c1 = self._conv1x1
        if self._h.use_gap:
            df = ws.get('df', (m, ch, cw, c1.cout))
            L.ni_gap_bwd(ptr(dcur), ptr(df), m, ch * cw, c1.cout, s)
        else:
            df = dcur.view(m, ch, cw, c1.cout)
        n_conv = len(self._convs)
        dp = ws.get('dp%d' % (n_conv - 1), (m, ch, cw, c1.cin))
        c1.bprop(acts['p%d' % (n_conv - 1)], acts['f'], df, dp, descs['f'])
        for i in reversed(range(n_conv)):
            conv = self._convs[i]
            d = descs['c%d' % i]
            dc = ws.get('dc%d' % i, (m, d.oh, d.ow, conv.cout))
            # pool backward + activation backward + bias gradient of conv i in one pass
            if acts['c%d' % i] is None:
                L.ni_maxpool2_code_bwd_bias(ptr(acts['code%d' % i]), ptr(dp), ptr(dc), conv.bias_grad_ptr(), m, d.oh // 2, d.ow // 2, conv.cout,
                                            d.act, d.act_alpha, s)
            else:
                L.ni_maxpool2_act_bwd_bias(ptr(acts['c%d' % i]), ptr(dp), None, ptr(dc), conv.bias_grad_ptr(), m, d.oh, d.ow, conv.cout, 0,
                                           conv.cout, 0, conv.cout, 0, 0, 0, conv.cout, 0, d.act, d.act_alpha, s)
            x_in = acts['p%d' % (i - 1)] if i > 0 else acts['r']
            dp = ws.get('dp%d' % (i - 1), (m, d.h, d.w, conv.cin)) if i > 0 else ws.get('dr', (m, h, w, 3))
            conv.bprop(x_in, acts['c%d' % i], dc, dp, d, act_bias_done=True)
        # constrained conv: filter gradient through the normalisation; input gradient through the mirrored pad
        d0 = descs['cconv']
        dx = None
        L.ni_cconv5_bwd_filter(ptr(acts['x']), ptr(dp), ptr(self._dnf), m, h, w, s)
        L.ni_constrained_filter_bwd(ptr(self._cconv.w.value), ptr(self._dnf), ptr(self._cconv.w.grad), 5, 3, self.filter_strength, s)
        if need_dx:
            dx = ws.get('dx', (m, h, w, 3))
            L.ni_cconv5_bwd_data(ptr(dp), ptr(self._nf), ptr(dx), m, h, w, 0, s)       # transpose of (SYMMETRIC pad + VALID conv) in one kernel
        return dx

    def loss(self, labels, probabilities):
        """SparseCategoricalCrossentropy()(labels, probabilities) on probabilities (Keras eager semantics)."""
        p = as_device(probabilities).double().clamp(1e-7, 1 - 1e-7)
        lab = as_device(np.asarray(labels), torch.int64)
        q = p / p.sum(dim=1, keepdim=True)
        return wrap((-torch.log(q.gather(1, lab.view(-1, 1)))).mean().float())

    def training_step(self, batch_x, target_labels, learning_rate=None):
        """One optimisation step on (images, class numbers); returns the loss (reference models/forensics.py:116-125)."""
        # (the reference calls self._model(batch_x) WITHOUT training=True here, so Dropout layers are inactive in its training step too)
        x = as_device(batch_x)
        labels = as_device(np.asarray(target_labels), torch.int32)
        probs, loss, dlogits = self.forward_loss(x, labels)
        self.backward(dlogits, need_dx=False)
        if learning_rate is not None:
            self.optimizer.lr = float(learning_rate)
        self.optimizer.apply([self._store])
        return wrap((loss / x.shape[0]).reshape(()))

    def summary(self):
        return '{kernel}x{kernel} CNN: 1+{conv}+1 conv layers {gap}+ {fc} fc layers [{params:,} parameters]'.format(
            kernel=self._h.kernel, conv=self._h.n_convolutions, fc=self._h.n_dense,
            gap='+ (GAP) ' if self._h.use_gap else '', params=self.count_parameters())
