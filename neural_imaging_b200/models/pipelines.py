"""Camera ISP models on the B200 path — API mirror of reference models/pipelines.py (NIPModel :27-166, UNet :169-230,
INet :233-295, DNet :298-350, ONet :353-362, ClassicISP :415-539 with models/layers.py:206-258). Explicit forward / backward kernel sequences; skip connections are concat-free (producers write
straight into channel slices of the concat buffer), Conv2DTranspose is a 1x1 conv with a depth_to_space epilogue.
"""
import inspect
import os
import sys

import numpy as np
import torch

from .. import _lib, nn, ops
from .._lib import ACT_CLIP01, MODE_BLOCK2, MODE_PLAIN, PAD_REFLECT
from ..helpers import kernels, paramspec, utils
from ..tensor import Workspace, as_device, empty, ptr, stream, wrap, zeros
from .tfmodel import Placeholder, TFModel

_ACTIVATIONS = {'leaky_relu', 'relu', 'tanh', 'sigmoid'}


class NIPModel(TFModel):
    """Abstract neural imaging pipeline. Sub-classes implement construct_model(), _forward(x, save) -> y and
    _backward(dy) (parameter gradients into the flat gradient buffer)."""

    def __init__(self, loss_metric='L2', patch_size=None, in_channels=4, seed=None, **kwargs):
        super().__init__()
        self.x = Placeholder((patch_size, patch_size, in_channels))
        self.in_channels = in_channels
        self.patch_size = patch_size
        self._rng = np.random.RandomState(seed)
        self._store = nn.ParamStore()
        self._ws = Workspace()
        self.construct_model(**kwargs)
        self._store.finalize()
        self._has_attributes(['y'])
        self.loss_metric = loss_metric
        self.construct_loss(loss_metric)
        self.optimizer = nn.AdamKeras()

    def construct_loss(self, loss_metric):
        from ..helpers import tf_helpers
        if loss_metric == 'L2':
            self.loss = tf_helpers.mse
        elif loss_metric == 'L1':
            self.loss = tf_helpers.mae
        elif loss_metric == 'SSIM':
            self.loss = tf_helpers.ssim_loss
        elif loss_metric == 'MS-SSIM':
            self.loss = tf_helpers.msssim_loss
        else:
            raise ValueError('Unsupported loss metric!')

    def construct_model(self):
        raise NotImplementedError()

    def _prep(self, batch_x):
        x = as_device(batch_x)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        return x

    def process(self, batch_x, training=False):
        """Develop RAW input and return the RGB image (device tensor with .numpy())."""
        y = self._forward(self._prep(batch_x), save=False)
        return wrap(y.clone())

    def training_step(self, batch_x, batch_y, learning_rate=None, grad_sync=None):
        """One optimisation step on (raw, rgb target); returns the loss (reference models/pipelines.py:77-90).
        grad_sync (parallel.GradSync): batch-sharded data parallelism — gradients are all-reduced between backward and Adam; the
        loss is a mean, so the 1 / world_size average rides in the Adam kernel's gradient scale."""
        x, t = self._prep(batch_x), self._prep(batch_y)
        y = self._forward(x, save=True)
        L = _lib.lib()
        acc = self._ws.get('loss_acc', (1,))
        L.ni_fill(ptr(acc), 0.0, 1, stream())
        dy = self._ws.get('dY', y.shape)
        self.loss_forward(y, t, acc, 1.0)
        self.loss_backward(y, t, dy, 1.0)
        self._backward(dy)
        if grad_sync is not None:
            grad_sync([self._store])
        if learning_rate is not None:
            self.optimizer.lr = float(learning_rate)
        self.optimizer.apply([self._store], getattr(grad_sync, 'gscale', 1.0))
        return wrap((acc / float(y.numel())).reshape(()))

    def loss_forward(self, y, t, acc, grad_scale=1.0):
        """acc += numel(y) * loss(y, t) on the device (callers divide by numel, the L2 / L1 convention of ni_image_loss).
        grad_scale is the factor loss_backward will be asked for (the structural losses fix it in the forward pass)."""
        if self.loss_metric in ('SSIM', 'MS-SSIM'):
            if getattr(self, '_sloss', None) is None:
                self._sloss = ops.StructuralLoss(self.loss_metric == 'MS-SSIM', self._ws)
            self._sloss.forward(y, t, acc, loss_scale=float(y.numel()), grad_scale=float(grad_scale))
        else:
            _lib.lib().ni_image_loss(ptr(y), ptr(t), ptr(acc), y.numel(), 0 if self.loss_metric == 'L2' else 1, stream())

    def loss_backward(self, y, t, dy, scale):
        """dy = scale * d loss(y, t) / dy (after loss_forward on the same tensors with grad_scale = scale)."""
        if self.loss_metric in ('SSIM', 'MS-SSIM'):
            self._sloss.backward(dy, accumulate=False)
        else:
            _lib.lib().ni_image_loss_grad(ptr(y), ptr(t), ptr(dy), y.numel(), 0 if self.loss_metric == 'L2' else 1, float(scale), 0, stream())

    def reset_performance_stats(self):
        self.performance = {'loss': {'training': [], 'validation': []}, 'psnr': {'validation': []}, 'ssim': {'validation': []}}

    def get_hyperparameters(self):
        p = {'in_channels': self.in_channels}
        if hasattr(self, '_h'):
            p.update(self._h.to_json())
        return p

    @property
    def patch_size_raw(self):
        return self.x.shape[1:]

    @property
    def patch_size_rgb(self):
        return self.y.shape[1:] if hasattr(self.y, 'shape') else None

    def summary(self):
        return '{:s} : {} -> {}'.format(super().summary(), utils.format_patch_shape(self.patch_size_raw),
                                        utils.format_patch_shape(self.patch_size_rgb))

    def load_model(self, dirname):
        if '/' not in dirname:
            dirname = os.path.join('data/models/nip', dirname)
        super().load_model(dirname)

    def save_model(self, dirname, epoch=0, quiet=False):
        if '/' not in dirname:
            dirname = os.path.join('data/models/nip', dirname)
        super().save_model(dirname, epoch=epoch, quiet=quiet)


class UNet(NIPModel):
    """5-level U-Net (reference models/pipelines.py:169-230): per level 2 x [3x3 SAME conv + act], 2x2 SAME max-pool;
    decoder: 2x2/s2 transposed conv, concat [upsampled, skip], 2 x conv; 3x3 conv -> 12, depth_to_space(2), STE clip."""

    def construct_model(self, **kwargs):
        self._h = paramspec.ParamSpec({
            'n_steps': (5, int, (2, 6)),
            'activation': ('leaky_relu', str, _ACTIVATIONS),
        })
        self._h.update(**kwargs)
        act, st, rng, S = self._h.activation, self._store, self._rng, self._h.n_steps
        self._enc, self._dec = [], []
        cin = self.in_channels
        for n in range(1, S + 1):
            c = 32 * 2 ** (n - 1)
            self._enc.append((nn.Conv2D(st, 'ec{}1'.format(n), 3, cin, c, activation=act, rng=rng),
                              nn.Conv2D(st, 'ec{}2'.format(n), 3, c, c, activation=act, rng=rng)))
            cin = c
        for n in range(1, S):
            c = 32 * 2 ** (S - n - 1)
            # Conv2DTranspose(c, 2x2, stride 2) == 1x1 conv (2c -> 4c) + depth_to_space(2); one bias per real feature
            up = nn.Conv2D(st, 'dct{}'.format(n), 1, 2 * c, 4 * c, padding='VALID', rng=None, bias_mod=c,
                           kernel_init=self._transposed_init(rng, 2 * c, c))
            self._dec.append((up, nn.Conv2D(st, 'dc{}1'.format(n), 3, 2 * c, c, activation=act, rng=rng),
                              nn.Conv2D(st, 'dc{}2'.format(n), 3, c, c, activation=act, rng=rng)))
        self._final = nn.Conv2D(st, 'dc{}'.format(S), 3, 32, 12, activation='clip01', rng=rng)
        p = self.patch_size
        self.y = Placeholder((None if p is None else 2 * p, None if p is None else 2 * p, 3))
        self._saved = None

    @staticmethod
    def _transposed_init(rng, cin, cout):
        """Glorot-uniform draw in the Keras Conv2DTranspose layout (2,2,cout,cin), re-packed as the 1x1 conv kernel."""
        k = nn.glorot_uniform(rng, (2, 2, cout, cin))
        return np.ascontiguousarray(k.transpose(3, 0, 1, 2).reshape(1, 1, cin, 4 * cout))

    @property
    def model_code(self):
        return '{}_{}'.format(self.class_name, self._h.n_steps)

    def _forward(self, x, save=False):
        L, ws, s, S = _lib.lib(), self._ws, stream(), self._h.n_steps
        B, h, w = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        if (h % (2 ** (S - 1))) or (w % (2 ** (S - 1))):
            raise ValueError('UNet input size must be a multiple of {}'.format(2 ** (S - 1)))
        acts, descs = {'ep0': x}, {}
        cur, ch, cw = x, h, w
        for n in range(1, S + 1):
            c1, c2 = self._enc[n - 1]
            c = c1.cout
            d1 = c1.desc(B, ch, cw)
            a1 = c1.fprop(cur, ws.get('ec%d1' % n, (B, ch, cw, c)), d1)
            if n < S:
                # second conv writes the skip connection straight into the 2nd half of decoder level (S-n)'s concat buffer
                cat = ws.get('cat%d' % (S - n), (B, ch, cw, 2 * c))
                d2 = c2.desc(B, ch, cw, out_pitch=2 * c, out_coff=c)
                c2.fprop(a1, cat, d2)
                ep = ws.get('ep%d' % n, (B, ch // 2, cw // 2, c))
                L.ni_maxpool2_fwd(ptr(cat), ptr(ep), B, ch, cw, c, 1, 2 * c, c, c, 0, s)
                acts['ep%d' % n] = ep
                cur, ch, cw = ep, ch // 2, cw // 2
            else:
                d2 = c2.desc(B, ch, cw)
                cur = c2.fprop(a1, ws.get('ec%d2' % n, (B, ch, cw, c)), d2)
                acts['dc02'] = cur
            acts['ec%d1' % n] = a1
            descs['ec%d1' % n], descs['ec%d2' % n] = d1, d2
        for n in range(1, S):
            up, c1, c2 = self._dec[n - 1]
            c = c1.cout
            cat = ws.get('cat%d' % n, (B, 2 * ch, 2 * cw, 2 * c))
            du = up.desc(B, ch, cw, out_pitch=2 * c, out_coff=0, out_mode=MODE_BLOCK2)
            up.fprop(cur, cat, du)
            ch, cw = 2 * ch, 2 * cw
            d1 = c1.desc(B, ch, cw)
            a1 = c1.fprop(cat, ws.get('dc%d1' % n, (B, ch, cw, c)), d1)
            d2 = c2.desc(B, ch, cw)
            cur = c2.fprop(a1, ws.get('dc%d2' % n, (B, ch, cw, c)), d2)
            acts['cat%d' % n], acts['dc%d1' % n], acts['dc%d2' % n] = cat, a1, cur
            descs['dct%d' % n], descs['dc%d1' % n], descs['dc%d2' % n] = du, d1, d2
        df = self._final.desc(B, ch, cw, out_pitch=3, out_mode=MODE_BLOCK2)
        y = self._final.fprop(cur, ws.get('y', (B, 2 * ch, 2 * cw, 3)), df)
        descs['final'] = df
        if save:
            self._saved = (acts, descs, (B, h, w))
        return y

    def _backward(self, dy):
        """dy: gradient w.r.t. the output (B,2h,2w,3); the STE clip passes it through. Parameter grads -> flat buffer."""
        with nn.deferred_wgrad_join():          # every gradient / activation buffer below is a distinct workspace buffer
            return self._backward_impl(dy)

    def _backward_impl(self, dy):
        L, ws, s, S = _lib.lib(), self._ws, stream(), self._h.n_steps
        acts, descs, (B, h, w) = self._saved
        # final conv (input dc{S-1}2)
        last = acts['dc%d2' % (S - 1)]
        dcur = ws.get('d_dc%d2' % (S - 1), last.shape)
        self._final.bprop(last, None, dy, dcur, descs['final'], fuse_prev=self._dec[S - 2][2].fuse_info(last) if S > 1 else None)
        dcats = {}
        done = self._final.fused_prev      # the gradient in `dcur` already carries act'(y) and the bias gradient of its layer has been accumulated
        for n in reversed(range(1, S)):
            up, c1, c2 = self._dec[n - 1]
            c = c1.cout
            d1, d2, du = descs['dc%d1' % n], descs['dc%d2' % n], descs['dct%d' % n]
            da1 = ws.get('d_dc%d1' % n, acts['dc%d1' % n].shape)
            # the dgrad of each convolution also applies the activation backward + bias gradient of the layer below it (tcgen05 epilogue)
            c2.bprop(acts['dc%d1' % n], acts['dc%d2' % n], dcur, da1, d2, act_bias_done=done, fuse_prev=c1.fuse_info(acts['dc%d1' % n]))
            dcat = ws.get('d_cat%d' % n, acts['cat%d' % n].shape)
            c1.bprop(acts['cat%d' % n], acts['dc%d1' % n], da1, dcat, d1, act_bias_done=c2.fused_prev)
            dcats[n] = dcat
            # transposed conv: dy = first half of dcat seen through depth_to_space addressing
            src = acts['dc%d2' % (n - 1)] if n > 1 else acts['dc02']
            below = self._dec[n - 2][2] if n > 1 else self._enc[S - 1][1]      # the layer that produced `src`
            dsrc = ws.get('d_dc%d2' % (n - 1), src.shape)
            up.bprop(src, None, dcat, dsrc, du, dy_addr=(2 * c, 0, MODE_BLOCK2), fuse_prev=below.fuse_info(src))
            dcur, done = dsrc, up.fused_prev
        # encoder
        ch, cw = h // 2 ** (S - 1), w // 2 ** (S - 1)
        for n in reversed(range(1, S + 1)):
            c1, c2 = self._enc[n - 1]
            c = c1.cout
            d1, d2 = descs['ec%d1' % n], descs['ec%d2' % n]
            a1 = acts['ec%d1' % n]
            if n < S:
                # d(ec_n2) = max-pool backward of d(ep_n) + skip-connection gradient (2nd half of d_cat)
                m = S - n
                cat, dcat = acts['cat%d' % m], dcats[m]
                dec2 = ws.get('d_ec%d2' % n, (B, ch, cw, c))
                # + activation backward and bias gradient of ec_n2 (its output is the second half of `cat`) in the same pass
                L.ni_maxpool2_act_bwd_bias(ptr(cat), ptr(dcur), ptr(dcat), ptr(dec2), c2.bias_grad_ptr(), B, ch, cw, c, 1, 2 * c, c, c, 0,
                                           2 * c, c, c, 0, d2.act, d2.act_alpha, s)
                y2, fused = cat, True
            else:
                dec2, y2, fused = dcur, acts['dc02'], done
            da1 = ws.get('d_ec%d1' % n, a1.shape)
            c2.bprop(a1, y2, dec2, da1, d2, dy_addr=(c, 0, MODE_PLAIN), act_bias_done=fused, fuse_prev=c1.fuse_info(a1))
            src = acts['ep%d' % (n - 1)]
            if n > 1:
                dsrc = ws.get('d_ep%d' % (n - 1), src.shape)
                c1.bprop(src, a1, da1, dsrc, d1, act_bias_done=c2.fused_prev)
                dcur = dsrc
                ch, cw = ch * 2, cw * 2
            else:
                c1.bprop(src, a1, da1, None, d1, need_dx=False, act_bias_done=c2.fused_prev)
        return None


class INet(NIPModel):
    """Neural pipeline replicating a standard ISP (reference models/pipelines.py:233-295): 1x1 CFA up-sampling 4 -> 12 +
    depth_to_space, REFLECT pad + k x k demosaicing (bilinear init), 1x1 colour matrix, gamma MLP 3 -> 12 (tanh) -> 3, STE clip."""

    def construct_model(self, random_init=False, kernel=5, trainable_upsampling=False, cfa_pattern='gbrg'):
        self._h = paramspec.ParamSpec({
            'random_init': (False, bool, None),
            'kernel': (5, int, (3, 11)),
            'trainable_upsampling': (False, bool, None),
            'cfa_pattern': ('gbrg', str, {'gbrg', 'rggb', 'bggr'})
        })
        self._h.update(random_init=random_init, kernel=kernel, trainable_upsampling=trainable_upsampling, cfa_pattern=cfa_pattern)
        k, rng, st = self._h.kernel, self._rng, self._store
        upk = kernels.upsampling_kernel(self._h.cfa_pattern)
        if self._h.random_init:
            dmf = rng.normal(0, 0.1, (k, k, 3, 3))
            d1k, d1b, d2k, d2b = rng.normal(0, 0.1, (3, 12)), np.zeros((12,)), rng.normal(0, 0.1, (12, 3)), np.zeros((3,))
            srgbk = np.eye(3)
        else:
            dmf = kernels.bilin_kernel(k)
            d1k, d1b, d2k, d2b = kernels.gamma_kernels()
            srgbk = np.array([[1.82691061, -0.65497452, -0.17193617],
                              [-0.00683982, 1.33216381, -0.32532394],
                              [0.06269717, -0.40055895, 1.33786178]]).transpose()      # models/pipelines.py:268-270
        f32 = lambda a, shape: np.asarray(a, np.float32).reshape(shape)
        pad = (k - 1) // 2
        self._up = nn.Conv2D(st, 'upsampling', 1, 4, 12, use_bias=False, kernel_init=f32(upk, (1, 1, 4, 12)),
                             trainable=self._h.trainable_upsampling)
        self._dm = nn.Conv2D(st, 'demosaicing', k, 3, 3, padding='VALID', use_bias=False, pad_mode=PAD_REFLECT, explicit_pad=pad,
                             kernel_init=f32(dmf, (k, k, 3, 3)))
        self._srgb = nn.Conv2D(st, 'srgb', 1, 3, 3, use_bias=False, kernel_init=f32(srgbk, (1, 1, 3, 3)))
        self._g1 = nn.Conv2D(st, 'gamma_d1', 1, 3, 12, activation='tanh', kernel_init=f32(d1k, (1, 1, 3, 12)), bias_init=f32(d1b, (12,)))
        self._g2 = nn.Conv2D(st, 'gamma_d2', 1, 12, 3, activation='clip01', kernel_init=f32(d2k, (1, 1, 12, 3)), bias_init=f32(d2b, (3,)))
        p = self.patch_size
        self.y = Placeholder((None if p is None else 2 * p, None if p is None else 2 * p, 3))
        self._saved = None

    @property
    def model_code(self):
        return '{c}_{cfa}{tu}{r}_{k}x{k}'.format(c=self.class_name, cfa=self._h.cfa_pattern, k=self._h.kernel,
                                                 tu='T' if self._h.trainable_upsampling else '', r='R' if self._h.random_init else '')

    def _forward(self, x, save=False):
        ws = self._ws
        B, h, w = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        d_up = self._up.desc(B, h, w, out_mode=MODE_BLOCK2)
        bayer = self._up.fprop(x, ws.get('bayer', (B, 2 * h, 2 * w, 3)), d_up)
        d_dm = self._dm.desc(B, 2 * h, 2 * w)
        rgb = self._dm.fprop(bayer, ws.get('rgb', (B, 2 * h, 2 * w, 3)), d_dm)
        d_s = self._srgb.desc(B, 2 * h, 2 * w)
        srgb = self._srgb.fprop(rgb, ws.get('srgb', (B, 2 * h, 2 * w, 3)), d_s)
        d_1 = self._g1.desc(B, 2 * h, 2 * w)
        g0 = self._g1.fprop(srgb, ws.get('g0', (B, 2 * h, 2 * w, 12)), d_1)
        d_2 = self._g2.desc(B, 2 * h, 2 * w)
        y = self._g2.fprop(g0, ws.get('y', (B, 2 * h, 2 * w, 3)), d_2)
        if save:
            self._saved = (x, bayer, rgb, srgb, g0, d_up, d_dm, d_s, d_1, d_2)
        return y

    def _backward(self, dy):
        ws = self._ws
        x, bayer, rgb, srgb, g0, d_up, d_dm, d_s, d_1, d_2 = self._saved
        dg0 = self._g2.bprop(g0, None, dy, ws.get('d_g0', g0.shape), d_2)
        dsrgb = self._g1.bprop(srgb, g0, dg0, ws.get('d_srgb', srgb.shape), d_1)
        drgb = self._srgb.bprop(rgb, None, dsrgb, ws.get('d_rgb', rgb.shape), d_s)
        if self._h.trainable_upsampling:
            pad = d_dm.pad_t
            dpad = ws.get('d_bayer_pad', (d_dm.n, d_dm.h + 2 * pad, d_dm.w + 2 * pad, 3))
            dbayer = self._dm.bprop(bayer, None, drgb, ws.get('d_bayer', bayer.shape), d_dm, dpad=dpad)
            self._up.bprop(x, None, dbayer, None, d_up, need_dx=False)
        else:
            self._dm.bprop(bayer, None, drgb, None, d_dm, need_dx=False)
        return None


class DNet(NIPModel):
    """Joint demosaicing-&-denoising pipeline after Gharbi et al. 2016 (reference models/pipelines.py:298-350):
    n_layers x [k x k VALID conv + ReLU, REFLECT pad] on the RAW stack (the last one with 12 features), depth_to_space,
    concat with the up-sampled Bayer planes, k x k VALID conv + ReLU, REFLECT pad, 1x1 conv -> RGB, STE clip.
    The `conv -> pad` pairs are evaluated as `pad -> conv` of the NEXT layer (mirrored padding folded into its addressing);
    the one pad that feeds depth_to_space is an identity 1x1 convolution with mirrored padding and a d2s epilogue."""

    def construct_model(self, n_layers=15, kernel=3, n_features=64):
        self._h = paramspec.ParamSpec({
            'n_layers': (15, int, (1, 32)),
            'kernel': (3, int, (3, 11)),
            'n_features': (64, int, (4, 128)),
        })
        self._h.update(n_layers=n_layers, kernel=kernel, n_features=n_features)
        k, nf, nl, st, rng = self._h.kernel, self._h.n_features, self._h.n_layers, self._store, self._rng
        pad = (k - 1) // 2
        vs = lambda cin, cout, kk=k: nn.variance_scaling(rng, (kk, kk, cin, cout))
        self._deep = []
        cin = self.in_channels
        for r in range(nl):
            cout = 12 if r == nl - 1 else nf
            kw = dict(padding='VALID') if r == 0 else dict(padding='VALID', pad_mode=PAD_REFLECT, explicit_pad=pad)
            self._deep.append(nn.Conv2D(st, 'conv2d_%d' % r, k, cin, cout, activation='relu', kernel_init=vs(cin, cout), **kw))
            cin = cout
        self._padid = nn.Conv2D(st, 'reflect_pad_d2s', 1, 12, 12, padding='VALID', use_bias=False, trainable=False, pad_mode=PAD_REFLECT,
                                explicit_pad=pad, kernel_init=np.eye(12, dtype=np.float32).reshape(1, 1, 12, 12), keras='internal')
        upk = kernels.upsampling_kernel()
        self._up = nn.Conv2D(st, 'upsampling', 1, 4, 12, use_bias=False, trainable=False,
                             kernel_init=np.asarray(upk, np.float32).reshape(1, 1, 4, 12))
        self._proj = nn.Conv2D(st, 'conv2d_%d' % nl, k, 6, nf, padding='VALID', activation='relu', kernel_init=vs(6, nf))
        self._final = nn.Conv2D(st, 'conv2d_%d' % (nl + 1), 1, nf, 3, padding='VALID', use_bias=False, activation='clip01',
                                pad_mode=PAD_REFLECT, explicit_pad=pad, kernel_init=np.ones((1, 1, nf, 3), np.float32))
        p = self.patch_size
        self.y = Placeholder((None if p is None else 2 * p, None if p is None else 2 * p, 3))
        self._saved = None

    @property
    def model_code(self):
        return '{c}_{k}x{k}_{l}x{f}f'.format(c=self.class_name, k=self._h.kernel, f=self._h.n_features, l=self._h.n_layers)

    def _forward(self, x, save=False):
        ws = self._ws
        B, h, w = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        acts, descs = [x], []
        cur, ch, cw = x, h, w
        for r, conv in enumerate(self._deep):
            d = conv.desc(B, ch, cw)
            cur = conv.fprop(cur, ws.get('deep%d' % r, (B, d.oh, d.ow, conv.cout)), d)
            ch, cw = d.oh, d.ow
            acts.append(cur)
            descs.append(d)
        cat = ws.get('cat', (B, 2 * h, 2 * w, 6))
        d_pad = self._padid.desc(B, ch, cw, out_pitch=6, out_coff=0, out_mode=MODE_BLOCK2)
        if (d_pad.oh, d_pad.ow) != (h, w):
            raise ValueError('DNet: patch too small for {} VALID layers'.format(len(self._deep)))
        self._padid.fprop(cur, cat, d_pad)
        d_up = self._up.desc(B, h, w, out_pitch=6, out_coff=3, out_mode=MODE_BLOCK2)
        self._up.fprop(x, cat, d_up)
        d_pr = self._proj.desc(B, 2 * h, 2 * w)
        pu = self._proj.fprop(cat, ws.get('pu', (B, d_pr.oh, d_pr.ow, self._proj.cout)), d_pr)
        d_f = self._final.desc(B, d_pr.oh, d_pr.ow)
        y = self._final.fprop(pu, ws.get('y', (B, 2 * h, 2 * w, 3)), d_f)
        if save:
            self._saved = (acts, descs, cat, pu, d_pad, d_pr, d_f)
        return y

    def _backward(self, dy):
        ws = self._ws
        acts, descs, cat, pu, d_pad, d_pr, d_f = self._saved
        padded = lambda name, d, c: ws.get(name, (d.n, d.h + 2 * d.pad_t, d.w + 2 * d.pad_l, c))
        dpu = self._final.bprop(pu, None, dy, ws.get('d_pu', pu.shape), d_f, dpad=padded('d_pu_pad', d_f, self._final.cin))
        dcat = self._proj.bprop(cat, pu, dpu, ws.get('d_cat', cat.shape), d_pr)
        dcur = self._padid.bprop(acts[-1], None, dcat, ws.get('d_deep%d' % (len(self._deep) - 1), acts[-1].shape), d_pad,
                                 dpad=padded('d_deep_pad12', d_pad, 12))
        for r in reversed(range(len(self._deep))):
            conv, d = self._deep[r], descs[r]
            if r > 0:
                dprev = ws.get('d_deep%d' % (r - 1), acts[r].shape)
                conv.bprop(acts[r], acts[r + 1], dcur, dprev, d, dpad=padded('d_deep_pad%d' % conv.cin, d, conv.cin))
                dcur = dprev
            else:
                conv.bprop(acts[0], acts[1], dcur, None, d, need_dx=False)
        return None


class ClassicISP(NIPModel):
    """Classic camera ISP (reference models/pipelines.py:415-539 over _ClassicISP :415-446 and DemosaicingLayer,
    models/layers.py:206-258): CFA up-sampling -> demosaicing (bilinear filter minus alpha x CNN residual, or a plain CNN) ->
    sRGB matrix -> straight-through clip to [1/255, 1] -> gamma 1/2.2."""

    def construct_model(self, srgb_mat=None, kernel=5, c_filters=(), cfa_pattern='gbrg', residual=True, brightness=None):
        self._h = paramspec.ParamSpec({
            'kernel': (5, int, (3, 11)),
            'c_filters': ((), tuple, paramspec.numbers_in_range(int, 1, 1024)),
            'cfa_pattern': ('gbrg', str, {'gbrg', 'rggb', 'bggr'}),
            'residual': (True, bool, None)
        })
        self._h.update(kernel=kernel, c_filters=tuple(c_filters), cfa_pattern=cfa_pattern, residual=residual)
        if brightness is not None:
            raise NotImplementedError('brightness normalisation is unreachable through ClassicISP in the reference '
                                      '(models/pipelines.py:476 drops the argument) and needs a host percentile')
        k, st, rng = self._h.kernel, self._store, self._rng
        up = np.asarray(kernels.upsampling_kernel(self._h.cfa_pattern), np.float32).reshape(1, 1, 4, 12)
        self._up = nn.Conv2D(st, 'upsampling', 1, 4, 12, use_bias=False, trainable=False, kernel_init=up)
        srgb = np.eye(3, dtype=np.float32) if srgb_mat is None else np.asarray(srgb_mat, np.float32)
        self._srgb = nn.Conv2D(st, 'srgb', 1, 3, 3, use_bias=False, trainable=False, kernel_init=srgb.T.reshape(1, 1, 3, 3).copy())
        if self._h.residual:
            self._alpha = st.add('demosaicing/alpha', (), np.float32(0.1), True)
            self._bilinear = nn.Conv2D(st, 'demosaicing/bilinear', k, 3, 3, padding='VALID', use_bias=False, trainable=False,
                                       pad_mode=PAD_REFLECT, explicit_pad=(k - 1) // 2, kernel_init=kernels.bilin_kernel(k))
        else:
            self._alpha = self._bilinear = None
        # the CNN branch is only evaluated (and, in Keras, only built) when it has more than the final 1x1 layer or when it
        # is the whole demosaicing model (models/layers.py:244-254)
        self._cnn = []
        if len(self._h.c_filters) > 0 or not self._h.residual:
            cin = 3
            for i, nf in enumerate(self._h.c_filters):
                self._cnn.append(nn.Conv2D(st, 'demosaicing/conv2d_%d' % i, k, cin, nf, activation='leaky_relu', rng=rng))
                cin = nf
            self._cnn.append(nn.Conv2D(st, 'demosaicing/conv2d_%d' % len(self._h.c_filters), 1, cin, 3,
                                       activation='tanh' if self._h.residual else 'sigmoid', rng=rng))
        p = self.patch_size
        self.y = Placeholder((None if p is None else 2 * p, None if p is None else 2 * p, 3))
        self._saved = None
        self._dalpha = None

    # ---- camera configuration (models/pipelines.py:480-512)
    def set_cfa_pattern(self, cfa_pattern):
        if cfa_pattern is not None:
            cfa_pattern = cfa_pattern.lower()
            up = np.asarray(kernels.upsampling_kernel(cfa_pattern), np.float32).reshape(1, 1, 4, 12)
            self._up.w.value.copy_(torch.from_numpy(up))
            self._h.update(cfa_pattern=cfa_pattern)

    def set_srgb_conversion(self, srgb_mat):
        if srgb_mat is not None:
            srgb = np.ascontiguousarray(np.asarray(srgb_mat, np.float32).T.reshape(1, 1, 3, 3))
            self._srgb.w.value.copy_(torch.from_numpy(srgb))

    def set_camera(self, camera, config='config/cameras.json'):
        import json
        with open(config) as f:
            cameras = json.load(f)
        self.set_cfa_pattern(cameras[camera]['cfa'])
        self.set_srgb_conversion(np.array(cameras[camera]['srgb']))

    def process(self, batch_x, training=False, cfa_pattern=None, srgb_mat=None):
        self.set_cfa_pattern(cfa_pattern)
        self.set_srgb_conversion(srgb_mat)
        return super().process(batch_x, training)

    @property
    def model_code(self):
        return 'ClassicISP_{cfa}_{k}x{k}_{fs}-{of}{r}'.format(fs='-'.join(['{:d}'.format(x) for x in self._h.c_filters]), of=3,
                                                             k=self._h.kernel, cfa=self._h.cfa_pattern, r='R' if self._h.residual else '')

    def summary(self):
        nf = len(self._h.c_filters)
        fs = self._h.c_filters[0] if len(set(self._h.c_filters)) == 1 else '*'
        k = self._h.kernel
        return '{}[{}] + CNN demosaicing [{}+1 layers : {}x{}x{} -> 1x1x3]'.format(self.class_name, self._h.cfa_pattern, nf, k, k, fs)

    def summary_compact(self):
        nf = len(self._h.c_filters)
        fs = self._h.c_filters[0] if len(set(self._h.c_filters)) == 1 else '*'
        k = self._h.kernel
        return '{}[{}, {}+1 conv2D {}x{}x{} > 1x1x3]'.format(self.class_name, self._h.cfa_pattern, nf, k, k, fs)

    def _forward(self, x, save=False):
        L, ws, s = _lib.lib(), self._ws, stream()
        B, h, w = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        H, W = 2 * h, 2 * w
        d_up = self._up.desc(B, h, w, out_mode=MODE_BLOCK2)
        bayer = self._up.fprop(x, ws.get('bayer', (B, H, W, 3)), d_up)
        feats, descs, cur = [bayer], [], bayer
        for i, conv in enumerate(self._cnn):
            d = conv.desc(B, H, W)
            cur = conv.fprop(cur, ws.get('cnn%d' % i, (B, H, W, conv.cout)), d)
            feats.append(cur)
            descs.append(d)
        dm = ws.get('demosaiced', (B, H, W, 3))
        if self._bilinear is not None:
            d_b = self._bilinear.desc(B, H, W)
            xb = self._bilinear.fprop(bayer, ws.get('bilinear', (B, H, W, 3)), d_b)
            if self._cnn:
                L.ni_residual_alpha_fwd(ptr(xb), ptr(cur), ptr(self._alpha.value), ptr(dm), xb.numel(), 1, s)
            else:
                L.ni_affine(ptr(xb), ptr(dm), 1.0, 0.0, 1, xb.numel(), s)          # f = 0: y = clip(x_bilinear)
        else:
            L.ni_affine(ptr(cur), ptr(dm), 1.0, 0.0, 1, cur.numel(), s)
        d_s = self._srgb.desc(B, H, W)
        rgb = self._srgb.fprop(dm, ws.get('rgb', (B, H, W, 3)), d_s)
        y = ws.get('y', (B, H, W, 3))
        L.ni_gamma_clip_fwd(ptr(rgb), ptr(y), rgb.numel(), 1.0 / 255, 1.0, 1.0 / 2.2, s)
        if save:
            self._saved = (feats, descs, rgb, dm, d_s)
        return y

    def _backward(self, dy):
        L, ws, s = _lib.lib(), self._ws, stream()
        feats, descs, rgb, dm, d_s = self._saved
        if not self._cnn:
            return None                              # only alpha is trainable and it multiplies f = 0
        drgb = ws.get('d_rgb', rgb.shape)
        L.ni_gamma_clip_bwd(ptr(rgb), ptr(dy), ptr(drgb), rgb.numel(), 1.0 / 255, 1.0, 1.0 / 2.2, s)
        ddm = self._srgb.bprop(dm, None, drgb, ws.get('d_dm', dm.shape), d_s, need_dw=False)
        if self._bilinear is not None:
            df = ws.get('d_f', feats[-1].shape)
            if self._dalpha is None:
                self._dalpha = zeros((1,))
            self._dalpha.zero_()
            L.ni_residual_alpha_bwd(ptr(ddm), ptr(feats[-1]), ptr(self._alpha.value), ptr(df), ptr(self._dalpha), ddm.numel(), s)
            self._alpha.grad.copy_(self._dalpha[0])
            dcur = df
        else:
            dcur = ddm
        for i in reversed(range(len(self._cnn))):
            conv = self._cnn[i]
            if i > 0:
                dprev = ws.get('d_cnn%d' % (i - 1), feats[i].shape)
                conv.bprop(feats[i], feats[i + 1], dcur, dprev, descs[i])
                dcur = dprev
            else:
                conv.bprop(feats[0], feats[1], dcur, None, descs[0], need_dx=False)
        return None


class ONet(NIPModel):
    """Dummy pipeline for RGB training: identity on (2p, 2p, 3) (reference models/pipelines.py:353-362)."""

    def construct_model(self):
        p = self.x.shape[1]
        self.x = Placeholder((None if p is None else 2 * p, None if p is None else 2 * p, 3))
        self.y = Placeholder(self.x.shape[1:])

    def _forward(self, x, save=False):
        return x

    def _backward(self, dy):
        return None


supported_models = [name for name, obj in inspect.getmembers(sys.modules[__name__])
                    if type(obj) is type and issubclass(obj, NIPModel) and name != 'NIPModel']
