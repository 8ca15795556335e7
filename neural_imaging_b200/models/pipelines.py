"""Camera ISP models on the B200 path — API mirror of reference models/pipelines.py (NIPModel :27-166, UNet :169-230,
ONet :353-362). Explicit forward / backward kernel sequences; skip connections are concat-free (producers write
straight into channel slices of the concat buffer), Conv2DTranspose is a 1x1 conv with a depth_to_space epilogue.
"""
import inspect
import os
import sys

import numpy as np
import torch

from .. import _lib, nn, ops
from .._lib import ACT_CLIP01, MODE_BLOCK2, MODE_PLAIN
from ..helpers import paramspec, utils
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
        elif loss_metric in ('SSIM', 'MS-SSIM'):
            raise NotImplementedError('SSIM losses (tf.image.ssim, reference helpers/tf_helpers.py:39-44) are not on the B200 path yet')
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

    def training_step(self, batch_x, batch_y, learning_rate=None):
        """One optimisation step on (raw, rgb target); returns the loss (reference models/pipelines.py:77-90)."""
        x, t = self._prep(batch_x), self._prep(batch_y)
        y = self._forward(x, save=True)
        kind = 0 if self.loss_metric == 'L2' else 1
        L = _lib.lib()
        acc = self._ws.get('loss_acc', (1,))
        L.ni_fill(ptr(acc), 0.0, 1, stream())
        L.ni_image_loss(ptr(y), ptr(t), ptr(acc), y.numel(), kind, stream())
        dy = self._ws.get('dY', y.shape)
        L.ni_image_loss_grad(ptr(y), ptr(t), ptr(dy), y.numel(), kind, 1.0, 0, stream())
        self._backward(dy)
        if learning_rate is not None:
            self.optimizer.lr = float(learning_rate)
        self.optimizer.apply([self._store])
        return wrap((acc / float(y.numel())).reshape(()))

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
        L, ws, s, S = _lib.lib(), self._ws, stream(), self._h.n_steps
        acts, descs, (B, h, w) = self._saved
        # final conv (input dc{S-1}2)
        last = acts['dc%d2' % (S - 1)]
        dcur = ws.get('d_dc%d2' % (S - 1), last.shape)
        self._final.bprop(last, None, dy, dcur, descs['final'])
        dcats = {}
        for n in reversed(range(1, S)):
            up, c1, c2 = self._dec[n - 1]
            c = c1.cout
            d1, d2, du = descs['dc%d1' % n], descs['dc%d2' % n], descs['dct%d' % n]
            da1 = ws.get('d_dc%d1' % n, acts['dc%d1' % n].shape)
            c2.bprop(acts['dc%d1' % n], acts['dc%d2' % n], dcur, da1, d2)
            dcat = ws.get('d_cat%d' % n, acts['cat%d' % n].shape)
            c1.bprop(acts['cat%d' % n], acts['dc%d1' % n], da1, dcat, d1)
            dcats[n] = dcat
            # transposed conv: dy = first half of dcat seen through depth_to_space addressing
            src = acts['dc%d2' % (n - 1)] if n > 1 else acts['dc02']
            dsrc = ws.get('d_dc%d2' % (n - 1), src.shape)
            up.bprop(src, None, dcat, dsrc, du, dy_addr=(2 * c, 0, MODE_BLOCK2))
            dcur = dsrc
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
                L.ni_maxpool2_bwd(ptr(cat), ptr(dcur), ptr(dcat), ptr(dec2), B, ch, cw, c, 1, 2 * c, c, c, 0, 2 * c, c, c, 0, s)
                y2 = cat
            else:
                dec2, y2 = dcur, acts['dc02']
            da1 = ws.get('d_ec%d1' % n, a1.shape)
            c2.bprop(a1, y2, dec2, da1, d2, dy_addr=(c, 0, MODE_PLAIN))
            src = acts['ep%d' % (n - 1)]
            if n > 1:
                dsrc = ws.get('d_ep%d' % (n - 1), src.shape)
                c1.bprop(src, a1, da1, dsrc, d1)
                dcur = dsrc
                ch, cw = ch * 2, cw * 2
            else:
                c1.bprop(src, a1, da1, None, d1, need_dx=False)
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
