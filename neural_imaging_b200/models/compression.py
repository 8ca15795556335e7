"""Learned image compression (DCN / TwitterDCN) on the B200 path — API mirror of reference models/compression.py:28-291.

Graph (models/compression.py:213-272), for a (M,H,W,3) batch in [0,1]:
  encoder : 2(x-0.5) -> Conv 5x5/2 (64, act) -> Conv 5x5/2 (128) -> 3 residual blocks [Conv3x3(act), Conv3x3] (the first
            branch starts from leaky_relu(net)) -> Conv 5x5/2 (n_features)
  latent  : DiscreteLatent (models/layers.py:176-203): trainable scalar scale -> soft-codebook quantisation (float64
            t-Student weights, straight-through hard values) -> soft-histogram entropy of the QUANTISED values
  decoder : Conv3x3 (512) -> depth_to_space -> 3 residual blocks -> Conv3x3 (256, act) -> d2s -> Conv3x3 (12) -> d2s ->
            (y+1)/2 with a straight-through clip to [0,1]
  loss    : tf.nn.l2_loss(x - y) + entropy_weight * H          (models/compression.py:89-92)

Forward and backward are explicit kernel sequences over libni_b200.so: depth_to_space is folded into the convolutions'
output addressing (NI_MODE_BLOCK2), residual adds into the second convolution's epilogue (accumulate), and the
quantiser + histogram is one fused kernel (csrc/latent.cu) instead of two (n_values x 32) float64 matrices.
"""
import numpy as np
import torch

from .. import _lib, nn
from .._lib import MODE_BLOCK2
from ..helpers import paramspec
from ..tensor import Workspace, as_device, empty, ptr, stream, wrap, zeros
from .tfmodel import Placeholder, TFModel

_ACTIVATIONS = {'leaky_relu', 'relu', 'tanh', 'sigmoid'}
_NU, _GAMMA = 50.0, 25.0           # DiscreteLatent defaults (models/layers.py:184)


class DCN(TFModel):
    """Abstract learned codec: hyper-parameters, quantiser set-up, loss, optimiser, rate statistics. Child classes
    implement construct_model / _encode / _decode / _backward."""

    def __init__(self, patch_size=128, latent_bpf=5, rounding='soft-codebook', train_codebook=False, entropy_weight=250,
                 scale_latent=True, use_batchnorm=False, loss_metric='L2', seed=None, **kwargs):
        super().__init__()
        self._h = paramspec.ParamSpec({
            'latent_bpf': (5, int, (1, 8)),
            'train_codebook': (False, bool, None),
            'entropy_weight': (250, float, (0, 1e6)),
            'scale_latent': (True, bool, None),
            'use_batchnorm': (False, bool, None),
            'loss_metric': ('L2', str, {'L2'}),
            'rounding': ('soft', str, {'identity', 'soft', 'soft-codebook', 'sin'})
        })
        self._h.update(latent_bpf=latent_bpf, train_codebook=train_codebook, entropy_weight=entropy_weight,
                       scale_latent=scale_latent, use_batchnorm=use_batchnorm, loss_metric=loss_metric, rounding=rounding)
        # latent quantiser of DiscreteLatent(rounding) (models/layers.py:118-170); the entropy estimate is the same for every mode
        self._rounding = {'soft-codebook': 0, 'sin': 1, 'soft': 2, 'identity': 3}[self._h.rounding]
        if self._h.train_codebook:
            raise NotImplementedError('train_codebook=True ("not tested" in the reference, models/compression.py:57) is not implemented')
        self.patch_size = patch_size
        self.x = Placeholder((patch_size, patch_size, 3))
        self._rng = np.random.RandomState(seed)
        self._store = nn.ParamStore()
        qmin, qmax = -2 ** (self._h.latent_bpf - 1) + 1, 2 ** (self._h.latent_bpf - 1)      # models/layers.py:108-109
        self._codebook_host = np.arange(qmin, qmax + 1).astype(np.float32)
        self.construct_model(**kwargs)
        self._has_attributes(['y', 'latent_shape', 'n_latent'])
        self._store.finalize()
        self._ws = Workspace()
        self._saved = None
        self._world = 1
        if not nn.HOST_ONLY:
            self._codebook = torch.from_numpy(self._codebook_host).to(self._store.flat.device)
            self._hist = zeros((len(self._codebook_host),), torch.float64)
            self._gh = zeros((len(self._codebook_host),), torch.float64)
            self._dscale = zeros((1,), torch.float64)
            self._entropy = zeros((1,))
        self.optimizer = nn.AdamKeras()

    def construct_model(self, **params):
        raise NotImplementedError('Not implemented!')

    def reset_performance_stats(self):
        self.performance = self._reset_performance(['loss', 'entropy', 'ssim', 'psnr'])

    # ---- data parallelism: the entropy is a function of the GLOBAL soft histogram => one 32-double all-reduce
    def set_data_parallel(self, world_size):
        self._world = int(world_size)

    # ---- loss (models/compression.py:89-92)
    def loss(self, image_target, image_compressed, entropy):
        a, b = as_device(image_target), as_device(image_compressed)
        acc = zeros((1,))
        _lib.lib().ni_image_loss(ptr(a), ptr(b), ptr(acc), a.numel(), 0, stream())
        ent = entropy if torch.is_tensor(entropy) else torch.tensor(float(entropy), device=acc.device)
        return wrap((acc / (2.0 * 255.0 * 255.0) + self._h.entropy_weight * ent.reshape(-1)[:1].to(acc.dtype)).reshape(()))

    @staticmethod
    def ssim(a, b):
        """mean tf.image.ssim(a, b, max_val=1): 11x11 Gaussian (sigma 1.5), VALID, k1=0.01, k2=0.03 (reporting metric of the training
        step, models/compression.py:89) — one fused kernel (csrc/metrics.cu) instead of five filtered copies of both images."""
        from ..helpers import metrics
        return wrap(metrics.ssim_tf(a, b).mean())

    # ---- public API (models/compression.py:106-139)
    def compress(self, batch_x):
        """(N)HW3 rgb -> quantised latent of the FIRST image (the reference indexes the encoder's output list with [0],
        which selects the latent tensor; kept batch-shaped here like tf does)."""
        x = as_device(batch_x)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        q, _ = self._encode(x, None)
        return wrap(q.clone())

    def decompress(self, batch_z):
        z = as_device(batch_z)
        if z.dim() == 3:
            z = z.unsqueeze(0)
        return wrap(self._decode(z, None).clone())

    def forward(self, x, save=False):
        """x: (M,H,W,3) device tensor -> (y, entropy) device tensors living in the model's workspace."""
        saved = {} if save else None
        q, ent = self._encode(x, saved)
        y = self._decode(q, saved)
        if save:
            self._saved = saved
        return y, ent

    def process(self, batch_x, return_entropy=False):
        x = as_device(batch_x)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        y, ent = self.forward(x)
        y = wrap(y.clone())
        return (y, wrap(ent.clone().reshape(()))) if return_entropy else y

    def training_step(self, batch_x, learning_rate=None, grad_sync=None):
        """One optimisation step; returns {'loss': sqrt(2 loss), 'ssim', 'entropy'} (models/compression.py:123-139).
        grad_sync (parallel.GradSync, after set_data_parallel(world)): the batch is sharded by rank; the soft histogram is all-reduced
        before the entropy is taken (so H is the single-device, batch-global value) and the gradients are summed — tf.nn.l2_loss is a
        SUM over the batch and dH/dz is already normalised by the global count, so no 1 / world_size factor applies."""
        L, s = _lib.lib(), stream()
        x = as_device(batch_x)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        y, ent = self.forward(x, save=True)
        acc = self._ws.get('l2acc', (1,))
        L.ni_fill(ptr(acc), 0.0, 1, s)
        L.ni_image_loss(ptr(x), ptr(y), ptr(acc), x.numel(), 0, s)
        loss = acc / (2.0 * 255.0 * 255.0) + self._h.entropy_weight * ent
        ssim = self.ssim(x, y)
        dy = self._ws.get('dy', y.shape)
        L.ni_image_loss_grad(ptr(y), ptr(x), ptr(dy), y.numel(), 0, y.numel() / (2.0 * 255.0 * 255.0), 0, s)     # dy = y - x
        self.backward(dy, float(self._h.entropy_weight), need_dx=False)
        if grad_sync is not None:
            grad_sync([self._store])
        if learning_rate is not None:
            self.optimizer.lr = float(learning_rate)
        self.optimizer.apply([self._store])
        return {'loss': wrap(torch.sqrt(2 * loss).reshape(())), 'ssim': ssim, 'entropy': wrap(ent.clone().reshape(()))}

    def backward(self, dy, entropy_upstream=0.0, need_dx=False, need_dw=True):
        return self._backward(dy, entropy_upstream, need_dx, need_dw)

    # ---- rate statistics
    def compression_stats(self, patch_size=None, n_latent_bytes=None):
        n_latent_bytes = n_latent_bytes or self._h.latent_bpf / 8
        ps = patch_size or self.patch_size
        if ps is None:
            raise ValueError('Patch size not specified!')
        bitmap_size = ps * ps * 3
        return {
            'rate': bitmap_size / (n_latent_bytes * self.n_latent),
            'bpp': 8 * self.n_latent * n_latent_bytes / (ps * ps),
            'bpf': 8 * n_latent_bytes,
            'bytes': self.n_latent * n_latent_bytes
        }

    def summary(self):
        l_shape = 'x'.join(str(x) for x in self.latent_shape if x is not None)
        return '{} : {}-D latent space @ {}-bpf [{:,.0f} params]'.format(self.class_name, l_shape, self._h.latent_bpf, self.count_parameters())

    def summary_compact(self):
        return '{} {}-D'.format(self.class_name, self.latent_shape[-1])

    @property
    def model_code(self):
        if not hasattr(self, 'n_latent'):
            raise ValueError('The model does not report the latent space dimensionality.')
        return '{}-{}C'.format(type(self).__name__, self._h.n_features)

    def get_codebook(self):
        return self._codebook_host.copy().reshape((-1,))

    # ---- the quantiser (shared by all DCN variants)
    def _quantise(self, z, saved):
        """z (M,h,w,F) -> (q, entropy[1]); histogram / entropy kept for the backward."""
        L, s = _lib.lib(), stream()
        q = self._ws.get('q', z.shape)
        self._hist.zero_()
        scale = ptr(self._scale.value) if self._scale is not None else None
        L.ni_latent_quantise_fwd(ptr(z), scale, ptr(self._codebook), ptr(q), ptr(self._hist), z.numel(), self._codebook.numel(),
                                 _NU, _GAMMA, self._rounding, s)
        n_total = z.numel()
        if self._world > 1:
            torch.distributed.all_reduce(self._hist)
            n_total *= self._world
        L.ni_entropy_from_hist(ptr(self._hist), n_total, self._codebook.numel(), 0.0, ptr(self._entropy), None, s)
        if saved is not None:
            saved['z'], saved['q'], saved['n_total'] = z, q, n_total
        return q, self._entropy

    def _quantise_bwd(self, dq, entropy_upstream, saved, need_dw):
        L, s = _lib.lib(), stream()
        z, q = saved['z'], saved['q']
        dz = self._ws.get('dz', z.shape)
        gh = None
        if entropy_upstream != 0.0:
            L.ni_entropy_from_hist(ptr(self._hist), saved['n_total'], self._codebook.numel(), float(entropy_upstream), None, ptr(self._gh), s)
            gh = ptr(self._gh)
        want_ds = need_dw and self._scale is not None and self._scale.trainable
        if want_ds:
            self._dscale.zero_()
        L.ni_latent_quantise_bwd(ptr(z), ptr(self._scale.value) if self._scale is not None else None, ptr(self._codebook), ptr(q),
                                 ptr(dq), gh, ptr(dz), ptr(self._dscale) if want_ds else None, z.numel(), self._codebook.numel(),
                                 _NU, _GAMMA, self._rounding, s)
        if want_ds:
            self._scale.grad.copy_(self._dscale[0])
        return dz


class TwitterDCN(DCN):
    """Auto-encoder of Theis et al. 2017 as adapted by the toolbox (models/compression.py:184-291)."""

    def construct_model(self, n_features=32, activation='leaky_relu'):
        self._h.add({
            'n_features': (32, int, (4, 128)),
            'activation': ('leaky_relu', str, _ACTIVATIONS)
        })
        self._h.update(n_features=n_features, activation=activation)
        nf = self._h.n_features
        if self.patch_size is None:
            self.latent_shape, self.n_latent = (None, None, nf), None
        else:
            self.latent_shape = (self.patch_size // 8, self.patch_size // 8, nf)
            self.n_latent = int(np.prod(self.latent_shape))
        st, rng, act = self._store, self._rng, self._h.activation
        C = nn.Conv2D
        self._e1 = C(st, 'encoder/conv2d', 5, 3, 64, stride=2, activation=act, rng=rng)
        # the two wide down-sampling layers run as 3x3 convolutions over space_to_depth(2) on the tcgen05 path (nn.StridedConv5)
        self._e2 = nn.StridedConv5(st, 'encoder/conv2d_1', 64, 128, rng=rng)
        self._eres = [(C(st, 'encoder/conv2d_%d' % (2 + 2 * i), 3, 128, 128, activation=act, rng=rng),
                       C(st, 'encoder/conv2d_%d' % (3 + 2 * i), 3, 128, 128, rng=rng)) for i in range(3)]
        self._eout = nn.StridedConv5(st, 'encoder/conv2d_8', 128, nf, rng=rng) if nf % 32 == 0 else C(st, 'encoder/conv2d_8', 5, 128, nf, stride=2, rng=rng)
        self._scale = st.add('encoder/discrete_latent/latent_scaling', (), np.float32(1.0), True) if self._h.scale_latent else None
        self._d1 = C(st, 'decoder/conv2d_9', 3, nf, 512, rng=rng)
        self._dres = [(C(st, 'decoder/conv2d_%d' % (10 + 2 * i), 3, 128, 128, activation=act, rng=rng),
                       C(st, 'decoder/conv2d_%d' % (11 + 2 * i), 3, 128, 128, rng=rng)) for i in range(3)]
        self._d2 = C(st, 'decoder/conv2d_16', 3, 128, 256, activation=act, rng=rng)
        self._d3 = C(st, 'decoder/conv2d_17', 3, 64, 12, rng=rng)
        self.y = Placeholder((self.patch_size, self.patch_size, 3))
        self.latent = Placeholder(self.latent_shape)

    # ------------------------------------------------------------------------------------------------ forward
    def _res_fwd(self, tag, pair, inp, base, m, hh, ww, saved):
        """out = base + conv_b(conv_a(inp)): the add is the second convolution's accumulate epilogue."""
        ca, cb = pair
        ws = self._ws
        da = ca.desc(m, hh, ww)
        r1 = ca.fprop(inp, ws.get(tag + '_r1', (m, hh, ww, ca.cout)), da)
        out = ws.get(tag + '_out', base.shape)
        out.copy_(base)
        db = cb.desc(m, hh, ww, accumulate=True)
        cb.fprop(r1, out, db)
        if saved is not None:
            saved[tag] = (inp, r1, da, db)
        return out

    def _encode(self, x, saved):
        L, ws, s = _lib.lib(), self._ws, stream()
        m, h, w = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        if h % 8 or w % 8:
            raise ValueError('TwitterDCN needs image sides divisible by 8, got {}x{}'.format(h, w))
        x0 = ws.get('x0', x.shape)
        L.ni_affine(ptr(x), ptr(x0), 2.0, -1.0, 0, x.numel(), s)
        d1 = self._e1.desc(m, h, w)
        a1 = self._e1.fprop(x0, ws.get('a1', (m, d1.oh, d1.ow, 64)), d1)
        d2 = self._e2.desc(m, d1.oh, d1.ow)
        net = self._e2.fprop(a1, ws.get('enet0', (m, d2.oh, d2.ow, 128)), d2)
        hh, ww = d2.oh, d2.ow
        nr = ws.get('enet0_act', net.shape)
        L.ni_leaky_relu_fwd(ptr(net), ptr(nr), net.numel(), 0.2, s)              # tf.nn.leaky_relu default alpha
        net0 = net
        for i, pair in enumerate(self._eres):
            net = self._res_fwd('eres%d' % i, pair, nr if i == 0 else net, net, m, hh, ww, saved)
        do = self._eout.desc(m, hh, ww)
        z = self._eout.fprop(net, ws.get('z', (m, do.oh, do.ow, self._eout.cout)), do)
        if saved is not None:
            saved.update(x0=x0, a1=a1, d_e1=d1, d_e2=d2, net0=net0, net3=net, d_eout=do, dims=(m, h, w, hh, ww))
        return self._quantise(z, saved)

    def _decode(self, q, saved):
        L, ws, s = _lib.lib(), self._ws, stream()
        m, lh, lw = int(q.shape[0]), int(q.shape[1]), int(q.shape[2])
        dd1 = self._d1.desc(m, lh, lw, out_mode=MODE_BLOCK2)
        inet = self._d1.fprop(q, ws.get('inet0', (m, 2 * lh, 2 * lw, 128)), dd1)
        hh, ww = 2 * lh, 2 * lw
        for i, pair in enumerate(self._dres):
            inet = self._res_fwd('dres%d' % i, pair, inet, inet, m, hh, ww, saved)
        dd2 = self._d2.desc(m, hh, ww, out_mode=MODE_BLOCK2)
        u2 = self._d2.fprop(inet, ws.get('u2', (m, 2 * hh, 2 * ww, 64)), dd2)
        dd3 = self._d3.desc(m, 2 * hh, 2 * ww, out_mode=MODE_BLOCK2)
        u3 = self._d3.fprop(u2, ws.get('u3', (m, 4 * hh, 4 * ww, 3)), dd3)
        y = ws.get('y', u3.shape)
        L.ni_affine(ptr(u3), ptr(y), 0.5, 0.5, 1, u3.numel(), s)               # (inet+1)/2 then the straight-through clip
        if saved is not None:
            saved.update(qin=q, d_d1=dd1, inet3=inet, d_d2=dd2, u2=u2, d_d3=dd3, u3=u3)
        return y

    # ------------------------------------------------------------------------------------------------ backward
    def _res_bwd(self, tag, pair, g, saved, need_dw, first_encoder_block=False):
        """g = d(out) on entry, d(base) on exit (in place)."""
        ca, cb = pair
        L, ws, s = _lib.lib(), self._ws, stream()
        inp, r1, da, db = saved[tag]
        dr1 = ws.get('dres_r1', r1.shape)
        # conv_a's activation backward + bias gradient ride in the epilogue of conv_b's input-gradient kernel when the tcgen05 path takes it
        cb.bprop(r1, None, g, dr1, db, need_dw=need_dw, fuse_prev=ca.fuse_info(r1, need_dw=need_dw))
        done = cb.fused_prev
        if first_encoder_block:
            dinp = ws.get('dres_in', inp.shape)
            ca.bprop(inp, r1, dr1, dinp, da, need_dw=need_dw, act_bias_done=done)
            L.ni_leaky_relu_bwd(ptr(saved['net0']), ptr(dinp), ptr(g), g.numel(), 0.2, 1, s)
        else:
            ca.bprop(inp, r1, dr1, g, da, dx_accumulate=True, need_dw=need_dw, act_bias_done=done)
        return g

    def _backward(self, dy, entropy_upstream, need_dx, need_dw):
        """dy: d loss / d y (M,H,W,3), overwritten. Parameter gradients -> the flat gradient buffer. Returns dx or None."""
        L, ws, s = _lib.lib(), self._ws, stream()
        sv = self._saved
        m, h, w, hh, ww = sv['dims']
        L.ni_affine(ptr(dy), ptr(dy), 0.5, 0.0, 0, dy.numel(), s)               # through (.+1)/2; the clip is straight-through
        du2 = ws.get('du2', sv['u2'].shape)
        self._d3.bprop(sv['u2'], None, dy, du2, sv['d_d3'], need_dw=need_dw)
        g = ws.get('g_dec', sv['inet3'].shape)
        self._d2.bprop(sv['inet3'], sv['u2'], du2, g, sv['d_d2'], need_dw=need_dw)
        for i in reversed(range(3)):
            self._res_bwd('dres%d' % i, self._dres[i], g, sv, need_dw)
        dq = ws.get('dq', sv['qin'].shape)
        self._d1.bprop(sv['qin'], None, g, dq, sv['d_d1'], need_dw=need_dw)
        dz = self._quantise_bwd(dq, entropy_upstream, sv, need_dw)
        g = ws.get('g_enc', sv['net3'].shape)
        self._eout.bprop(sv['net3'], None, dz, g, sv['d_eout'], need_dw=need_dw)
        for i in reversed(range(3)):
            self._res_bwd('eres%d' % i, self._eres[i], g, sv, need_dw, first_encoder_block=(i == 0))
        da1 = ws.get('da1', sv['a1'].shape)
        self._e2.bprop(sv['a1'], None, g, da1, sv['d_e2'], need_dw=need_dw)
        dx0 = ws.get('dx0', sv['x0'].shape) if need_dx else None
        self._e1.bprop(sv['x0'], sv['a1'], da1, dx0, sv['d_e1'], need_dx=need_dx, need_dw=need_dw)
        if need_dx:
            L.ni_affine(ptr(dx0), ptr(dx0), 2.0, 0.0, 0, dx0.numel(), s)
        return dx0

    @property
    def model_code(self):
        parts = [self._h.rounding, ('Q+{}bpf' if self._h.train_codebook else 'Q-{}bpf').format(self._h.latent_bpf),
                 'S+' if self._h.scale_latent else 'S-']
        if self._h.entropy_weight is not None:
            parts.append('H+{:.2f}'.format(self._h.entropy_weight))
        return '{}/{}'.format(super().model_code, '_'.join(parts))
