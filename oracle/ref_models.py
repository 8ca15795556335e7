"""PyTorch-CPU restatement of the reference models and of the joint training step (TEST INFRASTRUCTURE).

Weights are exchanged as ``{name: ndarray}`` dictionaries in the product's storage layout
(neural_imaging_b200.nn.ParamStore.state_dict(); Conv2D kernels HWIO, transposed convs as (1,1,cin,4*cout)).
"""
from collections import OrderedDict

import numpy as np
import torch

from . import ref_ops as R


def to_params(state, dtype=torch.float32, requires_grad=True):
    return OrderedDict((k, torch.tensor(np.asarray(v), dtype=dtype, requires_grad=requires_grad)) for k, v in state.items())


# ------------------------------------------------------------------------------------------------ UNet
def unet_forward(P, x, n_steps=5, activation='leaky_relu'):
    """models/pipelines.py:190-223."""
    act = R.ACT[activation]
    t = {'ep0': x}
    for n in range(1, n_steps + 1):
        a = act(R.conv2d(t['ep%d' % (n - 1)], P['ec%d1/kernel' % n], P['ec%d1/bias' % n]))
        t['ec%d2' % n] = act(R.conv2d(a, P['ec%d2/kernel' % n], P['ec%d2/bias' % n]))
        if n < n_steps:
            t['ep%d' % n] = R.max_pool(t['ec%d2' % n], same=True)
    cur = t['ec%d2' % n_steps]
    for n in range(1, n_steps):
        up = R.conv2d_transpose_2x2(cur, P['dct%d/kernel' % n], P['dct%d/bias' % n])
        cat = torch.cat((up, t['ec%d2' % (n_steps - n)]), dim=3)                     # [upsampled, skip] (:211)
        a = act(R.conv2d(cat, P['dc%d1/kernel' % n], P['dc%d1/bias' % n]))
        cur = act(R.conv2d(a, P['dc%d2/kernel' % n], P['dc%d2/bias' % n]))
    y = R.conv2d(cur, P['dc%d/kernel' % n_steps], P['dc%d/bias' % n_steps])
    return R.ste_clip(R.depth_to_space(y, 2))


# ------------------------------------------------------------------------------------------------ FAN
def constrained_filter(kernel, strength=100.0):
    """ConstrainedConv2D.call filter normalisation, models/layers.py:45-53."""
    ks, ch = kernel.shape[0], kernel.shape[2]
    ind = torch.zeros_like(kernel)
    for r in range(ch):
        ind[ks // 2, ks // 2, r, r] = 1
    nf = kernel * (1 - ind)
    df = nf.sum(dim=(0, 1, 2)).reshape(1, 1, 1, ch).repeat(ks, ks, ch, 1)
    nf = strength * nf / df
    return nf - strength * ind


def fan_forward(P, x, n_convolutions=4, n_dense=0, use_gap=True, activation='leaky_relu'):
    """models/forensics.py:61-90 -> class probabilities."""
    act = R.ACT[activation]
    nf = constrained_filter(P['constrained_conv2d/kernel'])
    net = R.conv2d(R.tf_pad(x, 2, 'SYMMETRIC'), nf, padding='VALID')
    for i in range(n_convolutions):
        net = act(R.conv2d(net, P['conv2d_%d/kernel' % i], P['conv2d_%d/bias' % i]))
        net = R.max_pool(net, same=False)
    net = act(R.conv2d(net, P['conv2d_1x1/kernel'], P['conv2d_1x1/bias'], padding='VALID'))
    net = net.mean(dim=(1, 2)) if use_gap else net.reshape(net.shape[0], -1)
    for i in range(n_dense):
        net = act(net @ P['dense_%d/kernel' % i][0, 0] + P['dense_%d/bias' % i])
    logits = net @ P['dense_out/kernel'][0, 0] + P['dense_out/bias']
    return torch.softmax(logits, dim=1)


# ------------------------------------------------------------------------------------------------ workflow
MANIPULATIONS = OrderedDict([
    ('sharpen', lambda x, s: R.manipulation_sharpen(x, s, hsv=True)),
    ('resample', lambda x, s: R.manipulation_resample(x, s)),
    ('gaussian', lambda x, s: R.manipulation_gaussian(x, 5, s)),
    ('jpeg', lambda x, s: R.jpeg_manipulation(x, s)),
    ('gamma', lambda x, s: R.manipulation_gamma(x, s)),
    ('median', lambda x, s: R.manipulation_median(x, s)),
])
DEFAULT_STRENGTHS = {'sharpen': 1, 'resample': 50, 'gaussian': 0.83, 'jpeg': 80, 'awgn': 5.1, 'gamma': 3, 'median': 3}


def run_manipulations(Y, names, strengths=None):
    """workflows/manipulation_classification.py:199-208: class-major concat [Y, op1(Y), ...]."""
    strengths = strengths or DEFAULT_STRENGTHS
    return torch.cat([Y] + [MANIPULATIONS[n](Y, strengths[n]) for n in names], dim=0)


def workflow_forward(P_nip, P_fan, x, names=('sharpen', 'resample', 'gaussian', 'jpeg'), quality=50, pool=2, nip='UNet', P_dcn=None):
    """run_workflow (:162-176) for the default distribution channel (pool:k + dJPEG(quality,'soft')), or, with P_dcn, the
    learned codec (compression='dcn'). Returns (Y, c, C, probs[, entropy])."""
    Y = unet_forward(P_nip, x) if nip == 'UNet' else x
    m = run_manipulations(Y, names)
    c = R.avg_pool(m, pool) if pool > 1 else m
    if P_dcn is not None:
        C, ent = twitter_dcn_forward(P_dcn, c)[:2]
        return Y, c, C, fan_forward(P_fan, C), ent
    C = R.djpeg(c, R.jpeg_qtable(quality, 0), R.jpeg_qtable(quality, 1), 'soft')[0] if quality else c
    probs = fan_forward(P_fan, C)
    return Y, c, C, probs


def batch_labels(batch_size, n_classes):
    return np.concatenate([k * np.ones((batch_size,), dtype=np.int64) for k in range(n_classes)])


def training_step(P_nip, P_fan, opt_state, x, y_target, lambda_nip=0.1, lr=1e-4, train_nip=True,
                  names=('sharpen', 'resample', 'gaussian', 'jpeg'), quality=50, pool=2, nip='UNet',
                  P_dcn=None, lambda_dcn=0.0, train_dcn=False, nip_loss=None):
    """ManipulationClassification.training_step (:260-285): loss = ce + lambda_nip * nip_loss (mse unless given); shared Keras Adam.
    opt_state = {'t': int, 'm': {name: tensor}, 'v': {...}}; parameters are updated in place. Returns loss dict + grads."""
    out = workflow_forward(P_nip, P_fan, x, names, quality, pool, nip, P_dcn)
    Y, c, C, probs = out[:4]
    n_classes = len(names) + 1
    loss_ce = R.sparse_categorical_crossentropy(batch_labels(x.shape[0], n_classes), probs)
    loss_nip = (nip_loss or R.mse)(y_target, Y)
    loss = loss_ce + (lambda_nip * loss_nip if train_nip else 0)
    loss_dcn = dcn_loss(c, C, out[4]) if P_dcn is not None else None
    if train_dcn:
        loss = loss + lambda_dcn * loss_dcn
    params = [('fan/' + k, v) for k, v in P_fan.items()]
    if train_nip and nip == 'UNet':
        params += [('nip/' + k, v) for k, v in P_nip.items()]
    if train_dcn:
        params += [('dcn/' + k, v) for k, v in P_dcn.items()]
    grads = torch.autograd.grad(loss, [p for _, p in params], allow_unused=True)
    grads = [torch.zeros_like(p) if g is None else g for g, (_, p) in zip(grads, params)]
    opt_state['t'] += 1
    with torch.no_grad():
        ms = [opt_state['m'].setdefault(k, torch.zeros_like(p)) for k, p in params]
        vs = [opt_state['v'].setdefault(k, torch.zeros_like(p)) for k, p in params]
        R.adam_keras_step([p for _, p in params], grads, ms, vs, opt_state['t'], lr)
    losses = {'loss': float(loss), 'ce': float(loss_ce), 'nip': float(loss_nip)}
    if loss_dcn is not None:
        losses['dcn'] = float(loss_dcn)
    return losses, OrderedDict((k, g) for (k, _), g in zip(params, grads))


# ------------------------------------------------------------------------------------------------ TwitterDCN
def dcn_codebook(latent_bpf=5, dtype=torch.float32):
    """models/layers.py:108-114."""
    return torch.arange(-2 ** (latent_bpf - 1) + 1, 2 ** (latent_bpf - 1) + 1, dtype=dtype)


def twitter_dcn_encode(P, x, latent_bpf=5, rounding='soft-codebook'):
    """models/compression.py:213-241. P uses the product's names (encoder/conv2d[_k], .../latent_scaling)."""
    act = R.ACT['leaky_relu']
    cv = lambda t, name, stride=1: R.conv2d(t, P[name + '/kernel'], P[name + '/bias'], stride=stride)
    net = 2 * (x - 0.5)
    net = act(cv(net, 'encoder/conv2d', 2))
    net = cv(net, 'encoder/conv2d_1', 2)
    for i in range(3):
        inp = R.leaky_relu(net) if i == 0 else net
        net = net + cv(act(cv(inp, 'encoder/conv2d_%d' % (2 + 2 * i))), 'encoder/conv2d_%d' % (3 + 2 * i))
    z = cv(net, 'encoder/conv2d_8', 2)
    return R.discrete_latent(z, P.get('encoder/discrete_latent/latent_scaling'), dcn_codebook(latent_bpf), rounding=rounding) + (z,)


def twitter_dcn_decode(P, q):
    """models/compression.py:247-272."""
    act = R.ACT['leaky_relu']
    cv = lambda t, name: R.conv2d(t, P[name + '/kernel'], P[name + '/bias'])
    inet = R.depth_to_space(cv(q, 'decoder/conv2d_9'), 2)
    for i in range(3):
        inet = inet + cv(act(cv(inet, 'decoder/conv2d_%d' % (10 + 2 * i))), 'decoder/conv2d_%d' % (11 + 2 * i))
    inet = R.depth_to_space(act(cv(inet, 'decoder/conv2d_16')), 2)
    inet = R.depth_to_space(cv(inet, 'decoder/conv2d_17'), 2)
    return R.ste_clip((inet + 1) / 2)


def twitter_dcn_forward(P, x, latent_bpf=5, rounding='soft-codebook'):
    q, ent, z = twitter_dcn_encode(P, x, latent_bpf, rounding)
    return twitter_dcn_decode(P, q), ent, q, z


def dcn_loss(x, y, ent, entropy_weight=250.0):
    """models/compression.py:89-92."""
    return R.l2_loss(x - y) + entropy_weight * ent.to(x.dtype)


def dcn_training_step(P, opt_state, x, lr=1e-3, entropy_weight=250.0, latent_bpf=5, rounding='soft-codebook'):
    """DCN.training_step (models/compression.py:123-139) with Keras Adam; parameters updated in place."""
    y, ent, q, z = twitter_dcn_forward(P, x, latent_bpf, rounding)
    loss = dcn_loss(x, y, ent, entropy_weight)
    names = list(P.keys())
    grads = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
    grads = [torch.zeros_like(P[k]) if g is None else g for g, k in zip(grads, names)]
    opt_state['t'] += 1
    with torch.no_grad():
        ms = [opt_state['m'].setdefault(k, torch.zeros_like(P[k])) for k in names]
        vs = [opt_state['v'].setdefault(k, torch.zeros_like(P[k])) for k in names]
        R.adam_keras_step([P[k] for k in names], grads, ms, vs, opt_state['t'], lr)
    return ({'loss': float(torch.sqrt(2 * loss)), 'entropy': float(ent), 'raw_loss': float(loss)},
            OrderedDict(zip(names, grads)), y.detach(), q.detach())


# ------------------------------------------------------------------------------------------------ INet / DNet / ClassicISP
def _conv_named(P, t, name, act=None, padding='SAME', pad=0):
    """Keras Conv2D by parameter name; pad > 0: tf.pad(REFLECT) then a VALID convolution."""
    if pad > 0:
        t, padding = R.tf_pad(t, pad, 'REFLECT'), 'VALID'
    y = R.conv2d(t, P[name + '/kernel'], P.get(name + '/bias'), padding=padding)
    return R.ACT[act](y) if act in R.ACT else y


def inet_forward(P, x, kernel=5):
    """models/pipelines.py:273-292."""
    bayer = R.depth_to_space(_conv_named(P, x, 'upsampling'), 2)
    rgb = _conv_named(P, bayer, 'demosaicing', pad=(kernel - 1) // 2)
    srgb = _conv_named(P, rgb, 'srgb')
    g0 = _conv_named(P, srgb, 'gamma_d1', 'tanh')
    return R.ste_clip(_conv_named(P, g0, 'gamma_d2'))


def dnet_forward(P, x, n_layers=15, kernel=3):
    """models/pipelines.py:318-345: conv VALID + ReLU followed by a REFLECT pad, n_layers times; d2s; concat; project; pad; 1x1."""
    pad = (kernel - 1) // 2
    deep = x
    for r in range(n_layers):
        deep = R.tf_pad(_conv_named(P, deep, 'conv2d_%d' % r, 'relu', padding='VALID'), pad, 'REFLECT')
    bayer = R.depth_to_space(_conv_named(P, x, 'upsampling'), 2)
    cat = torch.cat((R.depth_to_space(deep, 2), bayer), dim=3)
    pu = R.tf_pad(_conv_named(P, cat, 'conv2d_%d' % n_layers, 'relu', padding='VALID'), pad, 'REFLECT')
    return R.ste_clip(_conv_named(P, pu, 'conv2d_%d' % (n_layers + 1), padding='VALID'))


def classic_isp_forward(P, x, kernel=5, n_cnn=0, residual=True):
    """_ClassicISP.call (models/pipelines.py:432-446) over DemosaicingLayer.call (models/layers.py:238-258).
    n_cnn = len(c_filters); the CNN branch has n_cnn k x k layers + the final 1x1."""
    bayer = R.depth_to_space(_conv_named(P, x, 'upsampling'), 2)

    def cnn(t):
        for i in range(n_cnn):
            t = _conv_named(P, t, 'demosaicing/conv2d_%d' % i, 'leaky_relu')
        return _conv_named(P, t, 'demosaicing/conv2d_%d' % n_cnn, 'tanh' if residual else 'sigmoid')
    if residual:
        xb = _conv_named(P, bayer, 'demosaicing/bilinear', pad=(kernel - 1) // 2)
        f = cnn(bayer) if n_cnn > 0 else 0
        y = xb - P['demosaicing/alpha'] * f
    else:
        y = cnn(bayer)
    y = R.ste_clip(y)
    rgb = _conv_named(P, y, 'srgb')
    return torch.pow(R.ste_clip(rgb, 1.0 / 255, 1.0), 1 / 2.2)
