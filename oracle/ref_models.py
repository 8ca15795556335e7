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


def workflow_forward(P_nip, P_fan, x, names=('sharpen', 'resample', 'gaussian', 'jpeg'), quality=50, pool=2, nip='UNet'):
    """run_workflow (:162-176) for the default distribution channel (pool:k + dJPEG(quality,'soft'))."""
    Y = unet_forward(P_nip, x) if nip == 'UNet' else x
    m = run_manipulations(Y, names)
    c = R.avg_pool(m, pool) if pool > 1 else m
    C = R.djpeg(c, R.jpeg_qtable(quality, 0), R.jpeg_qtable(quality, 1), 'soft')[0] if quality else c
    probs = fan_forward(P_fan, C)
    return Y, c, C, probs


def batch_labels(batch_size, n_classes):
    return np.concatenate([k * np.ones((batch_size,), dtype=np.int64) for k in range(n_classes)])


def training_step(P_nip, P_fan, opt_state, x, y_target, lambda_nip=0.1, lr=1e-4, train_nip=True,
                  names=('sharpen', 'resample', 'gaussian', 'jpeg'), quality=50, pool=2, nip='UNet'):
    """ManipulationClassification.training_step (:260-285): loss = ce + lambda_nip * mse; shared Keras Adam.
    opt_state = {'t': int, 'm': {name: tensor}, 'v': {...}}; parameters are updated in place. Returns loss dict + grads."""
    Y, c, C, probs = workflow_forward(P_nip, P_fan, x, names, quality, pool, nip)
    n_classes = len(names) + 1
    loss_ce = R.sparse_categorical_crossentropy(batch_labels(x.shape[0], n_classes), probs)
    loss_nip = R.mse(y_target, Y)
    loss = loss_ce + (lambda_nip * loss_nip if train_nip else 0)
    params = [('fan/' + k, v) for k, v in P_fan.items()]
    if train_nip and nip == 'UNet':
        params += [('nip/' + k, v) for k, v in P_nip.items()]
    grads = torch.autograd.grad(loss, [p for _, p in params], allow_unused=True)
    grads = [torch.zeros_like(p) if g is None else g for g, (_, p) in zip(grads, params)]
    opt_state['t'] += 1
    with torch.no_grad():
        ms = [opt_state['m'].setdefault(k, torch.zeros_like(p)) for k, p in params]
        vs = [opt_state['v'].setdefault(k, torch.zeros_like(p)) for k, p in params]
        R.adam_keras_step([p for _, p in params], grads, ms, vs, opt_state['t'], lr)
    return ({'loss': float(loss), 'ce': float(loss_ce), 'nip': float(loss_nip)},
            OrderedDict((k, g) for (k, _), g in zip(params, grads)))
