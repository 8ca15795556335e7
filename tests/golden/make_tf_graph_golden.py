"""Generate tests/golden/tf_graph_golden.npz by EXECUTING the reference's own, unmodified TensorFlow model code.

Runs only in the build container (needs /root/reference). TensorFlow cannot be installed here, so `tests/tf_shim/tensorflow` — a
tests-only stand-in on torch-CPU tensors — supplies the ~90 tf.* / Keras entry points those files call; everything else (graph
wiring, constants, block order, loss composition, what carries gradient, optimizer calls) is the reference's code, imported from
where it lies:
    models/jpeg.py, models/layers.py, models/pipelines.py, models/forensics.py, models/compression.py, models/tfmodel.py,
    helpers/tf_helpers.py, helpers/kernels.py, helpers/paramspec.py, helpers/utils.py, compression/jpeg_helpers.py, compression/codec.py,
    workflows/manipulation_classification.py
Every case is run twice: float64 (`tf.set_precision('float64')`, the truth the parity tests compare against) and float32 (the
reference's dtype; only its distance from the float64 run is kept, as the drift budget). Large tensors are kept as 1024 seeded
samples + L2 norm + sum (tests/golden/tfgraph_common.py).

Environment adaptations (none touches reference source): NumPy-1.x aliases np.bool/np.float/np.int; scipy.signal.gaussian ->
scipy.signal.windows.gaussian; import stubs for absent third-party packages (tests/tf_shim/*); `jpeg_qtable` results cast to float32,
which is what NumPy 1.18 (the reference's pin) returns for float32-array x scalar — checked below to give identical table values.

    python tests/golden/make_tf_graph_golden.py
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, os.path.join(ROOT, 'tests', 'tf_shim'))
sys.path.insert(1, REF)
sys.path.insert(2, ROOT)
sys.path.insert(3, HERE)

import scipy.cluster.vq  # noqa: E402,F401
import scipy.fftpack  # noqa: E402,F401
import scipy.signal  # noqa: E402
import scipy.signal.windows  # noqa: E402
import scipy.stats  # noqa: E402,F401

np.bool = bool          # noqa: E305  NumPy 1.18 aliases used by helpers/utils.py:25
np.float = float
np.int = int
scipy.signal.gaussian = scipy.signal.windows.gaussian

import tensorflow as tf  # noqa: E402  (the shim)
import torch  # noqa: E402
from compression import jpeg_helpers  # noqa: E402
from helpers import tf_helpers  # noqa: E402
from models import compression, forensics, jpeg, layers, pipelines  # noqa: E402
from workflows import manipulation_classification as mc  # noqa: E402

import tfgraph_common as C  # noqa: E402

assert tf.__version__.endswith('torch-shim')

# ---- NumPy-1.18 behaviour of jpeg_qtable: float32 tables (see module docstring)
_qtable = jpeg_helpers.jpeg_qtable


def _qtable32(quality, channel=0):
    t = _qtable(quality, channel)
    q = np.maximum(np.minimum(100, quality), 1)
    s = np.float32(5000 / q if q < 50 else 200 - q * 2)          # legacy promotion: the scalar joins the float32 array's dtype
    base = _qtable(50, channel).astype(np.float32)               # scale 100 -> the base table itself
    t32 = np.floor((base * s + np.float32(50)) / np.float32(100))
    t32[t32 < 1] = 1
    t32[t32 > 255] = 255
    assert np.array_equal(t32, t), 'float32 and float64 evaluation of jpeg_qtable({}, {}) differ'.format(quality, channel)
    return t.astype(np.float32)


for _q in range(1, 101):
    _qtable32(_q, 0), _qtable32(_q, 1)
jpeg.jpeg_qtable = _qtable32
jpeg_helpers.jpeg_qtable = _qtable32

from neural_imaging_b200 import nn as product_nn  # noqa: E402

product_nn.HOST_ONLY = True
from neural_imaging_b200.models import compression as p_compression  # noqa: E402
from neural_imaging_b200.models import forensics as p_forensics  # noqa: E402
from neural_imaging_b200.models import pipelines as p_pipelines  # noqa: E402

OUT = {}
META = {}


def record(case, runs):
    """runs = {'float64': {name: array}, 'float32': {...}}"""
    r64, r32 = runs['float64'], runs['float32']
    for name, a in r64.items():
        key = '{}/{}'.format(case, name)
        s = C.summarize(a, key)
        OUT[key + '/v'], OUT[key + '/n'] = s['v'], s['n']
        s32 = C.summarize(r32[name], key)
        scale = max(float(np.max(np.abs(s['v']))), 1e-30) if s['v'].size else 1.0
        err = np.abs(s32['v'] - s['v']) / scale
        # drift of the reference's own float32 run from its float64 run: [max, 98 % quantile] (the quantile ignores isolated decision flips)
        OUT[key + '/d'] = np.array([float(err.max()) if err.size else 0.0, float(np.quantile(err, 0.98)) if err.size else 0.0])


def both(fn):
    runs = {}
    for prec in ('float64', 'float32'):
        tf.set_precision(prec)
        tf.keras.backend.clear_session()
        tf.random.set_seed(1234)
        tf.random.noise_log.clear()
        jpeg._common_codec = None
        np.random.seed(1234)
        runs[prec] = {k: np.asarray(v, dtype=np.float64) for k, v in fn(np.float64 if prec == 'float64' else np.float32).items()}
    tf.set_precision('float32')
    return runs


def T(a, dt, grad=False):
    t = tf.convert_to_tensor(np.asarray(a, dtype=dt))
    if grad:
        t.requires_grad_(True)
    return t


def grad_of(scalar, xs):
    if not scalar.requires_grad:          # e.g. manipulation_sharpen: RGBToHSV / HSVToRGB carry no gradient in TF 2.1 -> tape.gradient gives None
        return [np.zeros(tuple(x.shape)) for x in xs]
    g = torch.autograd.grad(scalar.as_subclass(torch.Tensor), [x for x in xs], allow_unused=True, retain_graph=True)
    return [np.zeros(tuple(x.shape)) if e is None else e.detach().numpy() for e, x in zip(g, xs)]


def N(t):
    return t.numpy() if hasattr(t, 'numpy') else np.asarray(t)


def load_state(keras_vars, product_specs, state):
    """Assign product-layout weights to the Keras variables (trainable ones in order; frozen ones are asserted equal)."""
    tr = [(n, s) for n, s, t, _ in product_specs if t]
    assert len(tr) == len(keras_vars), '{} product vs {} keras trainable variables'.format(len(tr), len(keras_vars))
    for (name, shape), v in zip(tr, keras_vars):
        v.assign(C.product_to_keras(state[name], tuple(v.shape)))


def grads_to_product(grads, product_specs):
    tr = [(n, s) for n, s, t, _ in product_specs if t]
    return {n: C.keras_to_product(N(g), s) for (n, s), g in zip(tr, grads)}


def check_constant_inits(keras_vars, product_specs, what):
    """The reference's constant initial values (helpers/kernels.py etc.) == the product's initial values."""
    tr = [(n, s, i) for n, s, t, i in product_specs if t]
    for (name, shape, init), v in zip(tr, keras_vars):
        if name in C.CONST_INIT:
            a = C.keras_to_product(N(v), shape)
            assert np.allclose(a, init, rtol=0, atol=1e-7), '{}: constant init of {} differs from the reference'.format(what, name)


# ================================================================================================================ dJPEG
def case_djpeg():
    rs = np.random.RandomState(11)
    x = rs.uniform(size=(2, 16, 24, 3)).astype(np.float32)
    w = rs.normal(size=x.shape).astype(np.float32)
    META['djpeg'] = {'seed': 11, 'shape': list(x.shape), 'draws': ['x uniform', 'w normal'],
                     'cases': [[50, 'soft'], [50, 'sin'], [80, 'harmonic'], [80, 'soft'], [30, 'sin'], [95, 'soft'], [10, 'harmonic']]}
    for q, mode in META['djpeg']['cases']:
        def run(dt):
            m = jpeg.DifferentiableJPEG(q, mode)
            xt = T(x, dt, True)
            y, X = m(xt)
            dx, = grad_of((y * T(w, dt)).sum(), [xt])
            # JPEG.process with a quality different from the constructor's swaps the tables and must give the same image (models/jpeg.py:235-243)
            y2 = jpeg.JPEG(77 if q != 77 else 50, mode).process(T(x, dt), quality=q)
            assert np.array_equal(N(y2), N(y))
            return {'y': N(y), 'X': N(X), 'dx': dx}
        record('djpeg_q{}_{}'.format(q, mode), both(run))
    # trainable quantisation tables (models/jpeg.py:58-62): gradients w.r.t. the two 8x8 weights
    META['djpeg']['trainable_cases'] = [[50, 'sin'], [80, 'soft'], [30, 'harmonic']]
    for q, mode in META['djpeg']['trainable_cases']:
        def run_t(dt):
            m = jpeg.DifferentiableJPEG(q, mode, trainable=True)
            assert len(m.trainable_weights) == 2
            xt = T(x, dt, True)
            y, X = m(xt)
            dx, dql, dqc = grad_of((y * T(w, dt)).sum(), [xt, m._q_mtx_luma, m._q_mtx_chroma])
            return {'y': N(y), 'dx': dx, 'dq_luma': dql, 'dq_chroma': dqc}
        record('djpeg_trainable_q{}_{}'.format(q, mode), both(run_t))
    # the lazily created module-level codec of the 'jpeg' manipulation is JPEG(None, 'soft') (models/jpeg.py:38-42)

    def run_common(dt):
        y = jpeg.differentiable_jpeg(T(x, dt), 80)
        assert jpeg._common_codec.codec == 'soft' and jpeg._common_codec.quality is None
        return {'y': N(y)}
    record('djpeg_common_q80', both(run_common))


# ================================================================================================================ manipulations
def case_manipulations():
    rs = np.random.RandomState(12)
    x = rs.uniform(size=(2, 32, 32, 3)).astype(np.float32)
    xr = rs.uniform(size=(2, 16, 24, 3)).astype(np.float32)          # non-square: catches H/W transpositions
    w = rs.normal(size=x.shape).astype(np.float32)
    wr = rs.normal(size=xr.shape).astype(np.float32)
    noise = rs.normal(size=x.shape).astype(np.float32)
    META['manip'] = {'seed': 12, 'draws': ['x (2,32,32,3) uniform', 'xr (2,16,24,3) uniform', 'w normal', 'wr normal', 'noise normal'],
                     'ops': {}}
    ops = {
        'sharpen_1': (lambda t: tf_helpers.manipulation_sharpen(t, 1, hsv=True), 'sq'),
        'sharpen_0p4': (lambda t: tf_helpers.manipulation_sharpen(t, 0.4, hsv=True), 'rect'),
        'resample_50': (lambda t: tf_helpers.manipulation_resample(t, 50), 'sq'),
        'resample_75': (lambda t: tf_helpers.manipulation_resample(t, 75), 'sq'),
        'resample_0p6': (lambda t: tf_helpers.manipulation_resample(t, 0.6), 'sq'),
        'gaussian_0p83': (lambda t: tf_helpers.manipulation_gaussian(t, 5, 0.83), 'sq'),
        'gaussian_2p5': (lambda t: tf_helpers.manipulation_gaussian(t, 5, 2.5), 'rect'),
        'gaussian_k3': (lambda t: tf_helpers.manipulation_gaussian(t, 3, 1.0), 'rect'),
        'gamma_3': (lambda t: tf_helpers.manipulation_gamma(t, 3.0), 'sq'),
        'gamma_0p7': (lambda t: tf_helpers.manipulation_gamma(t, 0.7), 'rect'),
        'median_3': (lambda t: tf_helpers.manipulation_median(t, 3), 'sq'),
        'median_5': (lambda t: tf_helpers.manipulation_median(t, 5), 'rect'),
        'median_4': (lambda t: tf_helpers.manipulation_median(t, 4), 'rect'),
        'soft_quantization': (lambda t: tf_helpers.soft_quantization(t), 'sq'),
        'quantize_and_clip': (lambda t: tf_helpers.quantize_and_clip(t * 1.2 - 0.1), 'sq'),
    }
    for name, (f, kind) in ops.items():
        def run(dt):
            xx, ww = (x, w) if kind == 'sq' else (xr, wr)
            xt = T(xx, dt, True)
            y = f(xt)
            dx, = grad_of((y * T(ww, dt)).sum(), [xt])
            return {'y': N(y), 'dx': dx}
        record('manip_' + name, both(run))
        META['manip']['ops'][name] = kind

    # AWGN: tf.random.normal cannot be reproduced; the shim records the noise it drew and the fixture pins y given that noise
    def run_awgn(dt):
        real = tf.random.normal
        tf.random.normal = lambda shape, *a, **k: T(noise, dt)
        try:
            xt = T(x, dt, True)
            y = tf_helpers.manipulation_awgn(xt, 5.1 / 255)
            dx, = grad_of((y * T(w, dt)).sum(), [xt])
        finally:
            tf.random.normal = real
        return {'y': N(y), 'dx': dx}
    record('manip_awgn_5p1', both(run_awgn))


# ================================================================================================================ layers
def case_layers():
    rs = np.random.RandomState(13)
    x = rs.uniform(size=(2, 12, 14, 3)).astype(np.float32)
    w = rs.normal(size=x.shape).astype(np.float32)
    k = rs.normal(size=(5, 5, 3, 3)).astype(np.float32)
    z = (rs.normal(size=(2, 4, 4, 8)) * 3).astype(np.float32)
    wz = rs.normal(size=z.shape).astype(np.float32)
    META['layers'] = {'seed': 13, 'draws': ['x (2,12,14,3) uniform', 'w normal', 'k (5,5,3,3) normal', 'z (2,4,4,8) normal*3', 'wz normal']}

    def run_cc(dt):
        layer = layers.ConstrainedConv2D()
        init = N(layer.kernel)
        layer.kernel.assign(init + 0.3 * k)
        xt = T(x, dt, True)
        y = layer(xt)
        dx, dk = grad_of((y * T(w, dt)).sum(), [xt, layer.kernel])
        return {'init': init, 'y': N(y), 'dx': dx, 'dkernel': dk}
    record('constrained_conv2d', both(run_cc))

    for v, gamma, bpf in ((50, 25, 5), (0, 5, 4), (50, 25, 3)):
        def run_dl(dt):
            layer = layers.DiscreteLatent('soft-codebook', v=v, gamma=gamma, latent_bpf=bpf)
            layer.scaling_factor.assign(np.float32(0.8))
            zt = T(z, dt, True)
            lat, ent = layer(zt)
            dz, ds = grad_of((lat * T(wz, dt)).sum() + 7.0 * ent, [zt, layer.scaling_factor])
            return {'latent': N(lat), 'entropy': N(ent), 'dz': dz, 'dscale': ds, 'codebook': N(layer.quantization.codebook)}
        record('discrete_latent_v{}_g{}_b{}'.format(v, gamma, bpf), both(run_dl))

    for mode in ('round', 'sin', 'soft', 'identity', 'harmonic'):
        def run_q(dt):
            layer = layers.Quantization(mode)
            zt = T(z, dt, True)
            y = layer(zt)
            g = grad_of((y * T(wz, dt)).sum(), [zt])[0] if mode != 'round' else np.zeros(z.shape)
            return {'y': N(y), 'dz': g}
        record('quantization_' + mode, both(run_q))

    # losses
    a = rs.uniform(size=(2, 48, 40, 3)).astype(np.float32)
    b = np.clip(a + 0.1 * rs.normal(size=a.shape), 0, 1).astype(np.float32)
    am = rs.uniform(size=(1, 192, 176, 3)).astype(np.float32)
    bm = np.clip(am + 0.05 * rs.normal(size=am.shape), 0, 1).astype(np.float32)
    META['layers']['draws'] += ['a (2,48,40,3) uniform', 'b = clip(a + 0.1 normal)', 'am (1,192,176,3) uniform', 'bm = clip(am + 0.05 normal)']
    for name, f, (p, q) in (('mse', tf_helpers.mse, (a, b)), ('mae', tf_helpers.mae, (a, b)), ('ssim_loss', tf_helpers.ssim_loss, (a, b)),
                            ('msssim_loss', tf_helpers.msssim_loss, (am, bm))):
        def run_loss(dt):
            pt = T(p, dt, True)
            val = f(pt, T(q, dt))
            dp, = grad_of(val, [pt])
            return {'loss': N(val), 'da': dp}
        record('loss_' + name, both(run_loss))

    def run_ssim(dt):
        return {'ssim': N(tf.image.ssim(T(a, dt), T(b, dt), max_val=1))}
    record('tf_image_ssim', both(run_ssim))


# ================================================================================================================ models
def case_nip_models():
    for cls, kw, ps, seed in (('UNet', {}, 32, 21), ('INet', {}, 16, 22), ('DNet', {'n_layers': 3}, 16, 23),
                              ('ClassicISP', {'c_filters': (8,)}, 16, 24), ('ClassicISP', {'c_filters': ()}, 16, 25),
                              ('INet', {'cfa_pattern': 'rggb', 'kernel': 3}, 16, 26), ('UNet', {'n_steps': 3, 'activation': 'relu'}, 16, 27)):
        case = 'nip_{}_{}'.format(cls, seed)
        pm = getattr(p_pipelines, cls)(patch_size=ps, seed=1, **kw)
        specs = C.specs_of(pm)
        state = C.golden_state(specs, seed, ones_names=('conv2d_4/kernel',) if cls == 'DNet' else ())
        rs = np.random.RandomState(seed)
        x = rs.uniform(size=(2, ps, ps, 4)).astype(np.float32)
        t = rs.uniform(size=(2, 2 * ps, 2 * ps, 3)).astype(np.float32)
        META[case] = {'cls': cls, 'kw': {k: list(v) if isinstance(v, tuple) else v for k, v in kw.items()}, 'patch_size': ps, 'seed': seed,
                      'draws': ['x (2,ps,ps,4) uniform', 't (2,2ps,2ps,3) uniform'], 'lr': 1e-3}

        def run(dt):
            m = getattr(pipelines, cls)(patch_size=ps, **kw)
            m.process(T(x, dt))          # subclassed Keras models (_ClassicISP) create their variables on the first call
            if tf.float32.torch == torch.float32:
                check_constant_inits(m.parameters, specs, case)
            load_state(m.parameters, specs, state)
            xt = T(x, dt, True)
            y = m.process(xt)
            loss = m.loss(y, T(t, dt))
            g = grad_of(loss, [xt] + list(m.parameters))
            out = {'y': N(y), 'loss': N(loss), 'dx': g[0]}
            for n, a in grads_to_product(g[1:], specs).items():
                out['grad/' + n] = a
            # two optimizer steps through the reference's own training_step (tape + Keras Adam)
            l1 = m.training_step(T(x, dt), T(t, dt), 1e-3)
            l2 = m.training_step(T(x, dt), T(t, dt), 5e-4)
            out['step_loss'] = np.array([float(N(l1)), float(N(l2))])
            for n, a in grads_to_product([N(v) for v in m.parameters], specs).items():
                out['param2/' + n] = a
            return out
        record(case, both(run))


def case_fan():
    for kw, seed, ps in (({}, 31, 32), (dict(n_filters=8, n_convolutions=2, kernel=3, n_dense=2, use_gap=False, activation='relu'), 32, 32),
                         (dict(n_filters=16, n_fscale=1.5, n_convolutions=3, n_dense=1, activation='tanh'), 33, 16)):
        case = 'fan_{}'.format(seed)
        pm = p_forensics.FAN(n_classes=5, patch_size=ps, seed=1, **kw)
        specs = C.specs_of(pm)
        state = C.golden_state(specs, seed)
        rs = np.random.RandomState(seed)
        x = rs.uniform(size=(6, ps, ps, 3)).astype(np.float32)
        labels = rs.randint(0, 5, size=(6,))
        META[case] = {'kw': kw, 'patch_size': ps, 'seed': seed, 'draws': ['x (6,ps,ps,3) uniform', 'labels randint(0,5,(6,))'], 'lr': 1e-3}

        def run(dt):
            m = forensics.FAN(n_classes=5, patch_size=ps, **kw)
            if tf.float32.torch == torch.float32:
                check_constant_inits(m.parameters, specs, case)
            load_state(m.parameters, specs, state)
            xt = T(x, dt, True)
            p = m.process(xt)
            loss = m.loss(labels, p)
            g = grad_of(loss, [xt] + list(m.parameters))
            out = {'probs': N(p), 'loss': N(loss), 'dx': g[0], 'decide': m.process_and_decide(T(x, dt)).astype(np.float64)}
            for n, a in grads_to_product(g[1:], specs).items():
                out['grad/' + n] = a
            l1 = m.training_step(T(x, dt), labels, 1e-3)
            l2 = m.training_step(T(x, dt), labels, 5e-4)
            out['step_loss'] = np.array([float(N(l1)), float(N(l2))])
            for n, a in grads_to_product([N(v) for v in m.parameters], specs).items():
                out['param2/' + n] = a
            return out
        record(case, both(run))


def case_dcn():
    for kw, seed, ps in (({}, 41, 32), (dict(n_features=8, latent_bpf=3, entropy_weight=100), 42, 16),
                         (dict(n_features=8, rounding='sin'), 43, 16), (dict(n_features=8, rounding='soft', latent_bpf=4), 44, 16),
                         (dict(n_features=4, rounding='identity', entropy_weight=50), 45, 16)):
        case = 'dcn_{}'.format(seed)
        pm = p_compression.TwitterDCN(patch_size=ps, seed=1, **kw)
        specs = C.specs_of(pm)
        state = C.golden_state(specs, seed)
        rs = np.random.RandomState(seed)
        x = rs.uniform(size=(2, ps, ps, 3)).astype(np.float32)
        META[case] = {'kw': kw, 'patch_size': ps, 'seed': seed, 'draws': ['x (2,ps,ps,3) uniform'], 'lr': 1e-3}

        def run(dt):
            m = compression.TwitterDCN(patch_size=ps, **kw)
            load_state(m.parameters, specs, state)
            xt = T(x, dt, True)
            y, ent = m.process(xt, return_entropy=True)
            loss = m.loss(xt, y, ent)
            g = grad_of(loss, [xt] + list(m.parameters))
            z = m.compress(T(x, dt))
            out = {'y': N(y), 'entropy': N(ent), 'loss': N(loss), 'dx': g[0], 'latent': N(z), 'decompressed': N(m.decompress(z)),
                   'codebook': m.get_codebook().astype(np.float64)}
            for n, a in grads_to_product(g[1:], specs).items():
                out['grad/' + n] = a
            s1 = m.training_step(T(x, dt), 1e-3)
            s2 = m.training_step(T(x, dt), 5e-4)
            out['step_loss'] = np.array([float(s1['loss']), float(s2['loss'])])
            out['step_ssim'] = np.array([float(N(s1['ssim'])), float(N(s2['ssim']))])
            out['step_entropy'] = np.array([float(N(s1['entropy'])), float(N(s2['entropy']))])
            for n, a in grads_to_product([N(v) for v in m.parameters], specs).items():
                out['param2/' + n] = a
            return out
        record(case, both(run))


# ================================================================================================================ joint workflow
def _flow_specs(ps, n_classes, nip='UNet'):
    pn = getattr(p_pipelines, nip)(patch_size=ps, seed=1)
    pf = p_forensics.FAN(n_classes=n_classes, patch_size=2 * ps // 2, seed=1)
    return C.specs_of(pn), C.specs_of(pf)


def case_workflow():
    ps, B = 32, 2
    variants = {
        # the BASELINE configuration (config 4) at a small patch: default manipulations, pool:2, dJPEG(50, 'soft'), trainable {fan, nip}
        'flow_default': dict(manipulations=None, codec='soft', trainable={'nip'}, lambda_nip=0.1, down='pool:2'),
        # continuous codec and no hard-rounding manipulation: every operation on the path is continuous -> tight tolerances
        'flow_sin': dict(manipulations=['sharpen', 'resample', 'gaussian'], codec='sin', trainable={'nip'}, lambda_nip=0.1, down='pool:2'),
        'flow_fan_only': dict(manipulations=['resample:70', 'gaussian:1.5', 'gamma', 'median'], codec='harmonic', trainable=set(), lambda_nip=0.0,
                              down='bilinear'),
    }
    for case, v in variants.items():
        manips = v['manipulations']
        n_classes = 1 + len(manips or ['sharpen', 'resample', 'gaussian', 'jpeg'])
        s_nip, s_fan = _flow_specs(ps, n_classes)
        seed = 50 + len(case)
        st_nip, st_fan = C.golden_state(s_nip, seed), C.golden_state(s_fan, seed + 1)
        rs = np.random.RandomState(seed)
        x = rs.uniform(size=(B, ps, ps, 4)).astype(np.float32)
        t = rs.uniform(size=(B, 2 * ps, 2 * ps, 3)).astype(np.float32)
        META[case] = {'manipulations': manips, 'codec': v['codec'], 'trainable': sorted(v['trainable']), 'lambda_nip': v['lambda_nip'],
                      'downsampling': v['down'], 'patch_size': ps, 'batch': B, 'seed': seed, 'fan_seed': seed + 1,
                      'draws': ['x (B,ps,ps,4) uniform', 't (B,2ps,2ps,3) uniform'], 'lr': [1e-4, 5e-5]}

        def run(dt):
            flow = mc.ManipulationClassification('UNet', manipulations=manips,
                                                 distribution={'downsampling': v['down'], 'compression': 'jpeg',
                                                               'compression_params': {'quality': 50, 'codec': v['codec']}},
                                                 fan_args={}, trainable=v['trainable'], raw_patch_size=ps)
            load_state(flow.nip.parameters, s_nip, st_nip)
            load_state(flow.fan.parameters, s_fan, st_fan)
            Y, c, Cc, ent, probs = flow.run_workflow(T(x, dt))
            assert np.isnan(ent)
            out = {'Y': N(Y), 'c': N(c), 'C': N(Cc), 'probs': N(probs), 'labels': flow._batch_labels(B).astype(np.float64),
                   'decisions': flow.run_workflow_to_decisions(T(x, dt)).astype(np.float64)}
            loss, parts = flow.training_step(T(x, dt), T(t, dt), lambda_nip=v['lambda_nip'], learning_rate=1e-4)
            src, grads = tf.GradientTape.last
            assert np.isnan(float(N(parts['dcn'])))                          # JPEG.loss receives entropy = NaN as sample_weight
            n_fan = len(flow.fan.parameters)
            for n, a in grads_to_product(grads[:n_fan], s_fan).items():
                out['grad/fan/' + n] = a
            if 'nip' in v['trainable']:
                for n, a in grads_to_product(grads[n_fan:], s_nip).items():
                    out['grad/nip/' + n] = a
            else:
                assert len(grads) == n_fan
            loss2, parts2 = flow.training_step(T(x, dt), T(t, dt), lambda_nip=v['lambda_nip'], learning_rate=5e-5)
            out['loss'] = np.array([float(N(loss)), float(N(loss2))])
            out['ce'] = np.array([float(N(parts['ce'])), float(N(parts2['ce']))])
            out['nip'] = np.array([float(N(parts['nip'])), float(N(parts2['nip']))])
            for n, a in grads_to_product([N(p) for p in flow.fan.parameters], s_fan).items():
                out['param2/fan/' + n] = a
            for n, a in grads_to_product([N(p) for p in flow.nip.parameters], s_nip).items():
                out['param2/nip/' + n] = a
            return out
        record(case, both(run))


def case_workflow_dcn():
    """config 5: compression='dcn' restored through compression.codec.restore -> models.tfmodel.restore from a model directory
    (written here with the shim's weight files), trainable {fan, nip, dcn}, lambda_dcn = 0.1 (config/tests/framework.json:55)."""
    from compression import codec as ref_codec
    ps, B, case = 32, 2, 'flow_dcn'
    s_nip, s_fan = _flow_specs(ps, 5)
    s_dcn = C.specs_of(p_compression.TwitterDCN(patch_size=ps, seed=1))
    seed = 61
    st_nip, st_fan, st_dcn = C.golden_state(s_nip, seed), C.golden_state(s_fan, seed + 1), C.golden_state(s_dcn, seed + 2)
    rs = np.random.RandomState(seed)
    x = rs.uniform(size=(B, ps, ps, 4)).astype(np.float32)
    t = rs.uniform(size=(B, 2 * ps, 2 * ps, 3)).astype(np.float32)
    META[case] = {'patch_size': ps, 'batch': B, 'seed': seed, 'fan_seed': seed + 1, 'dcn_seed': seed + 2, 'lambda_nip': 0.1, 'lambda_dcn': 0.1,
                  'draws': ['x (B,ps,ps,4) uniform', 't (B,2ps,2ps,3) uniform'], 'lr': [1e-4, 5e-5]}

    def run(dt):
        with tempfile.TemporaryDirectory() as d:
            m = compression.TwitterDCN(patch_size=ps)
            load_state(m.parameters, s_dcn, st_dcn)
            m.save_model(d, save_args=False, quiet=True)
            with open(os.path.join(d, 'progress.json'), 'w') as f:
                json.dump({'codec': {'model': 'TwitterDCN', 'args': m.get_hyperparameters(), 'performance': {}}}, f)
            flow = mc.ManipulationClassification('UNet', distribution={'downsampling': 'pool:2', 'compression': 'dcn',
                                                                       'compression_params': {'dirname': d}},
                                                 fan_args={}, trainable={'nip', 'dcn'}, raw_patch_size=ps)
        assert flow.codec.patch_size is None          # codec.restore passes patch_size=None
        load_state(flow.nip.parameters, s_nip, st_nip)
        load_state(flow.fan.parameters, s_fan, st_fan)
        for a, b in zip(flow.codec.parameters, m.parameters):
            assert np.array_equal(N(a), N(b))
        Y, c, Cc, ent, probs = flow.run_workflow(T(x, dt))
        out = {'Y': N(Y), 'c': N(c), 'C': N(Cc), 'entropy': N(ent), 'probs': N(probs)}
        loss, parts = flow.training_step(T(x, dt), T(t, dt), lambda_nip=0.1, lambda_dcn=0.1, learning_rate=1e-4)
        src, grads = tf.GradientTape.last
        n_fan, n_nip = len(flow.fan.parameters), len(flow.nip.parameters)
        for n, a in grads_to_product(grads[:n_fan], s_fan).items():
            out['grad/fan/' + n] = a
        for n, a in grads_to_product(grads[n_fan:n_fan + n_nip], s_nip).items():
            out['grad/nip/' + n] = a
        for n, a in grads_to_product(grads[n_fan + n_nip:], s_dcn).items():
            out['grad/dcn/' + n] = a
        loss2, parts2 = flow.training_step(T(x, dt), T(t, dt), lambda_nip=0.1, lambda_dcn=0.1, learning_rate=5e-5)
        out['loss'] = np.array([float(N(loss)), float(N(loss2))])
        for k in ('ce', 'nip', 'dcn'):
            out[k] = np.array([float(N(parts[k])), float(N(parts2[k]))])
        for n, a in grads_to_product([N(p) for p in flow.codec.parameters], s_dcn).items():
            out['param2/dcn/' + n] = a
        for n, a in grads_to_product([N(p) for p in flow.fan.parameters], s_fan).items():
            out['param2/fan/' + n] = a
        return out
    record(case, both(run))
    # reference `codec` helpers touched on the way
    assert callable(ref_codec.restore)


def main():
    case_djpeg()
    case_manipulations()
    case_layers()
    case_nip_models()
    case_fan()
    case_dcn()
    case_workflow()
    case_workflow_dcn()
    OUT['meta'] = np.frombuffer(json.dumps(META, sort_keys=True).encode(), dtype=np.uint8)
    path = os.path.join(HERE, 'tf_graph_golden.npz')
    np.savez_compressed(path, **OUT)
    cases = sorted({k.split('/')[0] for k in OUT if k != 'meta'})
    print('wrote {} ({:.2f} MB): {} cases, {} tensors'.format(path, os.path.getsize(path) / 1e6, len(cases), sum(k.endswith('/v') for k in OUT)))
    worst = sorted(((float(OUT[k][0]), k) for k in OUT if k.endswith('/d')), reverse=True)[:12]
    print('largest float32-vs-float64 drifts of the executed reference:')
    for d, k in worst:
        print('  {:.3e}  {}'.format(d, k[:-2]))


if __name__ == '__main__':
    main()
