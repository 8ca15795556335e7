"""The oracle pinned to the EXECUTED reference (CPU): `tests/golden/tf_graph_golden.npz` holds what the reference's own, unmodified model
code (models/jpeg.py, models/layers.py, helpers/tf_helpers.py, models/pipelines.py, models/forensics.py, models/compression.py,
workflows/manipulation_classification.py) computed when run through the tests-only TensorFlow stand-in `tests/tf_shim`
(`tests/golden/make_tf_graph_golden.py`). Here the float64 oracle restatement (`oracle/ref_ops.py`, `ref_models.py`) must reproduce
those values — outputs, losses, every gradient, parameters after two optimizer steps — to 1e-9 where the computation is float64
throughout, and 2e-6 where the reference itself casts to float32 in the middle (the latent path). What this moves from "re-typed" to
"executed reference": the graph wiring (transposes, block order, constants, paddings, loss composition, gradient-carrying branches,
optimizer calls). What stays restated (in the shim): the per-op semantics of tf.* kernels, listed in DESIGN.md section 2."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from golden import tfgraph_common as C
from oracle import ref_models as M
from oracle import ref_ops as R

from neural_imaging_b200 import nn as product_nn

TIGHT = 1e-9


@pytest.fixture(scope='module')
def G():
    g = C.Golden(os.path.join(GOLDEN, 'tf_graph_golden.npz'))
    g.meta = json.loads(bytes(g.d['meta']).decode())
    return g


@pytest.fixture(scope='module')
def host_models():
    """Product model classes in HOST_ONLY mode (parameter specs without device buffers)."""
    old = product_nn.HOST_ONLY
    product_nn.HOST_ONLY = True
    from neural_imaging_b200.models import compression, forensics, pipelines
    yield {'pipelines': pipelines, 'forensics': forensics, 'compression': compression}
    product_nn.HOST_ONLY = old


def t64(a, grad=False):
    return torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=grad)


def chk(G, case, name, got, tol=TIGHT):
    return G.check(case, name, got.detach().numpy() if isinstance(got, torch.Tensor) else got, tol=tol, slack=0.0)


# ------------------------------------------------------------------------------------------------------------------ dJPEG
def test_oracle_djpeg_matches_executed_reference(G):
    m = G.meta['djpeg']
    rs = np.random.RandomState(m['seed'])
    x = rs.uniform(size=m['shape']).astype(np.float32)
    w = rs.normal(size=x.shape).astype(np.float32)
    for q, mode in m['cases']:
        case = 'djpeg_q{}_{}'.format(q, mode)
        xt = t64(x, True)
        y, X = R.djpeg(xt, R.jpeg_qtable(q, 0), R.jpeg_qtable(q, 1), mode)
        dx, = torch.autograd.grad((y * t64(w)).sum(), xt)
        chk(G, case, 'y', y), chk(G, case, 'X', X), chk(G, case, 'dx', dx)
    chk(G, 'djpeg_common_q80', 'y', R.jpeg_manipulation(t64(x), 80))
    for q, mode in m['trainable_cases']:          # DifferentiableJPEG(trainable=True): gradients w.r.t. the quantisation tables
        case = 'djpeg_trainable_q{}_{}'.format(q, mode)
        xt, ql, qc = t64(x, True), t64(R.jpeg_qtable(q, 0), True), t64(R.jpeg_qtable(q, 1), True)
        y, _ = R.djpeg(xt, ql, qc, mode)
        dx, dql, dqc = torch.autograd.grad((y * t64(w)).sum(), [xt, ql, qc])
        chk(G, case, 'y', y), chk(G, case, 'dx', dx), chk(G, case, 'dq_luma', dql), chk(G, case, 'dq_chroma', dqc)


# ------------------------------------------------------------------------------------------------------------------ manipulations
MANIP = {
    'sharpen_1': lambda t: R.manipulation_sharpen(t, 1, hsv=True), 'sharpen_0p4': lambda t: R.manipulation_sharpen(t, 0.4, hsv=True),
    'resample_50': lambda t: R.manipulation_resample(t, 50), 'resample_75': lambda t: R.manipulation_resample(t, 75),
    'resample_0p6': lambda t: R.manipulation_resample(t, 0.6),
    'gaussian_0p83': lambda t: R.manipulation_gaussian(t, 5, 0.83), 'gaussian_2p5': lambda t: R.manipulation_gaussian(t, 5, 2.5),
    'gaussian_k3': lambda t: R.manipulation_gaussian(t, 3, 1.0),
    'gamma_3': lambda t: R.manipulation_gamma(t, 3.0), 'gamma_0p7': lambda t: R.manipulation_gamma(t, 0.7),
    'median_3': lambda t: R.manipulation_median(t, 3), 'median_5': lambda t: R.manipulation_median(t, 5), 'median_4': lambda t: R.manipulation_median(t, 4),
    'soft_quantization': lambda t: R.soft_quantization(t), 'quantize_and_clip': lambda t: R.soft_quantization(t * 1.2 - 0.1).clamp(0, 1),
}


def manip_inputs(G):
    rs = np.random.RandomState(G.meta['manip']['seed'])
    x = rs.uniform(size=(2, 32, 32, 3)).astype(np.float32)
    xr = rs.uniform(size=(2, 16, 24, 3)).astype(np.float32)
    w = rs.normal(size=x.shape).astype(np.float32)
    wr = rs.normal(size=xr.shape).astype(np.float32)
    noise = rs.normal(size=x.shape).astype(np.float32)
    return x, xr, w, wr, noise


def test_oracle_manipulations_match_executed_reference(G):
    x, xr, w, wr, noise = manip_inputs(G)
    assert set(G.meta['manip']['ops']) == set(MANIP)
    for name, kind in G.meta['manip']['ops'].items():
        xx, ww = (x, w) if kind == 'sq' else (xr, wr)
        xt = t64(xx, True)
        y = MANIP[name](xt)
        s = (y * t64(ww)).sum()
        dx = torch.autograd.grad(s, xt)[0] if s.requires_grad else torch.zeros_like(xt)
        # 1e-6: the reference builds its filter constants as tf.float32 (float64 in the truth run), the oracle rounds them to float32
        chk(G, 'manip_' + name, 'y', y, 1e-6), chk(G, 'manip_' + name, 'dx', dx, 1e-6)
        if name.startswith('sharpen'):
            assert float(dx.abs().max()) == 0.0 and float(np.abs(G.get('manip_' + name, 'dx')).max()) == 0.0          # no gradient through HSV in TF 2.1
    xt = t64(x, True)
    y = R.manipulation_awgn(xt, 5.1 / 255, t64(noise))
    dx, = torch.autograd.grad((y * t64(w)).sum(), xt)
    chk(G, 'manip_awgn_5p1', 'y', y), chk(G, 'manip_awgn_5p1', 'dx', dx)


# ------------------------------------------------------------------------------------------------------------------ layers and losses
def test_oracle_layers_match_executed_reference(G):
    rs = np.random.RandomState(G.meta['layers']['seed'])
    x = rs.uniform(size=(2, 12, 14, 3)).astype(np.float32)
    w = rs.normal(size=x.shape).astype(np.float32)
    k = rs.normal(size=(5, 5, 3, 3)).astype(np.float32)
    z = (rs.normal(size=(2, 4, 4, 8)) * 3).astype(np.float32)
    wz = rs.normal(size=z.shape).astype(np.float32)
    # constrained convolution (models/layers.py:36-57)
    init = G.get('constrained_conv2d', 'init').reshape(5, 5, 3, 3)
    kt, xt = t64(init + 0.3 * k, True), t64(x, True)
    y = R.conv2d(R.tf_pad(xt, 2, 'SYMMETRIC'), M.constrained_filter(kt), padding='VALID')
    dx, dk = torch.autograd.grad((y * t64(w)).sum(), [xt, kt])
    chk(G, 'constrained_conv2d', 'y', y), chk(G, 'constrained_conv2d', 'dx', dx), chk(G, 'constrained_conv2d', 'dkernel', dk)
    # soft-codebook latent + entropy (models/layers.py:139-203, helpers/tf_helpers.py:290-333)
    for v, gamma, bpf in ((50, 25, 5), (0, 5, 4), (50, 25, 3)):
        case = 'discrete_latent_v{}_g{}_b{}'.format(v, gamma, bpf)
        zt, sc = t64(z, True), t64(np.float32(0.8), True)
        cb = M.dcn_codebook(bpf, torch.float64)
        assert np.array_equal(cb.numpy(), G.get(case, 'codebook'))
        lat, ent = R.discrete_latent(zt, sc, cb, v, gamma)
        dz, ds = torch.autograd.grad((lat * t64(wz)).sum() + 7.0 * ent.double(), [zt, sc])
        chk(G, case, 'latent', lat), chk(G, case, 'entropy', ent, 2e-7), chk(G, case, 'dz', dz, 1e-6), chk(G, case, 'dscale', ds, 1e-6)
    for mode in ('round', 'sin', 'soft', 'identity', 'harmonic'):
        zt = t64(z, True)
        y = R.quantization(zt, mode)
        dz = torch.autograd.grad((y * t64(wz)).sum(), zt)[0] if mode != 'round' else torch.zeros_like(zt)
        chk(G, 'quantization_' + mode, 'y', y), chk(G, 'quantization_' + mode, 'dz', dz)
    # image losses (helpers/tf_helpers.py:31-44)
    a = rs.uniform(size=(2, 48, 40, 3)).astype(np.float32)
    b = np.clip(a + 0.1 * rs.normal(size=a.shape), 0, 1).astype(np.float32)
    am = rs.uniform(size=(1, 192, 176, 3)).astype(np.float32)
    bm = np.clip(am + 0.05 * rs.normal(size=am.shape), 0, 1).astype(np.float32)
    for name, f, (p, q) in (('mse', R.mse, (a, b)), ('mae', R.mae, (a, b)), ('ssim_loss', R.ssim_loss, (a, b)), ('msssim_loss', R.msssim_loss, (am, bm))):
        pt = t64(p, True)
        val = f(pt, t64(q))
        dp, = torch.autograd.grad(val, pt)
        chk(G, 'loss_' + name, 'loss', val), chk(G, 'loss_' + name, 'da', dp)
    chk(G, 'tf_image_ssim', 'ssim', R.ssim_tf(a, b, 1.0))


# ------------------------------------------------------------------------------------------------------------------ models
def _adam_two_steps(P, loss_fn, lrs):
    opt = {'t': 0, 'm': {}, 'v': {}}
    losses = []
    names = list(P.keys())
    for lr in lrs:
        loss = loss_fn(P)
        g = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
        g = [torch.zeros_like(P[k]) if e is None else e for e, k in zip(g, names)]
        opt['t'] += 1
        with torch.no_grad():
            ms = [opt['m'].setdefault(k, torch.zeros_like(P[k])) for k in names]
            vs = [opt['v'].setdefault(k, torch.zeros_like(P[k])) for k in names]
            R.adam_keras_step([P[k] for k in names], g, ms, vs, opt['t'], lr)
        losses.append(float(loss))
    return losses


def nip_forward(meta):
    cls, kw = meta['cls'], meta['kw']
    if cls == 'UNet':
        return lambda P, x: M.unet_forward(P, x, n_steps=kw.get('n_steps', 5), activation=kw.get('activation', 'leaky_relu'))
    if cls == 'INet':
        return lambda P, x: M.inet_forward(P, x, kernel=kw.get('kernel', 5))
    if cls == 'DNet':
        return lambda P, x: M.dnet_forward(P, x, n_layers=kw.get('n_layers', 15))
    return lambda P, x: M.classic_isp_forward(P, x, n_cnn=len(kw.get('c_filters', ())))


def test_oracle_nip_models_match_executed_reference(G, host_models):
    cases = [c for c in G.meta if c.startswith('nip_')]
    assert len(cases) == 7
    for case in cases:
        meta = G.meta[case]
        kw = {k: tuple(v) if isinstance(v, list) else v for k, v in meta['kw'].items()}
        ps = meta['patch_size']
        pm = getattr(host_models['pipelines'], meta['cls'])(patch_size=ps, seed=1, **kw)
        specs = C.specs_of(pm)
        state = C.golden_state(specs, meta['seed'], ones_names=('conv2d_4/kernel',) if meta['cls'] == 'DNet' else ())
        rs = np.random.RandomState(meta['seed'])
        x = rs.uniform(size=(2, ps, ps, 4)).astype(np.float32)
        t = rs.uniform(size=(2, 2 * ps, 2 * ps, 3)).astype(np.float32)
        fwd = nip_forward(meta)
        P = M.to_params(state, torch.float64)
        train = [n for n, s, tr, _ in specs if tr]
        xt = t64(x, True)
        y = fwd(P, xt)
        loss = R.mse(y, t64(t))
        g = torch.autograd.grad(loss, [xt] + [P[n] for n in train])
        chk(G, case, 'y', y), chk(G, case, 'loss', loss), chk(G, case, 'dx', g[0])
        for n, e in zip(train, g[1:]):
            chk(G, case, 'grad/' + n, e)
        Pt = {n: P[n] for n in train}
        frozen = {n: P[n].detach() for n in P if n not in Pt}
        losses = _adam_two_steps(Pt, lambda Q: R.mse(fwd({**frozen, **Q}, t64(x)), t64(t)), [1e-3, 5e-4])
        chk(G, case, 'step_loss', np.array(losses))
        for n in train:
            chk(G, case, 'param2/' + n, Pt[n])


def test_oracle_fan_matches_executed_reference(G, host_models):
    for case in [c for c in G.meta if c.startswith('fan_')]:
        meta = G.meta[case]
        kw, ps = meta['kw'], meta['patch_size']
        pm = host_models['forensics'].FAN(n_classes=5, patch_size=ps, seed=1, **kw)
        specs = C.specs_of(pm)
        state = C.golden_state(specs, meta['seed'])
        rs = np.random.RandomState(meta['seed'])
        x = rs.uniform(size=(6, ps, ps, 3)).astype(np.float32)
        labels = rs.randint(0, 5, size=(6,))
        okw = dict(n_convolutions=kw.get('n_convolutions', 4), n_dense=kw.get('n_dense', 0), use_gap=kw.get('use_gap', True),
                   activation=kw.get('activation', 'leaky_relu'))
        P = M.to_params(state, torch.float64)
        names = list(P.keys())
        xt = t64(x, True)
        p = M.fan_forward(P, xt, **okw)
        loss = R.sparse_categorical_crossentropy(labels, p)
        g = torch.autograd.grad(loss, [xt] + [P[n] for n in names])
        chk(G, case, 'probs', p), chk(G, case, 'loss', loss), chk(G, case, 'dx', g[0])
        assert np.array_equal(p.detach().numpy().argmax(axis=1), G.get(case, 'decide').astype(np.int64))
        for n, e in zip(names, g[1:]):
            chk(G, case, 'grad/' + n, e)
        losses = _adam_two_steps(P, lambda Q: R.sparse_categorical_crossentropy(labels, M.fan_forward(Q, t64(x), **okw)), [1e-3, 5e-4])
        chk(G, case, 'step_loss', np.array(losses))
        for n in names:
            chk(G, case, 'param2/' + n, P[n])


def test_oracle_twitter_dcn_matches_executed_reference(G, host_models):
    for case in [c for c in G.meta if c.startswith('dcn_')]:
        meta = G.meta[case]
        kw, ps = meta['kw'], meta['patch_size']
        pm = host_models['compression'].TwitterDCN(patch_size=ps, seed=1, **kw)
        specs = C.specs_of(pm)
        state = C.golden_state(specs, meta['seed'])
        x = np.random.RandomState(meta['seed']).uniform(size=(2, ps, ps, 3)).astype(np.float32)
        bpf, ew, rnd = kw.get('latent_bpf', 5), float(kw.get('entropy_weight', 250)), kw.get('rounding', 'soft-codebook')
        P = M.to_params(state, torch.float64)
        names = list(P.keys())
        xt = t64(x, True)
        y, ent, q, z = M.twitter_dcn_forward(P, xt, bpf, rnd)
        loss = M.dcn_loss(xt, y, ent, ew)
        g = torch.autograd.grad(loss, [xt] + [P[n] for n in names])
        # the reference casts the soft latent and the entropy to float32 in the middle of the graph (models/layers.py:161, tf_helpers.py:331)
        chk(G, case, 'y', y, 1e-6), chk(G, case, 'entropy', ent, 1e-6), chk(G, case, 'loss', loss, 1e-6), chk(G, case, 'dx', g[0], 2e-6)
        chk(G, case, 'latent', q, 1e-6)
        assert np.array_equal(M.dcn_codebook(bpf).numpy(), G.get(case, 'codebook'))
        for n, e in zip(names, g[1:]):
            chk(G, case, 'grad/' + n, e, 2e-6)
        opt = {'t': 0, 'm': {}, 'v': {}}
        s1 = M.dcn_training_step(P, opt, t64(x), 1e-3, ew, bpf, rnd)[0]
        s2 = M.dcn_training_step(P, opt, t64(x), 5e-4, ew, bpf, rnd)[0]
        chk(G, case, 'step_loss', np.array([s1['loss'], s2['loss']]), 1e-6)
        chk(G, case, 'step_entropy', np.array([s1['entropy'], s2['entropy']]), 1e-6)
        for n in names:
            G.check(case, 'param2/' + n, P[n].detach().numpy(), tol=2e-6, outliers=0.01, loose=2.1e-3)


# ------------------------------------------------------------------------------------------------------------------ joint workflow
FLOW = 5e-6          # the manipulations' filter constants are tf.float32 in the reference (float64 in the truth run), float32-rounded in the oracle


def fchk(G, case, name, got, tol=FLOW):
    return chk(G, case, name, got, tol)


def flow_setup(G, host_models, case, n_classes, with_dcn=False):
    meta = G.meta[case]
    ps, B = meta['patch_size'], meta['batch']
    s_nip = C.specs_of(host_models['pipelines'].UNet(patch_size=ps, seed=1))
    s_fan = C.specs_of(host_models['forensics'].FAN(n_classes=n_classes, patch_size=ps, seed=1))
    st = [C.golden_state(s_nip, meta['seed']), C.golden_state(s_fan, meta['fan_seed'])]
    if with_dcn:
        st.append(C.golden_state(C.specs_of(host_models['compression'].TwitterDCN(patch_size=ps, seed=1)), meta['dcn_seed']))
    rs = np.random.RandomState(meta['seed'])
    x = rs.uniform(size=(B, ps, ps, 4)).astype(np.float32)
    t = rs.uniform(size=(B, 2 * ps, 2 * ps, 3)).astype(np.float32)
    return meta, st, x, t


def test_oracle_joint_step_matches_executed_reference(G, host_models):
    """ManipulationClassification.run_workflow + training_step x 2 (workflows/manipulation_classification.py:162-285), config 4 wiring."""
    case = 'flow_default'
    meta, (st_nip, st_fan), x, t = flow_setup(G, host_models, case, 5)
    Pn, Pf = M.to_params(st_nip, torch.float64), M.to_params(st_fan, torch.float64)
    with torch.no_grad():
        Y, c, Cc, probs = M.workflow_forward(Pn, Pf, t64(x))
    fchk(G, case, 'Y', Y), fchk(G, case, 'c', c), fchk(G, case, 'C', Cc), fchk(G, case, 'probs', probs)
    assert np.array_equal(M.batch_labels(meta['batch'], 5), G.get(case, 'labels').astype(np.int64))
    opt = {'t': 0, 'm': {}, 'v': {}}
    l1, grads = M.training_step(Pn, Pf, opt, t64(x), t64(t), lambda_nip=meta['lambda_nip'], lr=meta['lr'][0], train_nip=True)
    for k, g in grads.items():
        fchk(G, case, 'grad/' + k, g)
    l2, _ = M.training_step(Pn, Pf, opt, t64(x), t64(t), lambda_nip=meta['lambda_nip'], lr=meta['lr'][1], train_nip=True)
    for key in ('loss', 'ce', 'nip'):
        chk(G, case, key, np.array([l1[key], l2[key]]))
    for n, p in Pf.items():
        fchk(G, case, 'param2/fan/' + n, p)
    for n, p in Pn.items():
        fchk(G, case, 'param2/nip/' + n, p)


def test_oracle_joint_step_continuous_variant(G, host_models):
    """Same step with a continuous codec ('sin') and no hard-rounding manipulation."""
    case = 'flow_sin'
    meta, (st_nip, st_fan), x, t = flow_setup(G, host_models, case, 4)
    Pn, Pf = M.to_params(st_nip, torch.float64), M.to_params(st_fan, torch.float64)
    names = tuple(meta['manipulations'])

    def forward(Pn, Pf, xt):
        Y = M.unet_forward(Pn, xt)
        c = R.avg_pool(M.run_manipulations(Y, names), 2)
        Cc = R.djpeg(c, R.jpeg_qtable(50, 0), R.jpeg_qtable(50, 1), 'sin')[0]
        return Y, c, Cc, M.fan_forward(Pf, Cc)
    Y, c, Cc, probs = forward(Pn, Pf, t64(x))
    fchk(G, case, 'Y', Y), fchk(G, case, 'c', c), fchk(G, case, 'C', Cc), fchk(G, case, 'probs', probs)
    loss_ce = R.sparse_categorical_crossentropy(M.batch_labels(meta['batch'], 4), probs)
    loss_nip = R.mse(t64(t), Y)
    loss = loss_ce + meta['lambda_nip'] * loss_nip
    params = [('fan/' + k, v) for k, v in Pf.items()] + [('nip/' + k, v) for k, v in Pn.items()]
    g = torch.autograd.grad(loss, [p for _, p in params])
    for (k, _), e in zip(params, g):
        fchk(G, case, 'grad/' + k, e)
    assert abs(float(loss) - G.get(case, 'loss')[0]) <= 1e-9 * float(loss)


def test_oracle_fan_only_variant_with_bilinear_downsampling(G, host_models):
    """trainable = {fan}; manipulations resample:70, gaussian:1.5, gamma, median; 'bilinear' down-sampling; codec 'harmonic'."""
    case = 'flow_fan_only'
    meta, (st_nip, st_fan), x, t = flow_setup(G, host_models, case, 5)
    Pn, Pf = M.to_params(st_nip, torch.float64), M.to_params(st_fan, torch.float64)
    Y = M.unet_forward(Pn, t64(x))
    ops = [lambda v: R.manipulation_resample(v, 70.0), lambda v: R.manipulation_gaussian(v, 5, 1.5), lambda v: R.manipulation_gamma(v, 3),
           lambda v: R.manipulation_median(v, 3)]
    m = torch.cat([Y] + [f(Y) for f in ops], dim=0)
    c = R.resize_bilinear(m, m.shape[1] // 2, m.shape[1] // 2)
    Cc = R.djpeg(c, R.jpeg_qtable(50, 0), R.jpeg_qtable(50, 1), 'harmonic')[0]
    probs = M.fan_forward(Pf, Cc)
    fchk(G, case, 'Y', Y), fchk(G, case, 'c', c), fchk(G, case, 'C', Cc), fchk(G, case, 'probs', probs)
    loss = R.sparse_categorical_crossentropy(M.batch_labels(meta['batch'], 5), probs)
    g = torch.autograd.grad(loss, list(Pf.values()))
    for k, e in zip(Pf, g):
        fchk(G, case, 'grad/fan/' + k, e)
    assert abs(float(loss) - G.get(case, 'loss')[0]) <= 1e-9 * float(loss)
    assert not G.has(case, 'grad/nip/ec11/kernel')          # the ISP is not in the trainable set: the reference passes no gradient to it


def test_oracle_joint_step_with_learned_codec(G, host_models):
    """config 5: compression='dcn' (restored through the reference's codec.restore / tfmodel.restore), trainable {fan, nip, dcn}."""
    case = 'flow_dcn'
    meta, (st_nip, st_fan, st_dcn), x, t = flow_setup(G, host_models, case, 5, with_dcn=True)
    Pn, Pf, Pd = (M.to_params(s, torch.float64) for s in (st_nip, st_fan, st_dcn))
    with torch.no_grad():
        Y, c, Cc, probs, ent = M.workflow_forward(Pn, Pf, t64(x), P_dcn=Pd)
    fchk(G, case, 'Y', Y), fchk(G, case, 'c', c), chk(G, case, 'C', Cc, 1e-6), chk(G, case, 'probs', probs, 1e-6), chk(G, case, 'entropy', ent, 1e-6)
    opt = {'t': 0, 'm': {}, 'v': {}}
    l1, grads = M.training_step(Pn, Pf, opt, t64(x), t64(t), lambda_nip=0.1, lr=meta['lr'][0], train_nip=True, P_dcn=Pd, lambda_dcn=0.1,
                                train_dcn=True)
    for k, g in grads.items():
        chk(G, case, 'grad/' + k, g, 2e-6)
    l2, _ = M.training_step(Pn, Pf, opt, t64(x), t64(t), lambda_nip=0.1, lr=meta['lr'][1], train_nip=True, P_dcn=Pd, lambda_dcn=0.1,
                            train_dcn=True)
    for key in ('loss', 'ce', 'nip', 'dcn'):
        chk(G, case, key, np.array([l1[key], l2[key]]), 1e-6)
