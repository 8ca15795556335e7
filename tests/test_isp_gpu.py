"""GPU parity of the remaining camera ISPs (INet, DNet, ClassicISP) against the CPU oracle restatement of
models/pipelines.py:233-350,415-539 and models/layers.py:206-258: forward, training-step loss, every gradient."""
import numpy as np
import pytest
import torch

from conftest import assert_parity
from oracle import ref_models as M
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def _check(model, fwd, x, t, tol_y=1e-5, tol_g=5e-5, lr=1e-4):
    state = model._store.state_dict()
    names = [p.name for p in model._store.trainable]
    y = model.process(x).numpy()
    out = {}
    for dt in (torch.float64, torch.float32):
        P = M.to_params(state, dt)
        yt = fwd(P, torch.tensor(x, dtype=dt))
        loss = R.mse(yt, torch.tensor(t, dtype=dt))
        g = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
        out[dt] = (yt.detach().numpy(), float(loss), {k: (np.zeros(P[k].shape) if v is None else v.numpy()) for k, v in zip(names, g)})
    assert y.shape == out[torch.float64][0].shape
    assert_parity(y, out[torch.float64][0], out[torch.float32][0], tol=tol_y, what='y')
    loss = model.training_step(x, t, learning_rate=lr)
    assert abs(float(loss.numpy()) - out[torch.float64][1]) < 1e-4 * out[torch.float64][1]
    for p in model._store.trainable:
        ref64, ref32 = out[torch.float64][2][p.name], out[torch.float32][2][p.name]
        assert_parity(p.grad.cpu().numpy().reshape(ref64.shape), ref64, ref32, tol=tol_g, slack=6.0, what='grad ' + p.name)
    # Adam moved every parameter with a non-negligible gradient by ~lr
    new = model._store.state_dict()
    for k in names:
        g = np.abs(out[torch.float64][2][k])
        if g.max() > 1e-6:
            assert 0 < np.abs(new[k] - state[k]).max() <= 1.01 * lr, k
    return out


@pytest.mark.parametrize('kw', [dict(), dict(random_init=True, kernel=3, trainable_upsampling=True, cfa_pattern='rggb')])
def test_inet(kw):
    from neural_imaging_b200.models import pipelines
    rs = np.random.RandomState(3)
    model = pipelines.INet(patch_size=16, seed=4, **kw)
    k = kw.get('kernel', 5)
    assert model.count_parameters() == (48 if kw.get('trainable_upsampling') else 0) + 9 * k * k + 9 + 36 + 12 + 36 + 3
    assert model.model_code == ('INet_rggbTR_3x3' if kw else 'INet_gbrg_5x5')
    x = rs.uniform(size=(2, 16, 16, 4)).astype(np.float32)
    t = rs.uniform(size=(2, 32, 32, 3)).astype(np.float32)
    _check(model, lambda P, xt: M.inet_forward(P, xt, k), x, t)


def test_inet_default_weights_act_like_an_isp():
    """Default INet = bilinear demosaicing + colour matrix + tone curve: a constant grey RAW patch develops to a constant image."""
    from neural_imaging_b200.models import pipelines
    model = pipelines.INet(patch_size=8)
    y = model.process(np.full((1, 8, 8, 4), 0.25, np.float32)).numpy()
    assert y.shape == (1, 16, 16, 3) and np.ptp(y[0, :, :, 1]) < 1e-5 and 0.0 < y.mean() < 1.0


@pytest.mark.parametrize('kw', [dict(n_layers=3, n_features=16), dict(n_layers=2, kernel=5, n_features=8)])
def test_dnet(kw):
    from neural_imaging_b200.models import pipelines
    rs = np.random.RandomState(5)
    model = pipelines.DNet(patch_size=24, seed=6, **kw)
    nl, k, nf = kw['n_layers'], kw.get('kernel', 3), kw['n_features']
    x = rs.uniform(size=(2, 24, 24, 4)).astype(np.float32)
    t = rs.uniform(size=(2, 48, 48, 3)).astype(np.float32)
    # the all-ones 1x1 projection of the reference saturates the clipped output at initialisation; scale it into [0, 1]
    state = model._store.state_dict()
    state['conv2d_%d/kernel' % (nl + 1)] = state['conv2d_%d/kernel' % (nl + 1)] * np.float32(0.6 / nf)
    model._store.load_state_dict(state)
    _check(model, lambda P, xt: M.dnet_forward(P, xt, nl, k), x, t)
    assert model.model_code == 'DNet_{k}x{k}_{l}x{f}f'.format(k=k, l=nl, f=nf)


def test_dnet_default_shape_and_parameter_count():
    from neural_imaging_b200.models import pipelines
    model = pipelines.DNet(patch_size=32, seed=1)
    # 15 layers: 4->64, 13 x 64->64, 64->12 (3x3 + bias), project 6->64 (3x3 + bias), final 64->3 (1x1, no bias)
    want = (36 * 64 + 64) + 13 * (576 * 64 + 64) + (576 * 12 + 12) + (54 * 64 + 64) + 192
    assert model.count_parameters() == want
    y = model.process(np.random.RandomState(0).uniform(size=(1, 32, 32, 4)).astype(np.float32))
    assert tuple(y.shape) == (1, 64, 64, 3)


@pytest.mark.parametrize('kw,n_cnn', [(dict(c_filters=(8, 8), kernel=3), 2), (dict(c_filters=(8,), residual=False), 1)])
def test_classic_isp_trainable(kw, n_cnn):
    from neural_imaging_b200.models import pipelines
    rs = np.random.RandomState(8)
    srgb = np.array([[1.6, -0.4, -0.2], [-0.2, 1.5, -0.3], [0.05, -0.45, 1.4]], np.float32)
    model = pipelines.ClassicISP(patch_size=16, seed=9, **kw)
    model.set_srgb_conversion(srgb)
    model.set_cfa_pattern('RGGB')
    x = rs.uniform(size=(2, 16, 16, 4)).astype(np.float32)
    t = rs.uniform(size=(2, 32, 32, 3)).astype(np.float32)
    k, residual = kw.get('kernel', 5), kw.get('residual', True)
    # the pow(., 1/2.2) tail has an unbounded derivative near the 1/255 clip: float32 drift is larger than for the other ISPs
    _check(model, lambda P, xt: M.classic_isp_forward(P, xt, k, n_cnn, residual), x, t, tol_y=2e-5, tol_g=1e-4)
    assert model.model_code == 'ClassicISP_rggb_{k}x{k}_{fs}-3{r}'.format(k=k, fs='-'.join(str(c) for c in kw['c_filters']), r='R' if residual else '')


def test_classic_isp_default_is_bilinear_plus_gamma():
    from neural_imaging_b200.models import pipelines
    model = pipelines.ClassicISP(patch_size=16)
    assert model.count_parameters() == 1                     # alpha only (the CNN branch is never built, models/layers.py:244-254)
    x = np.random.RandomState(1).uniform(size=(16, 16, 4)).astype(np.float32)       # 3-D input gets a batch axis
    y = model.process(x, cfa_pattern='GBRG').numpy()
    P = M.to_params(model._store.state_dict(), torch.float64, False)
    ref = M.classic_isp_forward(P, torch.tensor(x[None], dtype=torch.float64)).numpy()
    assert y.shape == (1, 32, 32, 3) and np.abs(y - ref).max() < 2e-6
    assert 'ClassicISP' in pipelines.supported_models and 'INet' in pipelines.supported_models and 'DNet' in pipelines.supported_models


def test_workflow_accepts_every_isp():
    from neural_imaging_b200.workflows.manipulation_classification import ManipulationClassification
    rs = np.random.RandomState(2)
    x = rs.uniform(size=(2, 32, 32, 4)).astype(np.float32)
    t = rs.uniform(size=(2, 64, 64, 3)).astype(np.float32)
    for nip in ('INet', 'DNet', 'ClassicISP'):
        flow = ManipulationClassification(nip, manipulations=['gaussian', 'gamma'], trainable={'nip'}, raw_patch_size=32, seed=3,
                                          fan_args=dict(n_filters=8, n_convolutions=2))
        loss, parts = flow.training_step(x, t, lambda_nip=0.1, learning_rate=1e-4)
        assert np.isfinite(float(loss.numpy())) and np.isfinite(float(parts['nip'].numpy())), nip
