"""GPU parity of the models (UNet, FAN) and of the joint training step against the CPU oracle restatement of
models/pipelines.py:169-230, models/forensics.py:29-125 and workflows/manipulation_classification.py:260-285."""
import os

import numpy as np
import pytest
import torch

from conftest import assert_parity, rel_err
from oracle import ref_models as M
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def _grads(store):
    return {p.name: p.grad.detach().cpu().numpy().copy() for p in store.trainable}


def test_unet_forward_backward():
    from neural_imaging_b200.models import pipelines
    rs = np.random.RandomState(1234)
    model = pipelines.UNet(patch_size=32, seed=1)
    assert model.count_parameters() == 7763820
    x = rs.uniform(size=(2, 32, 32, 4)).astype(np.float32)
    t = rs.uniform(size=(2, 64, 64, 3)).astype(np.float32)
    y = model.process(x).numpy()
    state = model._store.state_dict()
    out = {}
    for dt in (torch.float64, torch.float32):
        P = M.to_params(state, dt)
        yt = M.unet_forward(P, torch.tensor(x, dtype=dt))
        loss = R.mse(yt, torch.tensor(t, dtype=dt))
        g = torch.autograd.grad(loss, list(P.values()))
        out[dt] = (yt.detach().numpy(), float(loss), {k: v.numpy() for k, v in zip(P.keys(), g)})
    assert tuple(y.shape) == (2, 64, 64, 3)
    assert_parity(y, out[torch.float64][0], out[torch.float32][0], tol=1e-5, what='UNet y')
    # one training step: loss value + gradients (before Adam touches the weights, read them from the flat buffer)
    loss = model.training_step(x, t, learning_rate=1e-4)
    assert abs(float(loss.numpy()) - out[torch.float64][1]) < 1e-4 * out[torch.float64][1]
    g = _grads(model._store)
    for name, ref in out[torch.float64][2].items():
        # whole-network gradients through 23 tcgen05 (3xTF32) layers: measured <= 2.1e-5 scale-relative
        assert_parity(g[name], ref, out[torch.float32][2][name], tol=5e-5, slack=6.0, what='UNet grad ' + name)
    # Keras-Adam update of every parameter
    P = M.to_params(state, torch.float64, requires_grad=False)
    ms = [torch.zeros_like(p) for p in P.values()]
    vs = [torch.zeros_like(p) for p in P.values()]
    R.adam_keras_step(list(P.values()), [torch.tensor(out[torch.float64][2][k]) for k in P], ms, vs, 1, 1e-4)
    new = model._store.state_dict()
    for k, p in P.items():
        # Adam's first step is lr * g / (|g| + eps): a parameter whose gradient is at the rounding-noise level can move by
        # up to 2*lr in the "wrong" direction in ANY float32 implementation, so bound the bulk and the extreme separately
        delta = np.abs(new[k] - p.numpy())
        assert delta.max() <= 2.02e-4 and np.mean(delta > 2e-6) < 2e-3, k


@pytest.mark.parametrize('kw', [dict(), dict(n_filters=8, n_convolutions=2, kernel=3, n_dense=2, use_gap=False, activation='relu')])
def test_fan_forward_backward(kw):
    from neural_imaging_b200.models import forensics
    from neural_imaging_b200.tensor import as_device
    rs = np.random.RandomState(99)
    ps = 32
    fan = forensics.FAN(n_classes=5, patch_size=ps, seed=3, **kw)
    if not kw:
        assert fan.count_parameters() == 1145382         # SURVEY Appendix A (5 classes)
    m = 6
    x = rs.uniform(size=(m, ps, ps, 3)).astype(np.float32)
    labels = rs.randint(0, 5, size=(m,))
    probs = fan.process(x).numpy()
    state = fan._store.state_dict()
    okw = dict(n_convolutions=kw.get('n_convolutions', 4), n_dense=kw.get('n_dense', 0), use_gap=kw.get('use_gap', True),
               activation=kw.get('activation', 'leaky_relu'))
    out = {}
    for dt in (torch.float64, torch.float32):
        P = M.to_params(state, dt)
        xt = torch.tensor(x, dtype=dt, requires_grad=True)
        pt = M.fan_forward(P, xt, **okw)
        loss = R.sparse_categorical_crossentropy(labels, pt)
        g = torch.autograd.grad(loss, [xt] + list(P.values()))
        out[dt] = (pt.detach().numpy(), float(loss), g[0].numpy(), {k: v.numpy() for k, v in zip(P.keys(), g[1:])})
    assert_parity(probs, out[torch.float64][0], out[torch.float32][0], tol=1e-5, what='FAN probs')
    assert np.array_equal(fan.process_and_decide(x), probs.argmax(axis=1))
    assert abs(float(fan.loss(labels, probs).numpy()) - out[torch.float64][1]) < 2e-5 * max(1, out[torch.float64][1])
    pr, loss, dlogits = fan.forward_loss(as_device(x), as_device(labels.astype(np.int32), torch.int32))
    dx = fan.backward(dlogits, need_dx=True)
    assert abs(float(loss.item()) / m - out[torch.float64][1]) < 1e-5 * max(1, out[torch.float64][1])
    assert_parity(dx.cpu().numpy(), out[torch.float64][2], out[torch.float32][2], tol=5e-5, slack=6.0, what='FAN dx')
    g = _grads(fan._store)
    for name, ref in out[torch.float64][3].items():
        assert_parity(g[name], ref, out[torch.float32][3][name], tol=5e-5, slack=6.0, what='FAN grad ' + name)


@pytest.mark.parametrize('ps', [32, 44])
def test_fan_fused_first_block_equals_the_kernel_pair(ps):
    """conv 5x5 3->32 + bias + LeakyReLU + MaxPool 2x2 in one kernel (ni_conv2d_pool2_fwd, code byte instead of the full-resolution
    activation; backward ni_maxpool2_code_bwd_bias) against the conv -> ni_maxpool2_fwd / ni_maxpool2_act_bwd_bias pair: same
    accumulation order, so the pooled activations and the routed gradients are bit-equal; the bias gradient is a different
    reduction tree (tight tolerance)."""
    from neural_imaging_b200.models import forensics
    from neural_imaging_b200.tensor import as_device
    rs = np.random.RandomState(5)
    m = 5
    x = as_device(rs.uniform(size=(m, ps, ps, 3)).astype(np.float32))
    labels = as_device(rs.randint(0, 5, size=(m,)).astype(np.int32), torch.int32)
    res = []
    for fuse in (True, False):
        fan = forensics.FAN(n_classes=5, patch_size=ps, seed=3)
        fan._fuse_pool = fuse
        probs, loss, dlogits = fan.forward_loss(x, labels)
        acts = fan._saved[0]
        assert (acts['c0'] is None) == fuse and ('code0' in acts) == fuse
        dx = fan.backward(dlogits, need_dx=True)
        res.append((probs.cpu().numpy().copy(), acts['p0'].cpu().numpy().copy(), dx.cpu().numpy().copy(), _grads(fan._store)))
    (pa, p0a, dxa, ga), (pb, p0b, dxb, gb) = res
    assert np.array_equal(p0a, p0b) and np.array_equal(pa, pb)
    assert np.array_equal(dxa, dxb)
    for k in ga:            # (filter gradients: split-K atomics, bias gradients: another reduction tree -> not bit-reproducible)
        scale = float(np.abs(gb[k]).max()) + 1e-12
        assert float(np.abs(ga[k] - gb[k]).max()) <= 2e-5 * scale, k


@pytest.mark.parametrize('train_nip', [True, False])
def test_joint_training_step(train_nip):
    """ManipulationClassification.training_step: losses, every gradient and the Adam update vs the oracle."""
    from neural_imaging_b200.workflows.manipulation_classification import ManipulationClassification
    rs = np.random.RandomState(1234)
    B, ps = 2, 32
    flow = ManipulationClassification('UNet', trainable={'nip'} if train_nip else None, raw_patch_size=ps, seed=5)
    assert flow.n_classes == 5 and flow._forensics_classes[0] == 'native'
    x = rs.uniform(size=(B, ps, ps, 4)).astype(np.float32)
    t = rs.uniform(size=(B, 2 * ps, 2 * ps, 3)).astype(np.float32)
    s_nip, s_fan = flow.nip._store.state_dict(), flow.fan._store.state_dict()
    # inference composition first (run_workflow returns Y, c, C, entropy, probs)
    Y, c, C, ent, probs = flow.run_workflow(x)
    assert np.isnan(ent) and tuple(C.shape) == (5 * B, ps, ps, 3) and tuple(probs.shape) == (5 * B, 5)
    res = {}
    for dt in (torch.float64, torch.float32):
        Pn, Pf = M.to_params(s_nip, dt), M.to_params(s_fan, dt)
        opt = {'t': 0, 'm': {}, 'v': {}}
        losses, grads = M.training_step(Pn, Pf, opt, torch.tensor(x, dtype=dt), torch.tensor(t, dtype=dt), lambda_nip=0.1, lr=1e-4,
                                        train_nip=train_nip)
        with torch.no_grad():
            Yt, ct, Ct, pt = M.workflow_forward(M.to_params(s_nip, dt, False), M.to_params(s_fan, dt, False), torch.tensor(x, dtype=dt))
        res[dt] = (losses, grads, Pn, Pf, Yt.numpy(), ct.numpy(), Ct.numpy(), pt.numpy())
    r64, r32 = res[torch.float64], res[torch.float32]
    assert_parity(Y.numpy(), r64[4], r32[4], tol=1e-5, what='Y')
    # rounding inside the two dJPEG stages: compare the bulk, allow isolated 8x8 blocks to flip a coefficient
    for got, k in ((c.numpy(), 5), (C.numpy(), 6)):
        assert np.mean(np.abs(got - r64[k]) > 1e-4) < 2e-3
    loss, parts = flow.training_step(x, t, lambda_nip=0.1, learning_rate=1e-4)
    assert abs(float(parts['ce'].numpy()) - r64[0]['ce']) < 2e-3 * r64[0]['ce']
    assert abs(float(parts['nip'].numpy()) - r64[0]['nip']) < 1e-4 * r64[0]['nip']
    assert np.isnan(parts['dcn'])
    exp_loss = r64[0]['ce'] + (0.1 * r64[0]['nip'] if train_nip else 0)
    assert abs(float(loss.numpy()) - exp_loss) < 2e-3 * exp_loss
    gf = _grads(flow.fan._store)
    for name, g in gf.items():
        ref64, ref32 = r64[1]['fan/' + name].numpy(), r32[1]['fan/' + name].numpy()
        e = rel_err(g, ref64)
        assert e < max(5e-3, 4 * rel_err(ref32, ref64)), 'fan grad {} err {}'.format(name, e)
    if train_nip:
        gn = _grads(flow.nip._store)
        for name, g in gn.items():
            ref64, ref32 = r64[1]['nip/' + name].numpy(), r32[1]['nip/' + name].numpy()
            e = rel_err(g, ref64)
            assert e < max(2e-2, 4 * rel_err(ref32, ref64)), 'nip grad {} err {}'.format(name, e)
    else:
        assert float(np.abs(flow.nip._store.gflat.cpu().numpy()).max()) == 0.0
        assert np.array_equal(flow.nip._store.state_dict()['ec11/kernel'], s_nip['ec11/kernel'])
    # parameters moved by (at most) lr in the Adam direction
    new_fan = flow.fan._store.state_dict()
    moved = np.abs(new_fan['conv2d_0/kernel'] - s_fan['conv2d_0/kernel'])
    assert 0 < moved.max() <= 1.01e-4


def test_joint_training_step_with_ssim_loss():
    """loss_metric='SSIM' in the workflow (workflows/manipulation_classification.py:63-68 -> helpers/tf_helpers.py:39-40): the NIP part of
    the loss and the gradients that reach the ISP through BOTH the structural loss and the manipulation / codec / FAN branch."""
    from neural_imaging_b200.workflows.manipulation_classification import ManipulationClassification
    rs = np.random.RandomState(4321)
    B, ps = 2, 32
    flow = ManipulationClassification('UNet', trainable={'nip'}, raw_patch_size=ps, loss_metric='SSIM', seed=6)
    x = rs.uniform(size=(B, ps, ps, 4)).astype(np.float32)
    t = rs.uniform(size=(B, 2 * ps, 2 * ps, 3)).astype(np.float32)
    s_nip, s_fan = flow.nip._store.state_dict(), flow.fan._store.state_dict()
    res = {}
    for dt in (torch.float64, torch.float32):
        Pn, Pf = M.to_params(s_nip, dt), M.to_params(s_fan, dt)
        res[dt] = M.training_step(Pn, Pf, {'t': 0, 'm': {}, 'v': {}}, torch.tensor(x, dtype=dt), torch.tensor(t, dtype=dt), lambda_nip=0.5,
                                  lr=1e-4, train_nip=True, nip_loss=R.ssim_loss)
    (l64, g64), (l32, g32) = res[torch.float64], res[torch.float32]
    loss, parts = flow.training_step(x, t, lambda_nip=0.5, learning_rate=1e-4)
    assert abs(float(parts['nip'].numpy()) - l64['nip']) < 1e-4 * l64['nip']
    assert abs(float(loss.numpy()) - l64['loss']) < 2e-3 * l64['loss']
    for name, g in _grads(flow.nip._store).items():
        ref64, ref32 = g64['nip/' + name].numpy(), g32['nip/' + name].numpy()
        e = rel_err(g, ref64)
        assert e < max(2e-2, 4 * rel_err(ref32, ref64)), 'nip grad {} err {}'.format(name, e)
    with pytest.raises(ValueError):
        ManipulationClassification('UNet', raw_patch_size=ps, loss_metric='MS-SSIM')       # the reference accepts L2 / L1 / SSIM here (:63)


def test_onet_rgb_training_and_api_surface():
    from neural_imaging_b200.workflows.manipulation_classification import ManipulationClassification
    rs = np.random.RandomState(7)
    flow = ManipulationClassification('ONet', manipulations=['gaussian:1', 'jpeg:85', 'awgn', 'median'],
                                      distribution={'downsampling': 'none', 'compression': 'jpeg',
                                                    'compression_params': {'quality': (50, 90), 'codec': 'soft'}},
                                      trainable={'nip'}, raw_patch_size=16, seed=2)
    assert flow._forensics_classes == ['native', 'gaussian:1.0', 'jpeg:85.0', 'awgn:5.1', 'median:3']
    y = rs.uniform(size=(3, 32, 32, 3)).astype(np.float32)
    losses = []
    for _ in range(3):
        loss, parts = flow.training_step(y, y, lambda_nip=0.1, augment=True, learning_rate=1e-3)
        losses.append(float(loss.numpy()))
    assert all(np.isfinite(losses))
    assert flow.run_workflow_to_decisions(y).shape == (15,)
    assert flow.run_rgb_to_probabilities(y).shape == (15, 5)
    assert 'FAN' in flow.summary() and 'Manipulations' in flow.details()
    with pytest.raises(ValueError):
        ManipulationClassification('UNet', manipulations=['bogus'], raw_patch_size=32)
    with pytest.raises(ValueError):
        ManipulationClassification('NoSuchNet', raw_patch_size=32)
    with pytest.raises(ValueError):
        ManipulationClassification('UNet', raw_patch_size=8)


def test_cuda_graph_step_matches_eager():
    """The captured-graph replay of the joint step (ManipulationClassification.enable_cuda_graph) must train exactly like
    the kernel-by-kernel path: same losses and same parameters after several steps with a changing learning rate."""
    import torch
    from neural_imaging_b200.workflows.manipulation_classification import ManipulationClassification
    rs = np.random.RandomState(7)
    xs = [rs.uniform(size=(2, 32, 32, 4)).astype(np.float32) for _ in range(4)]
    ys = [rs.uniform(size=(2, 64, 64, 3)).astype(np.float32) for _ in range(4)]
    lrs = [1e-3, 1e-3, 5e-4, 2e-4]
    out = {}
    for mode in ('eager', 'graph'):
        flow = ManipulationClassification('UNet', trainable={'nip'}, raw_patch_size=32, seed=1234)
        if mode == 'graph':
            flow.enable_cuda_graph()
        losses = []
        for x, y, lr in zip(xs, ys, lrs):
            loss, parts = flow.training_step(x, y, lambda_nip=0.1, learning_rate=lr)
            losses.append((float(loss.numpy()), float(parts['ce'].numpy()), float(parts['nip'].numpy())))
        if mode == 'graph':
            assert flow.graph_launches_per_step > 100     # steps 3 and 4 were replays of the captured graphs
        out[mode] = (np.array(losses), flow.fan._store.flat.cpu().numpy().copy(), flow.nip._store.flat.cpu().numpy().copy())
    # the first steps agree to float32 rounding; later ones inherit the run-to-run differences of the atomically summed weight gradients,
    # amplified by the hard rounding of the 'soft' codec (a flipped coefficient changes a whole 8x8 block)
    np.testing.assert_allclose(out['graph'][0][:2], out['eager'][0][:2], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(out['graph'][0], out['eager'][0], rtol=1e-3, atol=1e-6)
    # weight gradients are summed with atomics (order varies run to run): compare the parameters with a small tolerance
    # (Adam moves a weight by ~lr per step whatever the size of its gradient: where a gradient sits at the rounding-noise level the two
    # runs may step in opposite directions, so bound the bulk tightly and the extremes by 2 x lr x steps)
    for a, b in ((out['graph'][1], out['eager'][1]), (out['graph'][2], out['eager'][2])):
        d = np.abs(a - b)
        assert np.mean(d > 2e-5 * max(1.0, float(np.max(np.abs(b))))) < 1e-3 and d.max() <= 2 * sum(lrs), (float(d.max()), float(np.mean(d > 2e-5)))


class _ToyData:
    """Stand-in for reference helpers/dataset.py Dataset (RAW + RGB patches), deterministic."""
    _loaded_data = 'xy'

    def __init__(self, n_train=8, n_valid=4, raw=16, seed=3):
        rs = np.random.RandomState(seed)
        self.count_training, self.count_validation = n_train, n_valid
        self._x = rs.uniform(size=(n_train + n_valid, raw, raw, 4)).astype(np.float32)
        self._y = rs.uniform(size=(n_train + n_valid, 2 * raw, 2 * raw, 3)).astype(np.float32)

    def is_raw_and_rgb(self):
        return True

    def summary(self):
        return 'toy data ({} + {})'.format(self.count_training, self.count_validation)

    def next_training_batch(self, batch_id, batch_size, rgb_patch_size):
        i = (batch_id * batch_size) % self.count_training
        return self._x[i:i + batch_size], self._y[i:i + batch_size]

    def next_validation_batch(self, batch_id, batch_size):
        i = self.count_training + batch_id * batch_size
        return self._x[i:i + batch_size], self._y[i:i + batch_size]


def test_validate_nip_grouped_equals_per_image():
    """validate_nip develops the validation set in groups of NIP_GROUP with device-side reductions: same per-image ssim / psnr / loss as
    developing the images one by one on the host (reference training/validation.py:112-131)."""
    from neural_imaging_b200.helpers import metrics
    from neural_imaging_b200.models import pipelines
    from neural_imaging_b200.training import validation
    data = _ToyData(n_train=4, n_valid=validation.NIP_GROUP + 3, raw=16)
    model = pipelines.UNet(patch_size=16, seed=5)
    for loss_type in ('L2', 'L1'):
        ssims, psnrs, losss = validation.validate_nip(model, data, loss_type=loss_type)
        assert len(ssims) == len(psnrs) == len(losss) == data.count_validation
        for b in range(data.count_validation):
            x, y = data.next_validation_batch(b, 1)
            dev = model.process(x).numpy().clip(0, 1).squeeze()
            ref = np.asarray(y).squeeze()
            mse = float(np.mean(np.power(ref.astype(np.float64) - dev, 2.0)))
            assert abs(psnrs[b] - 10.0 * np.log10(1.0 / mse)) < 1e-4
            assert abs(ssims[b] - metrics.ssim(ref, dev)) < 1e-5
            want = mse if loss_type == 'L2' else float(np.mean(np.abs(ref.astype(np.float64) - dev)))
            assert abs(losss[b] - want) < 1e-6 * max(1.0, want)


def test_training_loop_api(tmp_path):
    """training.manipulation.train_manipulation_nip (reference training/manipulation.py:36): runs epochs through
    flow.training_step, validates with validate_fan (confusion matrix rows sum to 1), snapshots the models."""
    from neural_imaging_b200.training import manipulation, validation
    from neural_imaging_b200.workflows.manipulation_classification import ManipulationClassification
    flow = ManipulationClassification('UNet', trainable={'nip'}, raw_patch_size=16, seed=1234)
    data = _ToyData()
    spec = {'camera_name': 'toy', 'use_pretrained_nip': False, 'patch_size': 16, 'batch_size': 4, 'n_epochs': 3, 'validation_schedule': 2,
            'learning_rate': 1e-3, 'lambda_nip': 0.1}
    out = manipulation.train_manipulation_nip(flow, spec, data, {'root': str(tmp_path)}, overwrite=True)
    assert out.endswith('models') and str(tmp_path) in out
    # validation at epochs 0 and 2 plus the closing round the reference always runs (training/manipulation.py:300-303)
    assert len(flow.fan.performance['loss']['training']) == 3 and len(flow.fan.performance['accuracy']['validation']) == 3
    assert os.path.isfile(os.path.join(out, 'fan', 'fan.npz')) or os.path.isdir(os.path.join(out, 'fan'))
    conf = np.array(flow.fan.performance['confusion'])
    assert conf.shape == (5, 5) and np.allclose(conf.sum(axis=1), 1.0)       # reference normalisation: per-class rates (:200-202)
    acc, conf2, labels = validation.validate_fan(flow, data, get_labels=True)
    assert len(labels) == 5 * 4 and 0.0 <= acc <= 1.0
    with pytest.raises(RuntimeError):          # 'camera_name' is a required key (reference :81-86)
        manipulation.train_manipulation_nip(flow, {k: v for k, v in spec.items() if k != 'camera_name'}, data, {'root': str(tmp_path)})


def test_training_loop_with_the_real_dataset_class(tmp_path):
    """helpers.dataset.Dataset (integer images, patch sampling) through train_manipulation_nip and validate_fan: the loops read
    count_training / count_validation / rgb_patch_size from it (reference helpers/dataset.py:160-185), and validate_fan's grouped device
    pass must give the numbers of the reference's batch-of-10 host loop."""
    from neural_imaging_b200.helpers.dataset import Dataset
    from neural_imaging_b200.training import manipulation, validation
    from neural_imaging_b200.workflows.manipulation_classification import ManipulationClassification
    rs = np.random.RandomState(11)
    n, H = 6, 96
    y = rs.randint(0, 256, size=(n, H, H, 3)).astype(np.uint8)
    x = rs.randint(0, 65536, size=(n, H // 2, H // 2, 4)).astype(np.uint16)
    vy = rs.randint(0, 256, size=(23, 32, 32, 3)).astype(np.uint8)
    vx = rs.randint(0, 65536, size=(23, 16, 16, 4)).astype(np.uint16)
    data = Dataset.from_arrays(x=x, y=y, val_x=vx, val_y=vy)
    assert (data.count_training, data.count_validation, data.rgb_patch_size) == (6, 23, 32)
    flow = ManipulationClassification('UNet', trainable={'nip'}, raw_patch_size=16, seed=4)
    spec = {'camera_name': 'toy', 'use_pretrained_nip': False, 'patch_size': 16, 'batch_size': 3, 'n_epochs': 2, 'validation_schedule': 5,
            'learning_rate': 1e-3, 'lambda_nip': 0.1}
    np.random.seed(3)
    out = manipulation.train_manipulation_nip(flow, spec, data, {'root': str(tmp_path)}, overwrite=True)
    assert len(flow.fan.performance['accuracy']['validation']) == 2          # epoch 0 + the closing round (weights of the last epoch are saved)
    assert os.path.isdir(os.path.join(out, flow.nip.scoped_name)) and os.path.isfile(os.path.join(os.path.dirname(out), 'training.json'))
    # grouped device pass == the reference's loop (batches of 10, 23 // 10 = 2 of them, decisions read batch by batch)
    acc, conf, labels = validation.validate_fan(flow, data, get_labels=True)
    ref_conf, ref_labels, accs = np.zeros((5, 5)), [], []
    for b in range(2):
        bx = data.next_validation_batch(b, 10)[0]
        pred = np.asarray(flow.run_workflow_to_decisions(bx))
        lab = flow._batch_labels(10)
        np.add.at(ref_conf, (lab, pred), 1)
        ref_labels += list(pred)
        accs.append(np.mean(pred == lab))
    assert labels == [int(v) for v in ref_labels]
    assert np.array_equal(conf, ref_conf / 20) and abs(acc - np.mean(accs)) < 1e-12
    old = validation.GROUP_IMAGES
    try:
        validation.GROUP_IMAGES = 10          # one batch per pass: same result whatever the grouping
        acc1, conf1 = validation.validate_fan(flow, data)
    finally:
        validation.GROUP_IMAGES = old
    assert acc1 == acc and np.array_equal(conf1, conf)


def test_fused_pooled_stack_inside_the_training_step():
    """The joint step with the manipulation stack + average pooling fused (default) against the operator-by-operator path
    (flow._fuse_pool = False): same losses, same gradients."""
    from neural_imaging_b200.workflows.manipulation_classification import ManipulationClassification
    rs = np.random.RandomState(77)
    x = rs.uniform(size=(2, 32, 32, 4)).astype(np.float32)
    t = rs.uniform(size=(2, 64, 64, 3)).astype(np.float32)
    res = {}
    for fused in (True, False):
        flow = ManipulationClassification('UNet', trainable={'nip'}, raw_patch_size=32, seed=9,
                                          distribution={'downsampling': 'pool:2', 'compression': 'jpeg', 'compression_params': {'quality': 50, 'codec': 'sin'}})
        flow._fuse_pool = fused
        Y, c, C, ent, probs = flow.run_workflow(x)
        loss, parts = flow.training_step(x, t, lambda_nip=0.1, learning_rate=1e-4)
        res[fused] = (float(loss.numpy()), float(parts['ce'].numpy()), _grads(flow.nip._store), _grads(flow.fan._store))
    assert abs(res[True][0] - res[False][0]) <= 1e-6 * abs(res[False][0]) and abs(res[True][1] - res[False][1]) <= 1e-5 * abs(res[False][1])
    for k in (2, 3):
        for name, g in res[True][k].items():
            ref = res[False][k][name]
            e = np.abs(g - ref) / max(float(np.abs(ref).max()), 1e-30)
            assert np.mean(e > 2e-5) <= 0.02, (name, float(e.max()))


def test_fan_dropout_semantics():
    """FAN(dropout > 0) (models/forensics.py:86-88): Dropout follows each hidden dense layer; it is the identity in process(), in
    training_step (the reference calls the model without training=True there) and only acts in process(x, training=True)."""
    from neural_imaging_b200.models import forensics
    rs = np.random.RandomState(5)
    x = rs.uniform(size=(64, 16, 16, 3)).astype(np.float32)
    labels = rs.randint(0, 5, size=(64,))
    kw = dict(n_classes=5, patch_size=16, n_filters=8, n_convolutions=2, n_dense=2, seed=3)
    a, b = forensics.FAN(dropout=0.0, **kw), forensics.FAN(dropout=0.5, **kw)
    assert np.array_equal(a.process(x).numpy(), b.process(x).numpy())
    la, lb = a.training_step(x, labels, 1e-3), b.training_step(x, labels, 1e-3)
    assert float(la.numpy()) == float(lb.numpy())
    # (weight gradients are summed with atomics: run-to-run rounding differences, amplified to +-lr by Adam where a gradient is ~0)
    assert np.mean(np.abs(a._store.flat.cpu().numpy() - b._store.flat.cpu().numpy()) > 1e-6) < 0.02
    p0, p1, p2 = b.process(x).numpy(), b.process(x, training=True).numpy(), b.process(x, training=True).numpy()
    assert not np.array_equal(p0, p1) and not np.array_equal(p1, p2)          # active, and a fresh mask per call
    assert np.allclose(p1.sum(axis=1), 1.0, atol=1e-5)
    # the kernel itself: keep rate and inverted scaling
    from neural_imaging_b200 import _lib
    from neural_imaging_b200.tensor import as_device, empty, ptr, stream
    v = as_device(np.ones((1 << 16,), np.float32))
    out = empty(v.shape)
    _lib.lib().ni_dropout(ptr(v), ptr(out), v.numel(), 0.25, 1234, stream())
    o = out.cpu().numpy()
    assert set(np.unique(o)).issubset({0.0, np.float32(1.0 / 0.75)}) and abs(float((o > 0).mean()) - 0.75) < 0.01
