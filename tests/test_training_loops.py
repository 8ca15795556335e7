"""Host logic of the stand-alone training loops (reference training/pipeline.py:105-258 `train_nip_model`, training/compression.py:123-309
`train_dcn`) with stand-in models: schedules, directory layout, resume, snapshots, progress.json, early stopping, error behaviour.
The real models' `training_step` / `process` / `compress` are covered by the GPU parity tests; nothing here needs a device (the SSIM
metric, a device kernel in the product, is replaced by a NumPy stand-in for these tests only)."""
import json
import os

import numpy as np
import pytest

from neural_imaging_b200.models.tfmodel import TFModel


class _Data:
    def __init__(self, n_train=8, n_val=4, raw=True, size=32):
        self.count_training, self.count_validation, self.raw, self.size = n_train, n_val, raw, size
        self.calls = []
        rs = np.random.RandomState(0)
        self._val_y = rs.uniform(size=(n_val, size, size, 3)).astype(np.float32)

    def __getitem__(self, key):
        n = self.count_training if key == 'training' else self.count_validation
        return {'x': np.zeros((n, self.size // 2, self.size // 2, 4)), 'y': np.zeros((n, self.size, self.size, 3))}

    def next_training_batch(self, batch_id, batch_size, rgb_patch_size, discard='flat'):
        self.calls.append((batch_id, batch_size, rgb_patch_size, discard))
        y = np.full((batch_size, rgb_patch_size, rgb_patch_size, 3), 0.25, np.float32)
        y[:, 0, 0, 0] = 1.0                     # a marker pixel: flips move it
        if not self.raw:
            return y
        return np.zeros((batch_size, rgb_patch_size // 2, rgb_patch_size // 2, 4), np.float32), y

    def next_validation_batch(self, batch_id, batch_size):
        y = self._val_y[batch_id * batch_size:(batch_id + 1) * batch_size]
        return (np.zeros((len(y), self.size // 2, self.size // 2, 4), np.float32), y) if self.raw else y

    def summary(self):
        return 'stand-in data'


class _NIP(TFModel):
    loss_metric = 'L2'
    model_code = 'FakeNet'

    def __init__(self, losses=None):
        super().__init__()
        self.performance = self._reset_performance(['loss', 'psnr', 'ssim'])
        self.lrs, self.saved, self.loaded, self.quality = [], [], None, 0.5
        self._losses = losses

    def training_step(self, x, y, lr):
        self.lrs.append(lr)
        return 10.0 / len(self.lrs)

    def process(self, x):
        n = len(x)

        class _T:
            def numpy(_):
                return np.full((n, 32, 32, 3), self.quality, np.float32)
        if self._losses:
            self.quality = self._losses[min(len(self.performance['loss']['validation']), len(self._losses) - 1)]
        return _T()

    def save_model(self, dirname, epoch=0, quiet=False):
        self.saved.append(epoch)

    def load_model(self, dirname):
        self.loaded = dirname

    def get_hyperparameters(self):
        return {'in_channels': 4}

    def summary(self):
        return 'FakeNet summary'


@pytest.fixture
def numpy_ssim(monkeypatch):
    from neural_imaging_b200.helpers import metrics
    monkeypatch.setattr(metrics, 'ssim', lambda a, b: 1.0 - float(np.mean(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)))))


class _DCN(TFModel):
    model_code = 'FakeDCN/8c'

    class _H:
        scale_latent = False
        train_codebook = False

    def __init__(self, ssims):
        super().__init__()
        self.performance = self._reset_performance(['loss', 'entropy', 'ssim', 'psnr'])
        self._h, self.lrs, self.saved, self.markers, self._ssims = self._H(), [], [], [], ssims

    def training_step(self, x, lr):
        self.lrs.append(lr)
        self.markers.append(tuple(int(v) for v in np.argwhere(x[0, :, :, 0] == 1.0)[0]))
        assert x.flags['C_CONTIGUOUS'] and x.dtype == np.float32
        return {'loss': 4.0, 'ssim': 0.5, 'entropy': 2.0}

    def compress(self, x):
        class _T:
            def numpy(_):
                return np.round(np.linspace(-3, 3, len(x) * 16)).reshape(len(x), 4, 4, 1)
        return _T()

    def decompress(self, z):
        n_val = len(self.performance['ssim']['validation'])
        err = 1.0 - self._ssims[min(n_val, len(self._ssims) - 1)]

        class _T:
            def numpy(_):
                return self._x - err
        return _T()

    def get_codebook(self):
        return np.arange(-15, 17, dtype=np.float32)

    def get_hyperparameters(self):
        return {'n_features': 8}

    def save_model(self, dirname, epoch=0, quiet=False):
        self.saved.append(epoch)


def test_train_dcn_loop(tmp_path, numpy_ssim, monkeypatch):
    from neural_imaging_b200.training import compression
    data = _Data(raw=False)
    spec = {'n_epochs': 30, 'batch_size': 4, 'patch_size': 32, 'learning_rate': 1e-3, 'learning_rate_reduction_schedule': 2,
            'learning_rate_reduction_factor': 0.5, 'validation_schedule': 1, 'convergence_threshold': 1e-4,
            'augmentation_probs': {'resize': 0.0, 'flip_h': 1.0, 'flip_v': 0.0, 'gamma': 0.0}}
    dcn = _DCN(ssims=[0.5, 0.6, 0.7, 0.8, 0.85, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9])
    orig = data.next_validation_batch
    monkeypatch.setattr(data, 'next_validation_batch', lambda b, n: setattr(dcn, '_x', orig(b, n)) or orig(b, n))
    out = compression.train_dcn(dcn, spec, data, directory=str(tmp_path))
    assert out == os.path.join(str(tmp_path), 'FakeDCN/8c', dcn.scoped_name) and os.path.isdir(out)
    assert dcn.lrs[:2] == [1e-3, 1e-3] and dcn.lrs[4] == 0.5e-3 and dcn.lrs[8] == 0.25e-3           # 2 batches / epoch, halved every 2 epochs
    assert all(m == (0, 31) for m in dcn.markers)                                                    # flip_h moved the marker to the last column
    n_epochs_run = len(dcn.performance['loss']['training'])
    assert 6 < n_epochs_run < 30 and dcn.saved == list(range(n_epochs_run))                          # early stop once the validation SSIM is flat
    assert dcn.performance['loss']['training'][0] == 4.0 and dcn.performance['entropy']['training'][0] == 2.0
    assert abs(dcn.performance['ssim']['validation'][0] - 0.5) < 1e-6 and spec['current_epoch'] == n_epochs_run - 1
    log = json.load(open(os.path.join(out, 'progress.json')))
    assert set(log) == {'training_spec', 'data', 'codec'} and set(log['codec']) == {'model', 'init', 'args', 'codebook', 'performance'}
    assert log['codec']['model'] == '_DCN' and len(log['codec']['codebook']) == 32 and log['data'] == 'stand-in data'
    assert 0 < log['codec']['performance']['entropy']['validation'][0] < 5
    # existing directory: skipped unless overwrite; deterioration by more than 10 % stops the loop; unavailable augmentation raises
    assert compression.train_dcn(_DCN([0.5]), spec, data, directory=str(tmp_path)) is None
    bad = _DCN(ssims=[0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.2, 0.2])
    monkeypatch.setattr(data, 'next_validation_batch', lambda b, n: setattr(bad, '_x', orig(b, n)) or orig(b, n))
    compression.train_dcn(bad, dict(spec, convergence_threshold=0.0), data, directory=str(tmp_path), overwrite=True)
    assert len(bad.performance['ssim']['validation']) == 7
    with pytest.raises(NotImplementedError):
        compression.train_dcn(_DCN([0.5]), dict(spec, augmentation_probs={'resize': 1.0, 'flip_h': 0, 'flip_v': 0, 'gamma': 0}), data,
                              directory=str(tmp_path / 'r'))
    assert abs(compression.latent_entropy(np.array([0.0, 0.0, 1.0, 1.0]), np.arange(-1, 3)) - (-(2 * 2 / 6 * np.log2(2 / 6) + 2 * 1 / 6 * np.log2(1 / 6)))) < 1e-12


def test_codec_statistics_helpers_match_reference_execution():
    """latent_entropy (helpers/stats.py:107-131, used by codec.compress_n_stats and train_dcn) and batch_gamma (helpers/image.py:22-28) against
    tests/golden/codec_stats.npz, which tests/golden/make_codec_stats_golden.py produced by executing the reference's own functions."""
    from conftest import GOLDEN
    from neural_imaging_b200.training import compression
    with np.load(os.path.join(GOLDEN, 'codec_stats.npz')) as d:
        g = {k: d[k] for k in d.files}
    n = 0
    while 'z_%d_0' % n in g:
        for j in range(3):
            key = '%d_%d' % (n, j)
            assert abs(compression.latent_entropy(g['z_' + key], g['cb_' + key]) - float(g['entropy_' + key])) < 1e-12
        n += 1
    assert n == 3
    assert np.array_equal(compression.batch_gamma(g['gamma_in'], 2.0), g['gamma_out_2p0'])
    assert np.array_equal(compression.batch_gamma(g['gamma_in'], g['gamma_vec']), g['gamma_out_vec'])
    np.random.seed(3)
    out = compression.batch_gamma(g['gamma_in'])
    assert out.shape == g['gamma_in'].shape and out.dtype == np.float32 and 0 <= out.min() and out.max() <= 1
    # the code path of codec.compress_n_stats imports the same function
    import inspect
    from neural_imaging_b200.compression import codec
    assert 'latent_entropy' in inspect.getsource(codec.compress_n_stats)


def test_save_training_progress_layout(tmp_path):
    """training/validation.py:301-352: the training.json written at every validation epoch of the manipulation-classification loop."""
    import types
    from neural_imaging_b200.helpers import paramspec
    from neural_imaging_b200.training import validation

    def model(name, perf, with_h=True):
        m = types.SimpleNamespace(class_name=name, performance=perf)
        if with_h:
            m._h = paramspec.ParamSpec({'n_filters': (32, int, (1, 1024)), 'activation': ('leaky_relu', str, {'leaky_relu', 'relu'})})
        return m
    flow = types.SimpleNamespace(
        _distribution={'downsampling': 'pool:2', 'compression': 'jpeg', 'compression_params': {'quality': (50, 90), 'codec': 'soft', 'odd': np.float32(1.5)}},
        _forensics_classes=['native', 'sharpen:1'],
        nip=model('UNet', {'loss': {'training': [1.0, 0.5], 'validation': [0.7]}}, with_h=False),
        fan=model('FAN', {'loss': {'training': [2.0], 'validation': []}, 'accuracy': {'training': [], 'validation': [0.4]}, 'confusion': [[1.0, 0.0], [0.5, 0.5]]}),
        codec=model('JPEG', {'entropy': {'training': [], 'validation': [float('nan')]}}, with_h=False))
    path = validation.save_training_progress({'Problem': 'toy', '# Epochs': 3}, flow, str(tmp_path / 'run' / '001'), quiet=True)
    log = json.load(open(path))
    assert os.path.basename(path) == 'training.json' and list(log) == ['summary', 'distribution', 'manipulations', 'nip', 'forensics', 'codec']
    assert log['nip']['args'] == {} and log['forensics']['args'] == {'n_filters': 32, 'activation': 'leaky_relu'} and 'args' not in log['codec']
    assert log['distribution']['compression_params']['quality'] == [50, 90] and log['forensics']['performance']['confusion'][1] == [0.5, 0.5]
    assert log['manipulations'] == ['native', 'sharpen:1'] and log['summary']['# Epochs'] == 3
    flow.codec = None
    assert 'codec' not in json.load(open(validation.save_training_progress({}, flow, str(tmp_path / 'run' / '002'), quiet=True)))
