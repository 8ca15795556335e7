"""Training loop of the learned codec (BASELINE config 3) — API mirror of reference training/compression.py:107-309 (`save_progress`,
`train_dcn`): same training dictionary, batch counting, flip / gamma augmentation, learning-rate reduction schedule, validation
schedule and statistics (||x - y||, SSIM, latent entropy on the code book), progress.json layout, snapshots, early stopping on the
validation SSIM. The step itself is `dcn.training_step` (explicit forward / backward / Adam kernel sequence on the device); the
validation reconstructions come from `dcn.compress` / `dcn.decompress`. Progress bars, thumbnails, TensorBoard summaries and matplotlib
figures of the reference are control-plane extras and are not reproduced; the anti-aliased 'resize' augmentation needs scikit-image,
which is not part of this stack, and raises if its probability is non-zero.
"""
import json
import os
from collections import deque

import numpy as np

from ..helpers import metrics


def _scalar(v):
    return float(np.asarray(v.numpy() if hasattr(v, 'numpy') else v).reshape(-1)[0])


def _count(data, split, fallback):
    try:
        return int(data[split]['y'].shape[0])
    except Exception:
        return int(getattr(data, fallback))


def latent_entropy(batch_z, code_book):
    """helpers/stats.py:107-131: entropy (bits) of the values quantised to the code-book centroids, empty bins counted once."""
    code_book = np.asarray(code_book, dtype=np.float64).reshape(-1)
    top = np.abs(code_book).max() * 2
    edges = np.concatenate(([-top], np.convolve(code_book, [0.5, 0.5], mode='valid'), [top]))
    counts = np.histogram(np.asarray(batch_z).ravel(), bins=edges)[0].clip(min=1)
    probs = counts / counts.sum()
    return float(-np.sum(probs * np.log2(probs)))


def batch_gamma(batch_p, gamma=None):
    """helpers/image.py:22-28 (host-side augmentation of the reference)."""
    if gamma is None:
        gamma = np.array(np.random.uniform(low=0.25, high=3, size=(len(batch_p), 1, 1, 1)), dtype=np.float32)
    elif type(gamma) is float:
        gamma = gamma * np.ones((len(batch_p), 1, 1, 1))
    return np.power(batch_p, 1 / gamma).clip(0, 1)


def save_progress(dcn, data, training, out_dir):
    output_stats = {
        'training_spec': training,
        'data': data.summary(),
        'codec': {
            'model': dcn.class_name,
            'init': repr(dcn),
            'args': dcn.get_hyperparameters(),
            'codebook': np.asarray(dcn.get_codebook()).tolist(),
            'performance': dcn.performance,
        },
    }
    with open(os.path.join(out_dir, 'progress.json'), 'w') as f:
        json.dump(output_stats, f, indent=4)


def train_dcn(dcn, training, data, directory='./data/models/dcn/playground/', overwrite=False, tensorboard=False, quiet=True):
    """training = {'n_epochs', 'batch_size', 'patch_size', 'learning_rate', 'learning_rate_reduction_schedule',
    'learning_rate_reduction_factor', 'validation_schedule', 'convergence_threshold', 'augmentation_probs': {'resize', 'flip_h', 'flip_v',
    'gamma'}}. Returns the output directory (None when it exists and overwrite is off, like the reference)."""
    if tensorboard:
        raise NotImplementedError('TensorBoard summaries are not part of the B200 path')
    if training['augmentation_probs'].get('resize', 0) > 0:         # fail before the first epoch, not part-way through training
        raise NotImplementedError("the 'resize' augmentation (skimage.transform.resize, anti-aliased) is not available on this stack: set "
                                  "augmentation_probs['resize'] = 0")
    n_batches = _count(data, 'training', 'count_training') // training['batch_size']
    v_batches = _count(data, 'validation', 'count_validation') // training['batch_size']
    perf = dcn.performance
    caches = {k: {'training': deque(maxlen=max(n_batches, 1)), 'validation': deque(maxlen=max(v_batches, 1))} for k in ('loss', 'entropy', 'ssim')}
    n_tail = 5
    learning_rate = training['learning_rate']
    model_output_dirname = os.path.join(directory, dcn.model_code, dcn.scoped_name)
    if os.path.isdir(model_output_dirname) and not overwrite:
        print('WARNING Directory {} exists, skipping... (use overwrite=True)'.format(model_output_dirname))
        return None
    os.makedirs(model_output_dirname, exist_ok=True)
    probs = training['augmentation_probs']
    for epoch in range(0, training['n_epochs']):
        training['current_epoch'] = epoch
        if epoch > 0 and epoch % training['learning_rate_reduction_schedule'] == 0:
            learning_rate *= training['learning_rate_reduction_factor']
        for batch_id in range(n_batches):
            np.random.uniform()           # the reference's draw for the 'resize' augmentation (probability 0 here): keeps the np.random sequence aligned
            batch_x = data.next_training_batch(batch_id, training['batch_size'], training['patch_size'])
            if isinstance(batch_x, tuple):
                batch_x = batch_x[-1]
            if np.random.uniform() < probs['flip_h']:
                batch_x = batch_x[:, :, ::-1, :]
            if np.random.uniform() < probs['flip_v']:
                batch_x = batch_x[:, ::-1, :, :]
            if np.random.uniform() < probs['gamma']:
                batch_x = batch_gamma(batch_x)
            values = dcn.training_step(np.ascontiguousarray(batch_x, dtype=np.float32), learning_rate)
            for key, value in values.items():
                caches[key]['training'].append(_scalar(value))
        for key in ('loss', 'ssim', 'entropy'):
            perf[key]['training'].append(float(np.mean(caches[key]['training'])))
        codebook = dcn.get_codebook()
        if epoch % training['validation_schedule'] == 0:
            for batch_id in range(v_batches):
                batch_x = data.next_validation_batch(batch_id, training['batch_size'])
                if isinstance(batch_x, tuple):
                    batch_x = batch_x[-1]
                batch_x = np.asarray(batch_x)
                batch_z = dcn.compress(batch_x).numpy()
                batch_y = dcn.decompress(batch_z).numpy()
                caches['loss']['validation'].append(float(np.linalg.norm(batch_x - batch_y)))
                caches['ssim']['validation'].append(float(metrics.batch(batch_x, batch_y, metrics.ssim)))
                caches['entropy']['validation'].append(latent_entropy(batch_z, codebook))
            for key in ('loss', 'ssim', 'entropy'):
                perf[key]['validation'].append(float(np.mean(caches[key]['validation'])))
            save_progress(dcn, data, training, model_output_dirname)
            dcn.save_model(model_output_dirname, epoch, quiet=True)
            if len(perf['ssim']['validation']) > 5:
                current = np.mean(perf['ssim']['validation'][-n_tail:])
                previous = np.mean(perf['ssim']['validation'][-(n_tail + 1):-1])
                perf_change = abs((current - previous) / previous)
                if perf_change < training['convergence_threshold']:
                    print('Early stopping - the model converged, validation SSIM change {:.4f}'.format(perf_change))
                    break
                if current < 0.9 * previous:
                    print('Error - SSIM deterioration by more than 10% {:.4f} -> {:.4f}'.format(previous, current))
                    break
        if not quiet:
            print('epoch {:5d}  L {:.3f}  Lv {:.3f}  lr {:.1e}  ssim {:.2f}  H {:.1f}'.format(
                epoch, np.mean(perf['loss']['training'][-3:]), np.mean(perf['loss']['validation'][-1:]), learning_rate,
                perf['ssim']['validation'][-1], np.mean(perf['entropy']['training'][-1:])), flush=True)
    return model_output_dirname
