"""Training loop of a stand-alone neural ISP (BASELINE config 2) — API mirror of reference training/pipeline.py:21-302 (`validate`,
`save_progress`, `train_nip_model`, `train_nip_bare`): same argument list, data-set checks, output-directory naming, resume from
progress.json, learning-rate schedule dictionary, validation schedule, best-checkpoint rule, the 0.95 learning-rate drop on a 20 %
deterioration and the early stop on a flat validation loss. The step is `model.training_step` (explicit forward / backward / Adam on
the device); validation develops every image with `model.process` and measures SSIM in the fused device kernel. Progress bars and
matplotlib figures are not reproduced.
"""
import json
import os
from collections import OrderedDict, deque

import numpy as np

from ..helpers import metrics


def _scalar(v):
    return float(np.asarray(v.numpy() if hasattr(v, 'numpy') else v).reshape(-1)[0])


def validate(model, data, out_directory, savefig=False, epoch=0, show_ref=False, loss_metric='L2'):
    """Develop the validation set one image at a time: (ssims, psnrs, losses, developed images) (reference :21-84, no figure)."""
    ssims, psnrs, losss = [], [], []
    if loss_metric not in ['L2', 'L1', 'SSIM', 'MS-SSIM']:
        raise ValueError('Unsupported loss ({})!'.format(loss_metric))
    developed_out = None
    for b in range(data.count_validation):
        example_x, example_y = data.next_validation_batch(b, 1)
        developed = np.asarray(model.process(example_x).numpy()).clip(0, 1)
        if developed_out is None:
            developed_out = np.zeros((data.count_validation,) + developed.shape[1:], dtype=np.float32)
        developed_out[b] = developed
        developed = developed.squeeze()
        reference = np.asarray(example_y).squeeze()
        ssim = float(metrics.ssim(reference, developed))
        psnr = float(metrics.psnr(reference, developed))
        if loss_metric == 'L2':
            loss = metrics.mse(255 * reference, 255 * developed)
        elif loss_metric == 'L1':
            loss = metrics.mae(255 * reference, 255 * developed)
        elif loss_metric == 'SSIM':
            loss = 255 * (1 - ssim)
        else:
            raise ValueError('Unsupported loss ({})!'.format(loss_metric))
        ssims.append(ssim)
        psnrs.append(psnr)
        losss.append(float(loss))
    return ssims, psnrs, losss, developed_out


def save_progress(model, training_summary, out_directory):
    output_stats = {
        'performance': model.performance,
        'args': model.get_hyperparameters(),
        'model': model.class_name,
        'init': repr(model),
        'summary': training_summary,
    }
    os.makedirs(out_directory, exist_ok=True)
    with open(os.path.join(out_directory, 'progress.json'), 'w') as f:
        json.dump(output_stats, f, indent=4)


def _shape(data, split, key):
    try:
        return tuple(int(v) for v in data[split][key].shape)
    except Exception:
        return None


def train_nip_model(model, camera_name, n_epochs=10000, lr_schedule=None, validation_loss_threshold=1e-3, validation_schedule=100,
                    resume=False, patch_size=64, batch_size=20, data=None, out_directory_root='./data/models/nip', save_best=False,
                    discard='flat', quiet=True):
    if data is None:
        raise ValueError('Training data seems not to be loaded!')
    try:
        batch_x, batch_y = data.next_training_batch(0, 5, patch_size * 2)
        if batch_x.shape != (5, patch_size, patch_size, 4) or batch_y.shape != (5, 2 * patch_size, 2 * patch_size, 3):
            raise ValueError('The training batch returned by the dataset instance is of invalid size!')
    except Exception as e:
        raise ValueError('Data set error: {}'.format(e))
    if batch_size > data.count_training or batch_size > data.count_validation:
        raise ValueError('Batch size ({}) exceeds dataset size ({}/{})!'.format(batch_size, data.count_training, data.count_validation))
    out_directory = os.path.join(out_directory_root, camera_name, model.model_code, model.scoped_name)
    if os.path.exists(out_directory) and not resume:
        print('WARNING directory {} exists, skipping...'.format(out_directory))
        return out_directory
    n_batches = data.count_training // batch_size
    n_tail = 5
    if not resume:
        start_epoch = 0
    else:
        summary_file = os.path.join(out_directory, 'progress.json')
        if not os.path.isfile(summary_file):
            raise FileNotFoundError('Could not open file {}'.format(summary_file))
        model.load_model(out_directory)
        with open(summary_file) as f:
            summary_data = json.load(f)
        model.performance = summary_data['performance']
        start_epoch = summary_data['summary']['Epoch']
    if lr_schedule is None:
        lr_schedule = {0: 1e-4}
    elif isinstance(lr_schedule, float):
        lr_schedule = {0: lr_schedule}
    training_summary = OrderedDict()
    training_summary['Camera'] = camera_name
    training_summary['Architecture'] = model.summary()
    training_summary['Max epochs'] = n_epochs
    training_summary['Learning rate'] = lr_schedule
    training_summary['Training data size'] = _shape(data, 'training', 'x')
    training_summary['Validation data size'] = _shape(data, 'validation', 'x')
    training_summary['# batches'] = n_batches
    training_summary['Patch size'] = patch_size
    training_summary['Batch size'] = batch_size
    training_summary['Validation schedule'] = validation_schedule
    training_summary['Start epoch'] = start_epoch
    training_summary['Saved checkpoint'] = None
    training_summary['Discarding policy'] = discard
    training_summary['Output directory'] = out_directory
    if not quiet:
        print('\n## Training summary')
        for k, v in training_summary.items():
            print('{:30s}: {}'.format(k, v))
    learning_rate = 1e-4
    epoch = start_epoch
    vloss = model.performance['loss']['validation']
    for epoch in range(start_epoch, n_epochs):
        if epoch in lr_schedule:
            learning_rate = lr_schedule[epoch]
        loss_local = []
        for batch_id in range(n_batches):
            batch_x, batch_y = data.next_training_batch(batch_id, batch_size, patch_size, discard=discard)
            loss_local.append(_scalar(model.training_step(batch_x, batch_y, learning_rate)))
        model.log_metric('loss', 'training', loss_local)
        if epoch % validation_schedule == 0:
            ssims, psnrs, v_losses, _ = validate(model, data, out_directory, True, epoch, True, loss_metric=model.loss_metric)
            model.log_metric('ssim', 'validation', ssims)
            model.log_metric('psnr', 'validation', psnrs)
            model.log_metric('loss', 'validation', v_losses)
            vloss = model.performance['loss']['validation']
            training_summary['Epoch'] = epoch
            save_progress(model, training_summary, out_directory)
            if not save_best or (len(vloss) > 2 and vloss[-1] <= min(vloss)):
                training_summary['Saved checkpoint'] = epoch
                model.save_model(out_directory, epoch, quiet=True)
            if len(vloss) > 5 and vloss[-1] > 1.2 * min(vloss):          # deteriorated by more than 20 %: drop the learning rate
                learning_rate = max((learning_rate * 0.95, 1e-7))
            if validation_loss_threshold is not None and len(vloss) > 10:
                current = np.mean(vloss[-n_tail:-1])
                previous = np.mean(vloss[-(n_tail + 1):-2])
                vloss_change = abs((current - previous) / previous)
                if vloss_change < validation_loss_threshold:
                    print('Early stopping - the model converged, validation loss change {}'.format(vloss_change))
                    break
            if not quiet:
                print('epoch {:5d}  loss {:.4f}  psnr {:.2f}  ssim {:.3f}  lr {:.1e}'.format(
                    epoch, model.pop_metric('loss', 'training'), model.pop_metric('psnr', 'validation'), model.pop_metric('ssim', 'validation'),
                    learning_rate), flush=True)
    training_summary['Epoch'] = epoch
    vloss = model.performance['loss']['validation']
    if not save_best or (vloss[-1] <= min(vloss)):
        training_summary['Saved checkpoint'] = epoch
        model.save_model(out_directory, epoch)
    save_progress(model, training_summary, out_directory)
    return out_directory


def train_nip_bare(model, camera_name, n_epochs=10000, lr_schedule=None, validation_loss_threshold=1e-3, validation_schedule=100,
                   resume=False, patch_size=64, batch_size=20, data=None, out_directory_root='./data/models/nip', save_best=False,
                   discard='flat'):
    """The loop without validation, logging or snapshots (reference :261-302; note that it trains at a fixed 1e-3, as upstream)."""
    out_directory = os.path.join(out_directory_root, camera_name, model.model_code, model.scoped_name)
    learning_rate = 1e-3
    for epoch in range(0, n_epochs):
        if hasattr(data, 'next_training_batch'):
            for batch_id in range(data.count_training // batch_size):
                batch_x, batch_y = data.next_training_batch(batch_id, batch_size, patch_size, discard=discard)
                model.training_step(batch_x, batch_y, learning_rate)
        else:
            for batch_x, batch_y in data:
                model.training_step(batch_x, batch_y, learning_rate)
    return out_directory
