"""Training loop of the manipulation-classification workflow on the B200 path.

API mirror of reference training/manipulation.py:21-300 (`default_training_specs`, `train_manipulation_nip`): same training
dictionary, required keys, data-set checks, output directory naming, learning-rate decay (x0.9 every 100 epochs, :136-138,
:268-269), validation schedule and model snapshots. The step itself is `flow.training_step` (one fused forward / backward /
Adam kernel sequence, replayed as CUDA graphs when `flow.enable_cuda_graph()` is on). Progress bars, matplotlib figures and
memory statistics of the reference are control-plane extras and are not reproduced.
"""
import os
import shutil
from collections import OrderedDict, deque

import numpy as np

from . import validation
from ..models.compression import DCN
from ..models.jpeg import JPEG


def default_training_specs():
    return {
        'use_pretrained_nip': True,
        'patch_size': 64,
        'batch_size': 10,
        'validation_schedule': 50,
        'n_epochs': 1001,
        'learning_rate': 1e-4,
        'run_number': 0,
        'lambda_nip': 0.1,
        'lambda_dcn': 0,
        'augment': False,
    }


def _scalar(v):
    return float(v.numpy()) if hasattr(v, 'numpy') else float(v)


def train_manipulation_nip(flow, training, data, directories=None, overwrite=False):
    """Train `flow` (ManipulationClassification) on `data` (object with next_training_batch / next_validation_batch /
    count_training / count_validation / is_raw_and_rgb / summary, as reference helpers/dataset.py). Returns the model directory."""
    directories_def = {'root': './data/m/', 'nip_snapshots': './data/models/nip/'}
    if directories is not None:
        directories_def.update(directories)
    directories = directories_def
    training_defaults = default_training_specs()
    if training is not None:
        training_defaults.update(training)
    training = training_defaults

    required_keys = {'camera_name', 'use_pretrained_nip', 'lambda_nip', 'lambda_dcn', 'run_number', 'n_epochs', 'learning_rate', 'augment'}
    if any(x not in training for x in required_keys):
        raise RuntimeError('Missing keys in the training dictionary! {}'.format(required_keys.difference(training.keys())))
    if data is None:
        raise ValueError('Training data seems not to be loaded!')
    try:
        ps = training['patch_size']
        if data.is_raw_and_rgb():
            batch_x, batch_y = data.next_training_batch(0, 1, ps * 2)
            if batch_x.shape != (1, ps, ps, 4) or batch_y.shape != (1, 2 * ps, 2 * ps, 3):
                raise ValueError('The RAW+RGB training batch is of invalid size! {}'.format(batch_x.shape))
        else:
            batch_x = data.next_training_batch(0, 1, ps * 2)
            if batch_x.shape != (1, 2 * ps, 2 * ps, 3):
                raise ValueError('The RGB training batch is of invalid size! {}'.format(batch_x.shape))
    except Exception as e:
        raise ValueError('Data set error: {}'.format(e))

    # root / camera_name / *Net / ln-0.1000 | fixed-nip / lc-0.1000 | fixed-codec / 001 /
    save_dir = [directories['root'], training['camera_name'], flow.nip.class_name]
    save_dir.append('ln-{:0.4f}'.format(training['lambda_nip']) if flow.is_trainable('nip') else 'fixed-nip')
    save_dir.append('lc-{:0.4f}'.format(training['lambda_dcn']) if flow.is_trainable('dcn') else 'fixed-codec')
    save_dir.append('{:03d}'.format(training['run_number']))
    save_dir = os.path.join(*save_dir)
    model_directory = os.path.join(save_dir, 'models')
    if os.path.exists(save_dir) and not overwrite:
        return model_directory
    if flow.is_trainable('nip') and flow.nip.count_parameters() == 0:
        raise ValueError('It looks like you`re trying to optimize a NIP with no trainable parameters!')

    learning_rate_decay_schedule, learning_rate_decay_rate = 100, 0.90
    learning_rate = training['learning_rate']
    n_batches = data.count_training // training['batch_size']
    if training['use_pretrained_nip'] and flow.nip.count_parameters() > 0:
        flow.nip.load_model(os.path.join(directories['nip_snapshots'], training['camera_name'], flow.nip.model_code))

    loss_epoch = {key: deque(maxlen=n_batches) for key in ('nip', 'fan')}
    summary = OrderedDict([('Problem', flow.summary()), ('Dataset', data.summary()), ('Camera name', training['camera_name']),
                           ('Classes', str(flow._forensics_classes)), ('Joint optimization', str(flow.trainable_models)),
                           ('# Epochs', training['n_epochs']), ('Batch size', training['batch_size']), ('Learning rate', training['learning_rate'])])
    raw_and_rgb = getattr(data, '_loaded_data', 'xy') == 'xy'
    learned_codec = flow.is_trainable('dcn') and isinstance(flow.codec, DCN)

    def validate(epoch, loss_type):
        """One validation round: FAN accuracy + confusion, ISP and codec quality when they are being trained (reference :224-246, :300-315)."""
        accuracy, conf = validation.validate_fan(flow, data)
        flow.fan.log_metric('accuracy', 'validation', accuracy)
        flow.fan.performance['confusion'] = conf.tolist()
        if flow.is_trainable('nip') and data.is_raw_and_rgb():
            for metric, values in zip(('ssim', 'psnr', 'loss'), validation.validate_nip(flow.nip, data, None, epoch=epoch, loss_type=loss_type)):
                flow.nip.log_metric(metric, 'validation', values)
        if flow.is_trainable('dcn'):
            if learned_codec:
                values = validation.validate_dcn(flow.codec, data, None, epoch=epoch)
            elif isinstance(flow.codec, JPEG):
                values = validation.validate_jpeg(flow.codec, data)
            else:
                raise NotImplementedError('Validation for {} codec doesn\'t seem to be implemented'.format(flow.codec))
            for metric, value in values.items():
                flow.codec.log_metric(metric, 'validation', value)

    def snapshot(epoch):
        validation.save_training_progress(summary, flow, save_dir, quiet=True)
        flow.fan.save_model(os.path.join(model_directory, flow.fan.scoped_name), epoch, quiet=True)
        if flow.is_trainable('nip'):
            flow.nip.save_model(os.path.join(model_directory, flow.nip.scoped_name), epoch, quiet=True)
        if learned_codec:
            flow.codec.save_model(os.path.join(model_directory, flow.codec.scoped_name), epoch, quiet=True)

    epoch = 0
    for epoch in range(0, training['n_epochs']):
        for batch_id in range(n_batches):
            if raw_and_rgb:
                batch_x, batch_y = data.next_training_batch(batch_id, training['batch_size'], 2 * training['patch_size'])
            else:
                batch_x = data.next_training_batch(batch_id, training['batch_size'], 2 * training['patch_size'])
                batch_y = batch_x
            comb_loss, comp_loss = flow.training_step(batch_x, batch_y, training['lambda_nip'], training['lambda_dcn'],
                                                      training['augment'], learning_rate)
            loss_epoch['fan'].append(_scalar(comb_loss))
            loss_epoch['nip'].append(_scalar(comp_loss['nip']))
        for name, model in (('nip', flow.nip), ('fan', flow.fan)):
            model.log_metric('loss', 'training', list(loss_epoch[name]))
        if epoch % training['validation_schedule'] == 0:
            validate(epoch, flow.nip.loss_metric)
            snapshot(epoch)
        if epoch % learning_rate_decay_schedule == 0:
            learning_rate *= learning_rate_decay_rate

    # the reference always closes with a validation round and a snapshot of the weights of the LAST epoch (:300-333), whatever the
    # validation schedule was (the ISP is scored with L2 here, as there)
    validate(epoch, 'L2')
    snapshot(epoch)
    progress = os.path.join(((flow._distribution.get('compression_params') or {}).get('dirname') or ''), flow.codec.scoped_name, 'progress.json') \
        if learned_codec else None
    if progress and os.path.isfile(progress):          # keep the codec's own training log next to its fine-tuned weights (:331-332)
        shutil.copyfile(progress, os.path.join(model_directory, flow.codec.scoped_name, 'progress.json'))
    return model_directory
