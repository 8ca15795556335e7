"""Base class of all toolbox models: performance-metric bookkeeping, parameter access / counting, save / load.

API mirror of reference models/tfmodel.py:86-294 (TFModel) without TensorFlow: weights live in a flat device
buffer (neural_imaging_b200.nn.ParamStore); snapshots are ``<dir>/<scoped_name>/<class>.npz`` (+ ``.json`` with the
hyper-parameters) because h5py / Keras are not part of this stack (h5 interchange is a "next" row).
"""
import json
import os
from pathlib import Path

import numpy as np

from ..helpers import utils


def restore(dir_name, module, key=None, patch_size=None, restore_perf=False, fetch_stats=False):
    """Restore a trained model from a training directory (*.json training log + weights) — reference models/tfmodel.py:16-83, same
    arguments, errors and return value; the weights are this stack's .npz snapshots (Keras .h5 interchange: SURVEY 8f N4)."""
    training_log_path = None
    if dir_name is None:
        raise ValueError('dcn directory cannot be None')
    if not os.path.exists(dir_name):
        preset_file = 'config/presets/{}.json'.format(module.__name__.split('.')[-1])
        if os.path.isfile(preset_file):
            with open(preset_file) as f:
                presets = json.load(f)
            if dir_name in presets:
                dir_name = presets[dir_name]
            else:
                raise ValueError('Directory {} does not exist & key not found in presets (config/presets/*)!'.format(dir_name))
        else:
            raise ValueError('Directory {} does not exist (presets not available)!'.format(dir_name))
    for filename in Path(dir_name).glob('**/*.json'):
        training_log_path = str(filename)
    if training_log_path is None:
        raise FileNotFoundError('Could not find a training log (JSON file) in {}'.format(dir_name))
    with open(training_log_path) as f:
        training_log = json.load(f)
    if key is not None:
        training_log = training_log[key]
    parameters = dict(training_log['args'])
    parameters['patch_size'] = patch_size
    for k, value in parameters.items():
        if isinstance(value, str) and value and value[0] == '(' and value[-1] == ')':
            parameters[k] = eval(value)         # tuples are stored as strings (JSON has no tuple type)
    model = getattr(module, training_log['model'])(**parameters)
    model.load_model(dir_name)
    if restore_perf:
        model.performance = training_log['performance']
    if fetch_stats:
        stats = {}
        for k, v in model.performance.items():
            if 'validation' in v and len(v['validation']) > 0:
                stats[k] = np.round(v['validation'][-1], 3)
            elif 'training' in v and len(v['training']) > 0:
                stats[k] = np.round(v['training'][-1], 3)
        return model, stats
    return model


class TFModel(object):

    def __init__(self, **kwargs):
        self._store = None
        self.reset_performance_stats()

    # ---- metrics (tfmodel.py:110-131)
    @staticmethod
    def _reset_performance(metrics):
        return {k: {'training': [], 'validation': []} for k in metrics}

    def reset_performance_stats(self):
        self.performance = self._reset_performance(['loss'])

    def log_metric(self, metric, scope, value, raw=False):
        if not raw:
            if hasattr(value, 'numpy'):
                value = value.numpy()
            value = float(value) if utils.is_number(value) else float(np.mean(value))
        self.performance[metric][scope].append(value)

    def pop_metric(self, metric, scope):
        return self.performance[metric][scope][-1]

    # ---- parameters
    @property
    def parameters(self):
        """Trainable parameters (device views, Keras layer order: kernel then bias)."""
        return [] if self._store is None else [p.value for p in self._store.trainable]

    @property
    def variables(self):
        return [] if self._store is None else [p.value for p in self._store.params]

    @property
    def parameter_names(self):
        return [] if self._store is None else [p.name for p in self._store.trainable]

    def count_parameters(self):
        return 0 if self._store is None else self._store.count()

    def count_parameters_breakdown(self):
        total = max(self.count_parameters(), 1)
        return [(p.name, p.shape, p.size, round(100 * p.size / total, 1)) for p in self._store.trainable]

    # ---- persistence
    def _weights_path(self, dirname):
        if not dirname.endswith(self.scoped_name):
            dirname = os.path.join(dirname, self.scoped_name)
        return dirname, os.path.join(dirname, '{}.npz'.format(self.class_name.lower()))

    def save_model(self, dirname, epoch=0, save_args=False, quiet=False):
        dirname, filename = self._weights_path(dirname)
        os.makedirs(dirname, exist_ok=True)
        np.savez(filename, **(self._store.state_dict() if self._store is not None else {}))
        if save_args:
            with open(os.path.join(dirname, '{}.json'.format(self.class_name.lower())), 'w') as f:
                json.dump({'model': self.class_name, 'args': self.get_hyperparameters()}, f, indent=4)

    def load_model(self, dirname, quiet=False):
        dirname, filename = self._weights_path(dirname)
        if not os.path.isfile(filename):
            raise FileNotFoundError('No weights found at {} (Keras .h5 snapshots are not readable here)'.format(filename))
        with np.load(filename) as data:
            self._store.load_state_dict({k: data[k] for k in data.files})
        self.reset_performance_stats()

    @classmethod
    def restore(cls, dir_name, *, key=None, patch_size=None):
        candidates = list(Path(dir_name).glob('**/*.json'))
        path = str(candidates[0]) if candidates else None
        if path is None or not os.path.isfile(path):
            raise FileNotFoundError('Could not find a training log (JSON file) in {}'.format(dir_name))
        with open(path) as f:
            log = json.load(f)
        if key is not None:
            log = log[key]
        params = log['args']
        if patch_size is not None:
            params['patch_size'] = patch_size
        for k, v in params.items():
            if isinstance(v, str) and v and v[0] == '(' and v[-1] == ')':
                params[k] = eval(v)    # tuples are stored as strings (JSON has no tuple type)
        instance = cls(**params)
        instance.load_model(dir_name)
        return instance

    # ---- naming
    @property
    def class_name(self):
        return type(self).__name__

    @property
    def scoped_name(self):
        return type(self).__name__.lower()

    @property
    def model_code(self):
        raise NotImplementedError()

    def summary(self):
        return '{} model [{:,.0f} parameters]'.format(self.class_name, self.count_parameters())

    def summary_compact(self):
        return '{}'.format(self.class_name)

    def get_hyperparameters(self):
        return self._h.to_json() if hasattr(self, '_h') else None

    def __repr__(self):
        try:
            extra = utils.join_args(self._h.changed_params())
        except Exception:
            extra = ''
        return '{}({})'.format(self.class_name, extra)

    def _has_attributes(self, attrs, message='Expected attributes not found: {}'):
        missing = [k for k in attrs if not hasattr(self, k)]
        if missing:
            raise NotImplementedError(message.format(missing))

    def process(self, x, training=False):
        raise NotImplementedError()


class Placeholder(object):
    """Stand-in for tf.keras.Input: callers only read ``.shape`` (e.g. flow.nip.x.shape[-1])."""

    def __init__(self, shape):
        self.shape = (None,) + tuple(shape)
