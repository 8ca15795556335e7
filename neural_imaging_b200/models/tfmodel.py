"""Base class of all toolbox models: performance-metric bookkeeping, parameter access / counting, save / load.

API mirror of reference models/tfmodel.py:86-294 (TFModel) without TensorFlow: weights live in a flat device
buffer (neural_imaging_b200.nn.ParamStore); snapshots are Keras-layout HDF5 weight files ``<dir>/<scoped_name>/<class>.h5``
(+ ``.json`` with the hyper-parameters) written and read by the dependency-free ``helpers/h5lite.py``; the ``.npz``
snapshots of earlier versions of this stack are still read.
"""
import json
import os
from pathlib import Path

import numpy as np

from ..helpers import h5lite, utils


# ---------------------------------------------------------------------------------------------------- Keras .h5 weight files
def _to_keras(p, a):
    """Product storage -> the Keras variable's shape (neural_imaging_b200/nn.py docstring)."""
    if p.keras == 'dense':                                   # (1, 1, in, out) -> (in, out)
        return a.reshape(a.shape[2], a.shape[3])
    if p.keras == 'conv2d_transpose':                        # 1x1 conv (1, 1, ci, (a*2+b)*F + f) -> (2, 2, F, ci)
        ci, f = a.shape[2], a.shape[3] // 4
        return np.ascontiguousarray(a.reshape(ci, 2, 2, f).transpose(1, 2, 3, 0))
    return a


def _from_keras(p, a):
    if p.keras == 'dense' and a.ndim == 2 and (1, 1) + a.shape == p.shape:
        return a.reshape(p.shape)
    if p.keras == 'conv2d_transpose' and a.ndim == 4 and a.shape[:2] == (2, 2) and p.shape == (1, 1, a.shape[3], 4 * a.shape[2]):
        return np.ascontiguousarray(a.transpose(3, 0, 1, 2)).reshape(p.shape)
    if tuple(a.shape) != p.shape:
        raise ValueError('shape mismatch for {}: weight file {} vs model {}'.format(p.name, tuple(a.shape), p.shape))
    return a


def save_weights_h5(store, filename):
    """``tf.keras.Model.save_weights(filename, save_format='h5')`` (reference models/tfmodel.py:159) for a ParamStore: the layout of
    Keras' ``save_weights_to_hdf5_group`` — root attributes ``layer_names`` / ``backend`` / ``keras_version``, one group per layer with
    a ``weight_names`` attribute and one dataset per variable at ``<layer>/<variable name>`` in the Keras storage shapes."""
    layers, order = {}, []
    for p in store.params:
        if p.keras == 'internal':
            continue
        layer = p.name.rsplit('/', 1)[0] if '/' in p.name else p.name
        if layer not in layers:
            layers[layer] = []
            order.append(layer)
        a = p.value.detach().cpu().numpy() if p.value is not None else p.init
        layers[layer].append((p.name + ':0', _to_keras(p, np.asarray(a, np.float32).reshape(p.shape))))
    tree, attrs = {}, {'': {'layer_names': np.array([n.encode('utf8') for n in order] or [b''], dtype='S'),
                            'backend': np.bytes_(b'tensorflow'), 'keras_version': np.bytes_(b'2.2.4-tf')}}

    def put(path, value):
        node = tree
        parts = path.split('/')
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        if value is None:
            node.setdefault(parts[-1], {})
        else:
            node[parts[-1]] = value
    for layer in order:
        put(layer, None)
        attrs[layer] = {'weight_names': np.array([n.encode('utf8') for n, _ in layers[layer]], dtype='S')}
        for n, a in layers[layer]:
            put(layer + '/' + n, np.asarray(a, np.float32))
    h5lite.write(filename, tree, attrs)


def load_weights_h5(store, filename):
    """``tf.keras.Model.load_weights`` on an h5 weight file (reference models/tfmodel.py:181): like Keras'
    ``load_weights_from_hdf5_group`` the variables are matched BY ORDER (layers in ``layer_names`` order, variables in
    ``weight_names`` order), not by name — Keras' automatic layer names depend on what else was built in the session. Chunked
    attributes (``layer_names0``, ``layer_names1`` ... for attributes above 64 KB) are joined as Keras does."""
    def attr_list(node, name):
        if name in node.attrs:
            vals = np.atleast_1d(node.attrs[name])
        else:
            vals, k = [], 0
            while '{}{}'.format(name, k) in node.attrs:
                vals.extend(np.atleast_1d(node.attrs['{}{}'.format(name, k)]))
                k += 1
        return [v.decode('utf8') if isinstance(v, bytes) else str(v) for v in vals]
    with h5lite.File(filename) as f:
        root = f['model_weights'] if 'layer_names' not in f.attrs and 'model_weights' in f else f   # model.save() files nest the weights
        values = []
        for layer in attr_list(root, 'layer_names'):
            if not layer:
                continue
            g = root[layer]
            for wn in attr_list(g, 'weight_names'):
                values.append((layer + '/' + wn, g[wn].read()))
    params = [p for p in store.params if p.keras != 'internal']
    if len(values) != len(params):
        raise ValueError('{} holds {} variables, the model has {}'.format(filename, len(values), len(params)))
    state = {p.name: _from_keras(p, np.asarray(a, dtype=np.float32)) for p, (_, a) in zip(params, values)}
    store.load_state_dict(state, strict=False)


def restore(dir_name, module, key=None, patch_size=None, restore_perf=False, fetch_stats=False):
    """Restore a trained model from a training directory (*.json training log + weights) — reference models/tfmodel.py:16-83, same
    arguments, errors and return value; the weights are Keras-layout .h5 files (load_model)."""
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
        return dirname, os.path.join(dirname, '{}.h5'.format(self.class_name.lower()))

    def save_model(self, dirname, epoch=0, save_args=False, quiet=False):
        """reference models/tfmodel.py:150-166: ``<dirname>/<scoped_name>/<class>.h5`` in the Keras weight-file layout (+ .json)."""
        dirname, filename = self._weights_path(dirname)
        os.makedirs(dirname, exist_ok=True)
        if self._store is not None:
            save_weights_h5(self._store, filename)
        if save_args:
            with open(os.path.join(dirname, '{}.json'.format(self.class_name.lower())), 'w') as f:
                json.dump({'model': self.class_name, 'args': self.get_hyperparameters()}, f, indent=4)

    def load_model(self, dirname, quiet=False):
        """reference models/tfmodel.py:168-182: the Keras h5 weight file; falls back to this stack's earlier .npz snapshots (the
        reference falls back to a TF checkpoint, which has no counterpart here)."""
        dirname, filename = self._weights_path(dirname)
        legacy = filename[:-3] + '.npz'
        if os.path.isfile(filename):
            load_weights_h5(self._store, filename)
        elif os.path.isfile(legacy):
            with np.load(legacy) as data:
                self._store.load_state_dict({k: data[k] for k in data.files})
        else:
            raise FileNotFoundError('No weights found at {} (or {})'.format(filename, legacy))
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
