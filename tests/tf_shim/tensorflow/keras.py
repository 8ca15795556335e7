"""tf.keras subset for the tests-only TensorFlow stand-in (see package docstring): functional + subclassed models, the layers,
losses and the optimizer the reference's model files use. TEST INFRASTRUCTURE."""
import numpy as np
import torch
import torch.nn.functional as F

import tensorflow as tf
from tensorflow import (KerasTensor, Tensor, Variable, _evaluate, _find_sym, _map_sym, _max_pool_nhwc, _raw, _same_pads, _strides2,
                        _symbolic_call, _t, _td, _wrap, _PROBE)


class _NS(object):
    pass


layers, activations, initializers, losses, optimizers, backend, utils = _NS(), _NS(), _NS(), _NS(), _NS(), _NS(), _NS()

_name_counters = {}
_init_rng = np.random.RandomState(20201017)


def _clear_session():
    _name_counters.clear()


backend.clear_session = _clear_session
backend.floatx = lambda: 'float32'


def _snake(name):
    out = []
    for i, ch in enumerate(name):
        if ch.isupper() and i and (not name[i - 1].isupper()):
            out.append('_')
        out.append(ch.lower())
    return ''.join(out).lstrip('_')


def _unique_name(base):
    n = _name_counters.get(base, 0)
    _name_counters[base] = n + 1
    return base if n == 0 else '{}_{}'.format(base, n)


# ----------------------------------------------------------------------------------------------------------------- initializers
class GlorotUniform(object):
    def __call__(self, shape, dtype=None):
        shape = tuple(int(s) for s in shape)
        if len(shape) < 1:
            fan_in = fan_out = 1
        elif len(shape) == 1:
            fan_in = fan_out = shape[0]
        elif len(shape) == 2:
            fan_in, fan_out = shape
        else:
            rec = int(np.prod(shape[:-2]))
            fan_in, fan_out = shape[-2] * rec, shape[-1] * rec
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return torch.from_numpy(_init_rng.uniform(-lim, lim, size=shape))


class VarianceScaling(object):
    """Keras default VarianceScaling(scale=1, mode='fan_in', distribution='truncated_normal')."""

    def __init__(self, scale=1.0, mode='fan_in', distribution='truncated_normal', seed=None):
        self.scale = scale

    def __call__(self, shape, dtype=None):
        shape = tuple(int(s) for s in shape)
        rec = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        fan_in = shape[-2] * rec if len(shape) > 1 else shape[0]
        std = np.sqrt(self.scale / max(1.0, fan_in)) / .87962566103423978
        v = _init_rng.normal(0, std, size=shape)
        return torch.from_numpy(np.clip(v, -2 * std, 2 * std))


class Zeros(object):
    def __call__(self, shape, dtype=None):
        return torch.zeros(tuple(int(s) for s in shape), dtype=torch.float64)


initializers.VarianceScaling, initializers.GlorotUniform, initializers.Zeros = VarianceScaling, GlorotUniform, Zeros
initializers.Constant = tf.constant_initializer


def _get_initializer(init, default):
    if init is None:
        return default()
    if isinstance(init, str):
        return {'glorot_uniform': GlorotUniform, 'zeros': Zeros}[init]()
    if isinstance(init, type):
        return init()
    return init


# ----------------------------------------------------------------------------------------------------------------- activations
def _act(fn):
    def wrapper(x, *a, **k):
        if isinstance(x, KerasTensor):
            return _symbolic_call(wrapper, (x,) + a, k)
        return _wrap(fn(_raw(_t(x)), *a, **k))
    return wrapper


activations.relu = _act(lambda x: torch.relu(x))
activations.tanh = _act(lambda x: torch.tanh(x))
activations.sigmoid = _act(lambda x: torch.sigmoid(x))
activations.softsign = _act(lambda x: x / (1 + torch.abs(x)))
activations.softmax = _act(lambda x, axis=-1: torch.softmax(x, dim=axis))
activations.linear = _act(lambda x: x)


# ----------------------------------------------------------------------------------------------------------------- Layer base
class Layer(object):
    def __init__(self, *args, trainable=True, name=None, dtype=None, **kwargs):
        object.__setattr__(self, '_sublayers', [])
        object.__setattr__(self, '_own_weights', [])
        self.name = name or _unique_name(_snake(type(self).__name__))
        self.trainable = trainable
        self.built = False

    def __setattr__(self, key, value):
        if isinstance(value, Layer):
            if not any(value is l for l in self._sublayers):
                self._sublayers.append(value)
        elif isinstance(value, (list, tuple)) and len(value) and all(isinstance(v, Layer) for v in value):
            for v in value:
                if not any(v is l for l in self._sublayers):
                    self._sublayers.append(v)
        object.__setattr__(self, key, value)

    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, trainable=True, **kwargs):
        shape = () if shape is None else tuple(int(s) for s in shape)
        init = _get_initializer(initializer, GlorotUniform)
        value = torch.as_tensor(init(shape)).to(_td(dtype) if dtype is not None else tf.float32.torch).reshape(shape)
        v = Variable._make(value, '{}/{}'.format(self.name, name or 'Variable'), trainable)
        self._own_weights.append(v)
        return v

    def _tracked_layers(self):
        # attributes appended to tracked lists after assignment (e.g. DemosaicingLayer._layers.append) are found by a re-scan
        for value in list(self.__dict__.values()):
            if isinstance(value, list):
                for v in value:
                    if isinstance(v, Layer) and not any(v is l for l in self._sublayers):
                        self._sublayers.append(v)
        return self._sublayers

    @property
    def weights(self):
        w = list(self._own_weights)
        for l in self._tracked_layers():
            w.extend(l.weights)
        return w

    variables = weights

    @property
    def trainable_weights(self):
        if not self.trainable:
            return []
        w = [v for v in self._own_weights if v.trainable]
        for l in self._tracked_layers():
            w.extend(l.trainable_weights)
        return w

    trainable_variables = trainable_weights

    def build(self, input_shape):
        pass

    def call(self, inputs, *args, **kwargs):
        raise NotImplementedError

    def _autocast(self, x):
        """base_layer.__call__: numpy / python inputs become tensors; floating inputs are cast to the layer's dtype (floatx)."""
        if isinstance(x, (np.ndarray, float, int)) and not isinstance(x, KerasTensor):
            x = _t(x)
        if isinstance(x, torch.Tensor) and _raw(x).dtype.is_floating_point and _raw(x).dtype != tf.float32.torch:
            x = _wrap(_raw(x).to(tf.float32.torch))
        return x

    def __call__(self, *args, **kwargs):
        if _find_sym((args, kwargs)) is not None:
            return _symbolic_call(self.__call__, args, kwargs)
        args = tuple(self._autocast(a) if not isinstance(a, (list, tuple)) else type(a)(self._autocast(e) for e in a) for a in args)
        if not self.built:
            first = args[0]
            self.build(first[0].shape if isinstance(first, (list, tuple)) else first.shape)
            self.built = True
        return self.call(*args, **kwargs)


layers.Layer = Layer


def _apply_activation(act, x):
    return x if act is None else act(x)


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', activation=None, use_bias=True,
                 kernel_initializer='glorot_uniform', bias_initializer='zeros', **kwargs):
        super().__init__(**kwargs)
        self.filters = int(filters)
        self.kernel_size = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(int(k) for k in kernel_size)
        self.strides = _strides2(strides)
        self.padding = padding.upper()
        self.activation = activation
        self.use_bias = use_bias
        self._ki, self._bi = kernel_initializer, bias_initializer

    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.kernel = self.add_weight('kernel', self.kernel_size + (cin, self.filters), initializer=self._ki, trainable=True)
        self.bias = self.add_weight('bias', (self.filters,), initializer=self._bi, trainable=True) if self.use_bias else None

    def call(self, inputs):
        y = tf.nn.conv2d(inputs, self.kernel, [1, self.strides[0], self.strides[1], 1], self.padding)
        if self.use_bias:
            y = y + self.bias
        return _apply_activation(self.activation, y)


class Conv2DTranspose(Layer):
    """Keras Conv2DTranspose, kernel (kh, kw, Cout, Cin); only kernel == stride (the reference's 2x2 / stride-2 up-convolution):
    out[s*i + a, s*j + b, f] = sum_c in[i, j, c] * K[a, b, f, c] + bias[f] (conv2d_backprop_input of a stride-s VALID conv)."""

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', activation=None, use_bias=True,
                 kernel_initializer='glorot_uniform', bias_initializer='zeros', **kwargs):
        super().__init__(**kwargs)
        self.filters = int(filters)
        self.kernel_size = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(int(k) for k in kernel_size)
        self.strides = _strides2(strides)
        if self.kernel_size != self.strides:
            raise NotImplementedError('tf shim: Conv2DTranspose with kernel != stride')
        self.activation, self.use_bias, self._ki, self._bi = activation, use_bias, kernel_initializer, bias_initializer

    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.kernel = self.add_weight('kernel', self.kernel_size + (self.filters, cin), initializer=self._ki)
        self.bias = self.add_weight('bias', (self.filters,), initializer=self._bi) if self.use_bias else None

    def call(self, inputs):
        x = _raw(inputs)
        n, h, w, c = x.shape
        kh, kw = self.kernel_size
        k = _raw(self.kernel)                                        # (a, b, f, c)
        y = torch.einsum('nijc,abfc->niajbf', x, k).reshape(n, h * kh, w * kw, self.filters)
        if self.use_bias:
            y = y + _raw(self.bias)
        return _apply_activation(self.activation, _wrap(y))


class MaxPool2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding='valid', **kwargs):
        super().__init__(**kwargs)
        self.pool = _strides2(pool_size)
        self.strides = self.pool if strides is None else _strides2(strides)
        self.padding = padding

    def call(self, inputs):
        return _wrap(_max_pool_nhwc(_raw(inputs), self.pool, self.strides, self.padding))


class Concatenate(Layer):
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def call(self, inputs):
        return tf.concat(list(inputs), axis=self.axis)


class LeakyReLU(Layer):
    def __init__(self, alpha=0.3, **kwargs):
        super().__init__(**kwargs)
        self.alpha = float(alpha)

    def call(self, inputs):
        x = _raw(inputs)
        return _wrap(torch.where(x > 0, x, x * self.alpha))          # LeakyRelu kernel: features > 0 ? features : alpha * features


class GlobalAveragePooling2D(Layer):
    def call(self, inputs):
        return _wrap(_raw(inputs).mean(dim=(1, 2)))


class Flatten(Layer):
    def call(self, inputs):
        x = _raw(inputs)
        return _wrap(x.reshape(x.shape[0], -1))


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer='glorot_uniform', bias_initializer='zeros', **kwargs):
        super().__init__(**kwargs)
        self.units, self.activation, self.use_bias, self._ki, self._bi = int(units), activation, use_bias, kernel_initializer, bias_initializer

    def build(self, input_shape):
        self.kernel = self.add_weight('kernel', (int(input_shape[-1]), self.units), initializer=self._ki)
        self.bias = self.add_weight('bias', (self.units,), initializer=self._bi) if self.use_bias else None

    def call(self, inputs):
        y = tf.matmul(inputs, self.kernel)
        if self.use_bias:
            y = y + self.bias
        return _apply_activation(self.activation, y)


class Dropout(Layer):
    """Identity unless called with training=True (the reference's training steps call `self._model(x)` without it)."""

    def __init__(self, rate, **kwargs):
        super().__init__(**kwargs)
        self.rate = rate

    def call(self, inputs, training=None):
        if training:
            raise NotImplementedError('tf shim: dropout in training mode')
        return inputs


layers.Conv2D, layers.Conv2DTranspose, layers.MaxPool2D, layers.MaxPooling2D = Conv2D, Conv2DTranspose, MaxPool2D, MaxPool2D
layers.Concatenate, layers.LeakyReLU, layers.GlobalAveragePooling2D, layers.Flatten = Concatenate, LeakyReLU, GlobalAveragePooling2D, Flatten
layers.Dense, layers.Dropout = Dense, Dropout


# ----------------------------------------------------------------------------------------------------------------- Input / Model
def Input(shape=None, batch_size=None, name=None, dtype=None, **kwargs):
    dims = [1] + [(_PROBE if s is None else int(s)) for s in shape]
    dyn = any(s is None for s in shape)
    probe = _wrap(torch.zeros(dims, dtype=_td(dtype) if dtype is not None else tf.float32.torch))
    return KerasTensor(None, (), {}, None, probe, dyn)


class Model(Layer):
    """tf.keras.Model: functional (inputs=, outputs=) or subclassed (call())."""

    def __init__(self, *args, **kwargs):
        inputs = kwargs.pop('inputs', None)
        outputs = kwargs.pop('outputs', None)
        if len(args) == 2:
            inputs, outputs = args
        name = kwargs.pop('name', None)
        super().__init__(name=name)
        self._functional = inputs is not None and outputs is not None
        if self._functional:
            # network.py _init_graph_network: single-element lists are unwrapped
            if isinstance(inputs, list) and len(inputs) == 1:
                inputs = inputs[0]
            if isinstance(outputs, list) and len(outputs) == 1:
                outputs = outputs[0]
            self._nested_inputs, self._nested_outputs = inputs, outputs
            self.inputs = [inputs] if isinstance(inputs, KerasTensor) else list(inputs)
            self.outputs = [outputs] if isinstance(outputs, KerasTensor) else list(outputs)
            self._graph_layers = []
            seen = set()

            def walk(sym):
                if id(sym) in seen or sym._fn is None:
                    return
                seen.add(id(sym))
                _map_sym((sym._args, sym._kwargs), walk)
                owner = getattr(sym._fn, '__self__', None)
                if isinstance(owner, Layer) and not any(owner is l for l in self._graph_layers):
                    self._graph_layers.append(owner)
            for o in self.outputs:
                walk(o)
            for l in self._graph_layers:
                if not any(l is s for s in self._sublayers):
                    self._sublayers.append(l)
            self.built = True

    @property
    def layers(self):
        return list(self._tracked_layers())

    def call(self, inputs, training=None):
        if not self._functional:
            raise NotImplementedError
        memo = {}
        vals = [inputs] if len(self.inputs) == 1 and not isinstance(inputs, (list, tuple)) else list(inputs)
        for sym, v in zip(self.inputs, vals):
            memo[id(sym)] = self._autocast(_t(v))
        out = _map_sym(self._nested_outputs, lambda s: _evaluate(s, memo))
        return out

    def __call__(self, *args, **kwargs):
        if self._functional:
            if _find_sym((args, kwargs)) is not None:
                return _symbolic_call(self.__call__, args, kwargs)
            return self.call(*args, **kwargs)
        return super().__call__(*args, **kwargs)

    def summary(self, *a, **k):
        print('Model {}: {} parameters'.format(self.name, int(sum(np.prod(v.shape) for v in self.trainable_weights))))

    def get_weights(self):
        return [v.numpy() for v in self.weights]

    def set_weights(self, values):
        for v, a in zip(self.weights, values):
            v.assign(a)

    def save_weights(self, filename, save_format=None):
        """Stand-in for the Keras h5 weight file: the variables in topological order (which is how Keras' h5 loader matches them)."""
        with open(filename, 'wb') as f:
            np.savez(f, *[v.numpy() for v in self.weights])

    def load_weights(self, filename):
        with np.load(filename) as d:
            ws = self.weights
            assert len(d.files) == len(ws)
            for i, v in enumerate(ws):
                v.assign(d['arr_%d' % i])


def _plot_model(*a, **k):
    raise NotImplementedError


utils.plot_model = _plot_model


# ----------------------------------------------------------------------------------------------------------------- losses
class SparseCategoricalCrossentropy(object):
    """Keras 2.1 eager path (backend.sparse_categorical_crossentropy, from_logits=False): q = clip(p, 1e-7, 1 - 1e-7);
    sparse_softmax_cross_entropy_with_logits(labels, log q) = -(log q[label] - log sum_k q_k); reduction SUM_OVER_BATCH_SIZE."""

    def __init__(self, from_logits=False):
        if from_logits:
            raise NotImplementedError

    def __call__(self, y_true, y_pred, sample_weight=None):
        eps = 1e-7
        q = tf.clip_by_value(y_pred, eps, 1 - eps)
        logq = _raw(tf.math.log(q))
        labels = torch.as_tensor(np.asarray(y_true)).to(torch.int64).reshape(-1)
        ce = torch.logsumexp(logq, dim=-1) - logq.gather(1, labels.reshape(-1, 1)).reshape(-1)
        if sample_weight is not None:
            ce = ce * _raw(_t(sample_weight)).to(ce.dtype)
        return _wrap(ce.sum() / ce.numel())


class MeanSquaredError(object):
    """losses.mean_squared_error = mean over the last axis of (y_pred - y_true)^2; optional sample_weight multiplies the per-element
    losses (losses_utils.compute_weighted_loss) before sum / number of elements."""

    def __call__(self, y_true, y_pred, sample_weight=None):
        a, b = _raw(_t(y_true)), _raw(_t(y_pred))
        if a.dtype != b.dtype:
            a = a.to(b.dtype)          # losses cast y_true to y_pred.dtype
        per = ((b - a) ** 2).mean(dim=-1)
        if sample_weight is not None:
            per = per * _raw(_t(np.float32(sample_weight) if not isinstance(sample_weight, torch.Tensor) else sample_weight)).to(per.dtype)
        return _wrap(per.sum() / per.numel())


losses.SparseCategoricalCrossentropy, losses.MeanSquaredError = SparseCategoricalCrossentropy, MeanSquaredError


# ----------------------------------------------------------------------------------------------------------------- optimizer
class _HyperVar(object):
    def __init__(self, v):
        self.v = float(v)

    def assign(self, v):
        self.v = float(np.asarray(v))

    def numpy(self):
        return np.float32(self.v)


class Adam(object):
    """optimizer_v2/adam.py + training_ops ApplyAdam (non-amsgrad, dense): per variable
        lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m += (g - m)(1 - b1);  v += (g^2 - v)(1 - b2);  var -= lr_t * m / (sqrt(v) + eps)
    with t = iterations + 1 shared by all variables of one apply_gradients call; None gradients are skipped (_filter_grads)."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, **kwargs):
        self.lr = _HyperVar(kwargs.get('lr', learning_rate))
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.iterations = 0
        self._slots = {}

    @property
    def learning_rate(self):
        return self.lr

    def apply_gradients(self, grads_and_vars, name=None):
        t = self.iterations + 1
        # coefficients are formed in the variable's dtype as the kernel's scalar inputs are cast to it
        for g, var in grads_and_vars:
            if g is None:
                continue
            dt = _raw(var).dtype
            np_dt = np.float32 if dt == torch.float32 else np.float64
            b1, b2, eps, lr = np_dt(self.beta_1), np_dt(self.beta_2), np_dt(self.epsilon), np_dt(self.lr.v)
            b1p, b2p = np.power(b1, np_dt(t)), np.power(b2, np_dt(t))
            alpha = lr * np.sqrt(np_dt(1) - b2p) / (np_dt(1) - b1p)
            slot = self._slots.get(id(var))
            if slot is None:
                slot = self._slots[id(var)] = (torch.zeros_like(_raw(var).detach()), torch.zeros_like(_raw(var).detach()), var)
            m, v, _ = slot
            gg = _raw(g).to(dt)
            m += (gg - m) * float(np_dt(1) - b1)
            v += (gg * gg - v) * float(np_dt(1) - b2)
            with torch.no_grad():
                _raw(var).sub_((m * float(alpha)) / (torch.sqrt(v) + float(eps)))
        self.iterations = t


optimizers.Adam = Adam
