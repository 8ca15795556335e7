"""A tests-only stand-in for the `tensorflow` package (TEST INFRASTRUCTURE — never imported by the product).

Purpose: let `tests/golden/make_tf_graph_golden.py` import and EXECUTE the reference's own, unmodified TF-2.1 model code
(`/root/reference/models/*.py`, `helpers/tf_helpers.py`, `workflows/manipulation_classification.py`) where TensorFlow itself
cannot be installed. The graph wiring — transposes, block order, constants, loss composition, which tensors carry gradient, the
optimizer calls — is then the reference's, executed; only the per-op semantics of the ~90 `tf.*` entry points those files use are
restated here, on torch-CPU tensors, each from the TF 2.1 kernel / Python definition named in its docstring. The restatement is
written independently of `oracle/ref_ops.py` (different formulation wherever there was a choice), so that agreement between the two
is evidence and not a tautology.

Precision: `set_precision('float64')` maps tf.float32 (and Keras' floatx) to float64, so the same reference code yields a
float64 "truth" run; `set_precision('float32')` reproduces the reference's dtype.
Dtype strictness follows TF: tensor (op) tensor with different dtypes raises; non-tensor operands adopt the tensor's dtype.
"""
import contextlib
import functools
from collections import namedtuple

import numpy as np
import torch
import torch.nn.functional as F

__version__ = '2.1.2+torch-shim'

# ----------------------------------------------------------------------------------------------------------------- dtypes
_PRECISION = {'f32': torch.float32}


def set_precision(name):
    """'float32' (reference dtype) or 'float64' (truth run: every tf.float32 becomes float64)."""
    assert name in ('float32', 'float64')
    _PRECISION['f32'] = torch.float32 if name == 'float32' else torch.float64


class DType(object):
    def __init__(self, name, torch_dtype):
        self.name, self._t = name, torch_dtype

    @property
    def torch(self):
        return _PRECISION['f32'] if self.name == 'float32' else self._t

    @property
    def base_dtype(self):
        return self

    @property
    def max(self):
        return float(np.finfo(np.float32).max) if self.name == 'float32' else float(torch.finfo(self._t).max) \
            if self._t.is_floating_point else int(torch.iinfo(self._t).max)

    @property
    def as_numpy_dtype(self):
        return {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64,
                torch.bool: np.bool_, torch.uint8: np.uint8}[self.torch]

    def __repr__(self):
        return 'tf.' + self.name

    def __eq__(self, o):
        return isinstance(o, DType) and o.name == self.name

    def __hash__(self):
        return hash(self.name)


float32 = DType('float32', torch.float32)
float64 = DType('float64', torch.float64)
int32 = DType('int32', torch.int32)
int64 = DType('int64', torch.int64)
uint8 = DType('uint8', torch.uint8)
bool = DType('bool', torch.bool)          # noqa: A001  (tf.bool)
_py_bool = (1 == 1).__class__


def _td(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, DType):
        return dtype.torch
    if isinstance(dtype, torch.dtype):
        return dtype
    if isinstance(dtype, str):
        return {'float32': float32, 'float64': float64, 'int32': int32, 'int64': int64}[dtype].torch
    return {np.float32: float32.torch, np.float64: torch.float64, np.int32: torch.int32, np.int64: torch.int64}[dtype]


class InvalidArgumentError(Exception):
    """tf.errors.InvalidArgumentError (dtype mismatches between two tensors)."""


# ----------------------------------------------------------------------------------------------------------------- tensors
class TensorShape(tuple):
    def as_list(self):
        return list(self)

    @property
    def rank(self):
        return len(self)

    def __getitem__(self, i):
        r = tuple.__getitem__(self, i)
        return TensorShape(r) if isinstance(i, slice) else r


_ARITH = {'add', 'sub', 'mul', 'div', 'true_divide', 'pow', 'maximum', 'minimum', 'matmul', 'rsub',
          '__add__', '__radd__', '__sub__', '__rsub__', '__mul__', '__rmul__', '__truediv__', '__rtruediv__', '__pow__',
          '__rpow__', '__matmul__', '__rmatmul__', '__iadd__', '__isub__', '__imul__', '__itruediv__'}


class Tensor(torch.Tensor):
    """Eager tensor: a torch.Tensor with TensorFlow's surface (.numpy() on graph-attached values, .shape.as_list(), TF dtype rules)."""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, '__name__', '')
        if name in _ARITH:
            ref = None
            for a in args:
                if isinstance(a, torch.Tensor):
                    ref = a
                    break
            conv = []
            for a in args:
                if isinstance(a, np.ndarray) or isinstance(a, (list, tuple)) or isinstance(a, np.generic):
                    a = torch.as_tensor(np.asarray(a)).to(ref.dtype)          # non-tensor operands adopt the tensor's dtype (TF binary-op wrapper)
                elif isinstance(a, torch.Tensor) and ref is not None and a.dtype != ref.dtype:
                    raise InvalidArgumentError('tf shim: dtype mismatch in {}: {} vs {}'.format(name, ref.dtype, a.dtype))
                conv.append(a)
            args = tuple(conv)
        with torch._C.DisableTorchFunctionSubclass():
            ret = func(*args, **kwargs)
        if func in torch.overrides.get_default_nowrap_functions():
            return ret
        return torch._tensor._convert(ret, Tensor)          # results (also of arithmetic on Variables) are plain eager tensors

    # -- TensorFlow surface
    def numpy(self):
        return self.detach().as_subclass(torch.Tensor).numpy().copy()

    def __array__(self, dtype=None, copy=None):
        a = self.detach().as_subclass(torch.Tensor).numpy()
        return a.astype(dtype) if dtype is not None else a

    # eager tensors are immutable in TensorFlow: `a += b` rebinds the name to a new tensor and never mutates `a`
    def __iadd__(self, o): return self + o
    def __isub__(self, o): return self - o
    def __imul__(self, o): return self * o
    def __itruediv__(self, o): return self / o

    def __array_wrap__(self, array, context=None, return_scalar=False):
        return array          # NumPy ufuncs on eager tensors give ndarrays (np.isnan(grad) in the workflow's NaN scan)

    @property
    def shape(self):
        return TensorShape(super().shape)

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return super().dtype

    def __bool__(self):
        return _py_bool(self.detach().as_subclass(torch.Tensor).item())

    def __hash__(self):
        return id(self)

    def __repr__(self):
        return 'tf.Tensor(shape={}, dtype={})'.format(tuple(self.shape), super().dtype)


def _wrap(t):
    return t if isinstance(t, Tensor) else t.as_subclass(Tensor)


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, Tensor) else t


def _t(x, dtype=None):
    """convert_to_tensor: numpy float arrays keep THEIR dtype (float64 stays float64) unless a dtype is given; python floats -> float32."""
    if isinstance(x, KerasTensor):
        return x
    if isinstance(x, torch.Tensor):
        r = x
    elif isinstance(x, np.ndarray) or isinstance(x, np.generic):
        a = np.asarray(x)
        r = torch.from_numpy(np.ascontiguousarray(a))
        if a.dtype == np.float32:
            r = r.to(float32.torch)
    elif isinstance(x, (float,)):
        r = torch.tensor(x, dtype=float32.torch)
    elif isinstance(x, (int, _py_bool)):
        r = torch.tensor(x, dtype=torch.int32)
    elif isinstance(x, (list, tuple)):
        if len(x) and any(isinstance(e, torch.Tensor) for e in x):
            r = torch.stack([_raw(_t(e)) for e in x])
        else:
            a = np.asarray(x)
            r = torch.from_numpy(a)
            if a.dtype == np.float64:
                r = r.to(float32.torch)          # python float lists default to float32 in TF
            elif a.dtype == np.int64:
                r = r.to(torch.int32)
    else:
        raise TypeError('tf shim: cannot convert {} to a tensor'.format(type(x)))
    if dtype is not None and r.dtype != _td(dtype):
        r = r.to(_td(dtype))
    return _wrap(r)


class Variable(Tensor):
    """tf.Variable: leaf tensor with assign(); `.name`, `.trainable`."""

    @staticmethod
    def _make(value, name, trainable):
        v = torch.Tensor._make_subclass(Variable, _raw(value).detach().clone(), trainable and value.dtype.is_floating_point)
        v._vname = name + ':0'
        v._trainable = trainable
        return v

    @property
    def name(self):
        return self._vname

    @property
    def trainable(self):
        return self._trainable

    def assign(self, value):
        with torch.no_grad():
            self.copy_(_raw(_t(value)).to(super().dtype).reshape(super().shape))
        return self

    def assign_sub(self, value):
        with torch.no_grad():
            self.sub_(_raw(_t(value)).to(super().dtype))
        return self

    def __repr__(self):
        return '<tf.Variable {} shape={}>'.format(self._vname, tuple(self.shape))


# ----------------------------------------------------------------------------------------------------------------- Keras symbolic tensors
_PROBE = 16          # spatial probe size for Input dimensions declared as None


class KerasTensor(object):
    """Symbolic tensor of the Keras functional API: a deferred call (fn, args, kwargs)[index] plus a concrete probe for shapes."""

    def __init__(self, fn, args, kwargs, index, probe, dyn):
        self._fn, self._args, self._kwargs, self._index, self._probe, self._dyn = fn, args, kwargs, index, probe, dyn

    @property
    def shape(self):
        s = [None] + list(self._probe.shape[1:])
        if self._dyn and len(s) == 4:
            s[1] = s[2] = None
        return TensorShape(s)

    @property
    def dtype(self):
        return self._probe.dtype

    def __repr__(self):
        return '<KerasTensor shape={}>'.format(tuple(self.shape))

    def __add__(self, o): return add(self, o)
    def __radd__(self, o): return add(o, self)
    def __sub__(self, o): return subtract(self, o)
    def __rsub__(self, o): return subtract(o, self)
    def __mul__(self, o): return multiply(self, o)
    def __rmul__(self, o): return multiply(o, self)
    def __truediv__(self, o): return divide(self, o)
    def __rtruediv__(self, o): return divide(o, self)
    def __neg__(self): return multiply(self, -1.0)
    def __getitem__(self, idx): return _getitem(self, idx)


def _find_sym(obj):
    if isinstance(obj, KerasTensor):
        return obj
    if isinstance(obj, (list, tuple)):
        for e in obj:
            r = _find_sym(e)
            if r is not None:
                return r
    if isinstance(obj, dict):
        for e in obj.values():
            r = _find_sym(e)
            if r is not None:
                return r
    return None


def _map_sym(obj, f):
    if isinstance(obj, KerasTensor):
        return f(obj)
    if isinstance(obj, list):
        return [_map_sym(e, f) for e in obj]
    if isinstance(obj, tuple):
        return tuple(_map_sym(e, f) for e in obj)
    if isinstance(obj, dict):
        return {k: _map_sym(v, f) for k, v in obj.items()}
    return obj


def _symbolic_call(fn, args, kwargs):
    sym = _find_sym((args, kwargs))
    with torch.no_grad():
        pa = _map_sym(args, lambda s: s._probe)
        pk = _map_sym(kwargs, lambda s: s._probe)
        out = fn(*pa, **pk)
    dyn = any_dyn((args, kwargs))
    if isinstance(out, (tuple, list)):
        return type(out)(KerasTensor(fn, args, kwargs, i, o, dyn) for i, o in enumerate(out))
    return KerasTensor(fn, args, kwargs, None, out, dyn)


def any_dyn(obj):
    found = []
    _map_sym(obj, lambda s: found.append(s._dyn))
    return any(found)


def op(fn):
    """Decorator: a shim op called on symbolic (Keras functional) tensors records a graph node instead of computing."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if _find_sym((args, kwargs)) is not None:
            return _symbolic_call(fn, args, kwargs)
        return fn(*args, **kwargs)
    return wrapper


def _evaluate(sym, memo):
    key = id(sym)
    if key in memo:
        return memo[key]
    if sym._fn is None:
        raise ValueError('tf shim: symbolic input not fed')
    node_key = ('node', id(sym._args), id(sym._kwargs))
    if node_key not in memo:
        a = _map_sym(sym._args, lambda s: _evaluate(s, memo))
        k = _map_sym(sym._kwargs, lambda s: _evaluate(s, memo))
        memo[node_key] = sym._fn(*a, **k)
    out = memo[node_key]
    val = out if sym._index is None else out[sym._index]
    memo[key] = val
    return val


# ----------------------------------------------------------------------------------------------------------------- basic ops
def shape(x):
    """tf.shape in eager mode: integers (the reference only does integer arithmetic / indexing on the result)."""
    return TensorShape(int(s) for s in _raw(_t(x)).shape)


@op
def convert_to_tensor(value, dtype=None, dtype_hint=None, name=None):
    return _t(value, dtype)


@op
def constant(value, dtype=None, shape=None, name=None):
    """tf.constant: python / numpy value -> tensor; python floats and float64 arrays WITHOUT dtype keep numpy's dtype for arrays
    (tf.constant(np.float64 array) is float64) and become float32 for python scalars/lists."""
    r = _t(value, dtype)
    if shape is not None:
        r = _wrap(_raw(r).reshape(tuple(shape)))
    return r


@op
def cast(x, dtype, name=None):
    return _wrap(_raw(_t(x)).to(_td(dtype)))


@op
def identity(x, name=None):
    return _wrap(_raw(_t(x)) * 1) if _t(x).dtype.is_floating_point else _t(x)


@op
def stop_gradient(x, name=None):
    return _wrap(_raw(_t(x)).detach())


def _bin(a, b):
    """TF binary-op operand conversion: the tensor operand fixes the dtype of a non-tensor operand; two tensors must agree."""
    ta, tb = isinstance(a, torch.Tensor), isinstance(b, torch.Tensor)
    if ta and not tb:
        b = _raw(_t(b)).to(a.dtype)
    elif tb and not ta:
        a = _raw(_t(a)).to(b.dtype)
    elif not ta and not tb:
        a, b = _raw(_t(a)), _raw(_t(b))
    if a.dtype != b.dtype:
        raise InvalidArgumentError('tf shim: dtype mismatch {} vs {}'.format(a.dtype, b.dtype))
    return _raw(a), _raw(b)


@op
def add(a, b, name=None):
    a, b = _bin(a, b)
    return _wrap(a + b)


@op
def subtract(a, b, name=None):
    a, b = _bin(a, b)
    return _wrap(a - b)


@op
def multiply(a, b, name=None):
    a, b = _bin(a, b)
    return _wrap(a * b)


@op
def divide(a, b, name=None):
    a, b = _bin(a, b)
    return _wrap(a / b)


@op
def pow(x, y, name=None):          # noqa: A001
    x, y = _bin(x, y)
    return _wrap(torch.pow(x, y))


@op
def sin(x, name=None):
    return _wrap(torch.sin(_raw(_t(x))))


@op
def exp(x, name=None):
    return _wrap(torch.exp(_raw(_t(x))))


@op
def round(x, name=None):          # noqa: A001
    """tf.round: half to even (cwise_ops.h round_half_to_even); NO gradient is registered... the Round op's gradient is None."""
    return _wrap(torch.round(_raw(_t(x)).detach()))


@op
def reshape(x, shape, name=None):          # noqa: A002
    return _wrap(_raw(_t(x)).reshape([int(s) for s in shape]))


@op
def transpose(x, perm=None, name=None):
    t = _raw(_t(x))
    if perm is None:
        perm = list(range(t.dim()))[::-1]
    return _wrap(t.permute(*[int(p) for p in perm]))


@op
def expand_dims(x, axis, name=None):
    return _wrap(_raw(_t(x)).unsqueeze(int(axis)))


@op
def tile(x, multiples, name=None):
    return _wrap(_raw(_t(x)).repeat(*[int(m) for m in multiples]))


@op
def concat(values, axis, name=None):
    ts = [_raw(_t(v)) for v in values]
    if len({t.dtype for t in ts}) != 1:
        raise TypeError('tf shim: concat of mixed dtypes')
    return _wrap(torch.cat(ts, dim=int(axis)))


@op
def matmul(a, b, name=None):
    a, b = _bin(a, b)
    return _wrap(torch.matmul(a, b))


@op
def gather(params, indices, axis=0, name=None):
    return _wrap(torch.index_select(_raw(_t(params)), int(axis), _raw(_t(indices)).to(torch.int64).reshape(-1)))


@op
def argmax(x, axis=None, output_type=None, name=None):
    return _wrap(torch.argmax(_raw(_t(x)), dim=0 if axis is None else int(axis)))


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return (int(axis),)


@op
def reduce_sum(x, axis=None, keepdims=False, name=None):
    t = _raw(_t(x))
    return _wrap(t.sum() if axis is None else t.sum(dim=_axes(axis), keepdim=keepdims))


@op
def reduce_mean(x, axis=None, keepdims=False, name=None):
    t = _raw(_t(x))
    return _wrap(t.mean() if axis is None else t.mean(dim=_axes(axis), keepdim=keepdims))


@op
def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    """clip_ops.py (TF 2.1): maximum(minimum(t, max), min); Minimum/Maximum gradients use <= / >= masks, i.e. the gradient passes
    wherever min <= t <= max (inclusive) and is zero outside."""
    x = _raw(_t(t))
    lo = _raw(_t(clip_value_min)).to(x.dtype) if not isinstance(clip_value_min, torch.Tensor) else _raw(clip_value_min)
    hi = _raw(_t(clip_value_max)).to(x.dtype) if not isinstance(clip_value_max, torch.Tensor) else _raw(clip_value_max)
    mask = ((x <= hi) & (torch.minimum(x, hi) >= lo)).to(x.dtype)
    val = torch.maximum(torch.minimum(x.detach(), hi), lo)
    return _wrap(val + (x - x.detach()) * mask)


@op
def _getitem(x, idx):
    return _wrap(_raw(x)[idx])


def _pad_index(n, before, after, mode):
    idx = np.arange(-before, n + after)
    if mode == 'REFLECT':          # mirror without repeating the edge sample
        period = 2 * (n - 1) if n > 1 else 1
        idx = np.abs(idx) % period
        idx = np.where(idx >= n, period - idx, idx)
    else:                          # SYMMETRIC: mirror repeating the edge sample
        period = 2 * n
        idx = np.where(idx < 0, -idx - 1, idx) % period
        idx = np.where(idx >= n, period - 1 - idx, idx)
    return torch.from_numpy(idx.astype(np.int64))


@op
def pad(tensor, paddings, mode='CONSTANT', constant_values=0, name=None):
    """tf.pad (mirror_pad_op.h for REFLECT / SYMMETRIC)."""
    x = _raw(_t(tensor))
    p = np.asarray(_raw(paddings).numpy() if isinstance(paddings, torch.Tensor) else paddings).astype(int).reshape(-1, 2)
    mode = mode.upper()
    if mode == 'CONSTANT':
        flat = []
        for b, a in p[::-1]:
            flat += [int(b), int(a)]
        return _wrap(F.pad(x, flat, mode='constant', value=float(constant_values)))
    for d, (b, a) in enumerate(p):
        if b or a:
            x = torch.index_select(x, d, _pad_index(x.shape[d], int(b), int(a), mode))
    return _wrap(x)


# ----------------------------------------------------------------------------------------------------------------- tf.math
class _Namespace(object):
    pass


math_ns = _Namespace()


@op
def _math_pow(x, y, name=None):
    x, y = _bin(x, y)
    return _wrap(torch.pow(x, y))


@op
def _math_abs(x, name=None):
    return _wrap(torch.abs(_raw(_t(x))))


@op
def _math_log(x, name=None):
    return _wrap(torch.log(_raw(_t(x))))


@op
def _reduce_std(x, axis=None, keepdims=False, name=None):
    t = _raw(_t(x))
    m = t.mean() if axis is None else t.mean(dim=_axes(axis), keepdim=True)
    v = ((t - m) ** 2)
    v = v.mean() if axis is None else v.mean(dim=_axes(axis), keepdim=keepdims)
    return _wrap(torch.sqrt(v))


math_ns.pow, math_ns.abs, math_ns.log, math_ns.reduce_std = _math_pow, _math_abs, _math_log, _reduce_std
math_ns.reduce_mean, math_ns.reduce_sum, math_ns.round, math_ns.sin, math_ns.exp = reduce_mean, reduce_sum, round, sin, exp
abs = _math_abs          # noqa: A001


# ----------------------------------------------------------------------------------------------------------------- tf.nn
nn = _Namespace()


def _same_pads(in_size, k, s):
    """TensorFlow SAME: out = ceil(in / s); total = max((out-1)*s + k - in, 0); before = total // 2 (extra goes after)."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2


def _strides2(strides):
    if isinstance(strides, int):
        return strides, strides
    s = list(strides)
    if len(s) == 4:
        return int(s[1]), int(s[2])
    if len(s) == 2:
        return int(s[0]), int(s[1])
    return int(s[0]), int(s[0])


@op
def _conv2d(input, filters, strides, padding, data_format='NHWC', dilations=None, name=None):          # noqa: A002
    """tf.nn.conv2d, NHWC activations x HWIO filters (cross-correlation)."""
    x = _raw(_t(input))
    w = _raw(_t(filters))
    if w.dtype != x.dtype:
        if isinstance(filters, torch.Tensor):
            raise TypeError('tf shim: conv2d dtype mismatch {} vs {}'.format(x.dtype, w.dtype))
        w = w.to(x.dtype)
    sh, sw = _strides2(strides)
    xn = x.permute(0, 3, 1, 2)
    if padding.upper() == 'SAME':
        pt, pb = _same_pads(x.shape[1], w.shape[0], sh)
        pl, pr = _same_pads(x.shape[2], w.shape[1], sw)
        xn = F.pad(xn, [pl, pr, pt, pb])
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), stride=(sh, sw))
    return _wrap(y.permute(0, 2, 3, 1))


@op
def _space_to_depth(input, block_size, name=None):          # noqa: A002
    """out[n, i, j, (di*b + dj)*C + c] = in[n, i*b + di, j*b + dj, c] (spacetodepth_op.cc, NHWC)."""
    x = _raw(_t(input))
    n, h, w, c = x.shape
    b = int(block_size)
    y = x.reshape(n, h // b, b, w // b, b, c).permute(0, 1, 3, 2, 4, 5)
    return _wrap(y.reshape(n, h // b, w // b, b * b * c))


@op
def _depth_to_space(input, block_size, name=None):          # noqa: A002
    x = _raw(_t(input))
    n, h, w, c = x.shape
    b = int(block_size)
    co = c // (b * b)
    y = x.reshape(n, h, w, b, b, co).permute(0, 1, 3, 2, 4, 5)
    return _wrap(y.reshape(n, h * b, w * b, co))


@op
def _leaky_relu(features, alpha=0.2, name=None):
    x = _raw(_t(features))
    return _wrap(torch.where(x > 0, x, x * alpha))


@op
def _l2_loss(t, name=None):
    x = _raw(_t(t))
    return _wrap((x * x).sum() / 2)


@op
def _avg_pool(input, ksize, strides, padding, name=None):          # noqa: A002
    """tf.nn.avg_pool: SAME padding cells are excluded from the divisor (avgpooling_op / Eigen SpatialAvgPooling)."""
    x = _raw(_t(input)).permute(0, 3, 1, 2)
    kh, kw = _strides2(ksize)
    sh, sw = _strides2(strides)
    if padding.upper() == 'SAME':
        pt, pb = _same_pads(x.shape[2], kh, sh)
        pl, pr = _same_pads(x.shape[3], kw, sw)
        ones = F.pad(torch.ones_like(x[:1, :1]), [pl, pr, pt, pb])
        x = F.pad(x, [pl, pr, pt, pb])
        cnt = F.avg_pool2d(ones, (kh, kw), (sh, sw)) * (kh * kw)
        y = F.avg_pool2d(x, (kh, kw), (sh, sw)) * (kh * kw) / cnt
    else:
        y = F.avg_pool2d(x, (kh, kw), (sh, sw))
    return _wrap(y.permute(0, 2, 3, 1))


def _max_pool_nhwc(x, pool, strides, padding):
    xn = x.permute(0, 3, 1, 2)
    kh, kw = pool
    sh, sw = strides
    if padding.upper() == 'SAME':
        pt, pb = _same_pads(x.shape[1], kh, sh)
        pl, pr = _same_pads(x.shape[2], kw, sw)
        xn = F.pad(xn, [pl, pr, pt, pb], value=float('-inf'))
    return F.max_pool2d(xn, (kh, kw), (sh, sw)).permute(0, 2, 3, 1)


_TopK = namedtuple('TopKV2', ['values', 'indices'])


@op
def _top_k(input, k=1, sorted=True, name=None):          # noqa: A002
    v, i = torch.topk(_raw(_t(input)), int(k), dim=-1, largest=True, sorted=True)
    return _TopK(_wrap(v), _wrap(i))


@op
def _softmax(x, axis=-1):
    return _wrap(torch.softmax(_raw(_t(x)), dim=axis))


nn.conv2d, nn.space_to_depth, nn.depth_to_space, nn.leaky_relu = _conv2d, _space_to_depth, _depth_to_space, _leaky_relu
nn.l2_loss, nn.avg_pool, nn.top_k, nn.softmax = _l2_loss, _avg_pool, _top_k, _softmax


# ----------------------------------------------------------------------------------------------------------------- tf.image
image = _Namespace()


def _interp_weights(out_size, in_size):
    """resize_bilinear_op.cc compute_interpolation_weights with HalfPixelScaler, float32 arithmetic as in the kernel."""
    scale = np.float32(in_size) / np.float32(out_size)
    i = np.arange(out_size, dtype=np.float32)
    src = (i + np.float32(0.5)) * scale - np.float32(0.5)
    fl = np.floor(src)
    lower = np.maximum(fl, 0).astype(np.int64)
    upper = np.minimum(np.ceil(src), in_size - 1).astype(np.int64)
    lerp = (src - fl).astype(np.float32)
    return torch.from_numpy(lower), torch.from_numpy(upper), torch.from_numpy(lerp.astype(np.float64))


@op
def _resize(images, size, method='bilinear', preserve_aspect_ratio=False, antialias=False, name=None):
    """tf.image.resize (v2): half-pixel centres, no antialiasing; only bilinear is used by the reference."""
    if method != 'bilinear':
        raise NotImplementedError('tf shim: resize method ' + str(method))
    x = _raw(_t(images))
    oh, ow = int(size[0]), int(size[1])
    ylo, yhi, yl = _interp_weights(oh, x.shape[1])
    xlo, xhi, xl = _interp_weights(ow, x.shape[2])
    yl = yl.to(x.dtype).reshape(1, -1, 1, 1)
    xl = xl.to(x.dtype).reshape(1, 1, -1, 1)
    top, bot = torch.index_select(x, 1, ylo), torch.index_select(x, 1, yhi)
    tl, tr = torch.index_select(top, 2, xlo), torch.index_select(top, 2, xhi)
    bl, br = torch.index_select(bot, 2, xlo), torch.index_select(bot, 2, xhi)
    t = tl + (tr - tl) * xl
    b = bl + (br - bl) * xl
    return _wrap(t + (b - t) * yl)


@op
def _rgb_to_hsv(images, name=None):
    """colorspace_op.h RGBToHSV. Registered ops.NotDifferentiable('RGBToHSV') in TF 2.1 (python/ops/image_ops_impl.py): no gradient."""
    x = _raw(_t(images)).detach()
    r, g, b = x[..., 0], x[..., 1], x[..., 2]
    v = torch.amax(x, dim=-1)
    rng = v - torch.amin(x, dim=-1)
    s = torch.where(v > 0, rng / torch.where(v > 0, v, torch.ones_like(v)), torch.zeros_like(v))
    norm = 1.0 / (6.0 * torch.where(rng > 0, rng, torch.ones_like(rng)))
    h = torch.where(r == v, norm * (g - b), torch.where(g == v, norm * (b - r) + 2.0 / 6.0, norm * (r - g) + 4.0 / 6.0))
    h = torch.where(rng > 0, h, torch.zeros_like(h))
    h = torch.where(h < 0, h + 1.0, h)
    return _wrap(torch.stack((h, s, v), dim=-1))


@op
def _hsv_to_rgb(images, name=None):
    """colorspace_op.h HSVToRGB. ops.NotDifferentiable('HSVToRGB'): no gradient."""
    x = _raw(_t(images)).detach()
    h, s, v = x[..., 0], x[..., 1], x[..., 2]
    dh = h * 6.0
    dr = torch.clamp(torch.abs(dh - 3.0) - 1.0, 0.0, 1.0)
    dg = torch.clamp(-torch.abs(dh - 2.0) + 2.0, 0.0, 1.0)
    db = torch.clamp(-torch.abs(dh - 4.0) + 2.0, 0.0, 1.0)
    one_s = -s + 1.0
    return _wrap(torch.stack(((one_s + s * dr) * v, (one_s + s * dg) * v, (one_s + s * db) * v), dim=-1))


@op
def _extract_patches(images, sizes, strides, rates, padding, name=None):
    """tf.image.extract_patches, VALID, stride 1: depth index = (kr * kw + kc) * C + c."""
    x = _raw(_t(images))
    kh, kw = int(sizes[1]), int(sizes[2])
    if padding.upper() != 'VALID' or list(strides) != [1, 1, 1, 1]:
        raise NotImplementedError
    n, h, w, c = x.shape
    cols = [x[:, i:h - kh + 1 + i, j:w - kw + 1 + j, :] for i in range(kh) for j in range(kw)]
    return _wrap(torch.cat(cols, dim=3))


def _fspecial_gauss(size, sigma, dtype):
    """image_ops_impl._fspecial_gauss: softmax of -(x^2 + y^2) / (2 sigma^2)."""
    coords = torch.arange(size, dtype=dtype) - (size - 1) / 2.0
    g = coords ** 2 * (-0.5 / (sigma * sigma))
    g2 = (g.reshape(1, -1) + g.reshape(-1, 1)).reshape(1, -1)
    return torch.softmax(g2, dim=-1).reshape(size, size)


def _ssim_per_channel(a, b, max_val, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03):
    """image_ops_impl._ssim_per_channel / _ssim_helper (compensation = 1.0), VALID depthwise Gaussian windows."""
    c = a.shape[-1]
    kern = _fspecial_gauss(filter_size, filter_sigma, a.dtype).reshape(1, 1, filter_size, filter_size).repeat(c, 1, 1, 1)

    def red(t):
        return F.conv2d(t.permute(0, 3, 1, 2), kern, groups=c).permute(0, 2, 3, 1)
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    mean0, mean1 = red(a), red(b)
    num0 = mean0 * mean1 * 2.0
    den0 = mean0 * mean0 + mean1 * mean1
    lum = (num0 + c1) / (den0 + c1)
    num1 = red(a * b) * 2.0
    den1 = red(a * a + b * b)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    return (lum * cs).mean(dim=(1, 2)), cs.mean(dim=(1, 2))


@op
def _ssim(img1, img2, max_val, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03):
    a, b = _bin(_t(img1), _t(img2))
    s, _ = _ssim_per_channel(a, b, float(max_val), filter_size, filter_sigma, k1, k2)
    return _wrap(s.mean(dim=-1))


_MSSSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


@op
def _ssim_multiscale(img1, img2, max_val, power_factors=_MSSSIM_WEIGHTS, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03):
    """image_ops_impl.ssim_multiscale: per scale relu(cs) (last scale relu(ssim)), 2x2 SAME average pooling between scales with
    an odd-size remainder pad, product of powers, mean over channels."""
    a, b = _bin(_t(img1), _t(img2))
    mcs = []
    for k in range(len(power_factors)):
        if k > 0:
            rem_h, rem_w = a.shape[1] % 2, a.shape[2] % 2
            if rem_h or rem_w:
                a = _raw(pad(_wrap(a), [[0, 0], [0, rem_h], [0, rem_w], [0, 0]], 'SYMMETRIC'))
                b = _raw(pad(_wrap(b), [[0, 0], [0, rem_h], [0, rem_w], [0, 0]], 'SYMMETRIC'))
            a = F.avg_pool2d(a.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
            b = F.avg_pool2d(b.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
        s, cs = _ssim_per_channel(a, b, float(max_val), filter_size, filter_sigma, k1, k2)
        mcs.append(torch.relu(cs))
    mcs.pop()
    stack = torch.stack(mcs + [torch.relu(s)], dim=-1)
    w = torch.tensor(power_factors, dtype=a.dtype)
    return _wrap(torch.prod(stack ** w, dim=-1).mean(dim=-1))


image.resize, image.rgb_to_hsv, image.hsv_to_rgb, image.extract_patches = _resize, _rgb_to_hsv, _hsv_to_rgb, _extract_patches
image.ssim, image.ssim_multiscale = _ssim, _ssim_multiscale


# ----------------------------------------------------------------------------------------------------------------- tf.random
random = _Namespace()
random.noise_log = []          # every tf.random.normal draw is recorded so a golden file can carry the noise that was used
random.generator = torch.Generator().manual_seed(1234)


def _normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):          # noqa: A002
    n = torch.randn([int(s) for s in shape], generator=random.generator, dtype=torch.float64).to(_td(dtype)) * stddev + mean
    random.noise_log.append(n.numpy().copy())
    return _wrap(n)


random.normal = _normal
random.set_seed = lambda s: random.generator.manual_seed(int(s))


# ----------------------------------------------------------------------------------------------------------------- misc
@contextlib.contextmanager
def name_scope(name, *a, **k):
    yield name


def function(f=None, **kwargs):
    """tf.function: tracing compiler — semantics of the eager function are unchanged."""
    if f is None:
        return lambda g: g
    return f


class GradientTape(object):
    """tf.GradientTape over torch autograd: variables are leaves; tape.gradient = torch.autograd.grad (None for unconnected sources)."""

    last = None

    def __init__(self, persistent=False, watch_accessed_variables=True):
        self._persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, t):
        pass

    def gradient(self, target, sources, output_gradients=None):
        single = isinstance(sources, torch.Tensor)
        src = [sources] if single else list(sources)
        g = torch.autograd.grad(_raw(target), [s for s in src], allow_unused=True, retain_graph=self._persistent)
        g = [None if e is None else _wrap(e.detach()) for e in g]
        GradientTape.last = (src, g)          # kept so a golden generator can read the gradients a reference training_step computed
        return g[0] if single else g


class _Config(object):
    @staticmethod
    def list_physical_devices(kind=None):
        return []

    @staticmethod
    def set_visible_devices(devices, kind=None):
        return None


config = _Config()


class constant_initializer(object):
    def __init__(self, value=0):
        self.value = value

    def __call__(self, shape, dtype=None):          # noqa: A002
        v = np.asarray(self.value, dtype=np.float64)
        return torch.from_numpy(np.broadcast_to(v.reshape(shape) if v.size == int(np.prod(shape)) and v.size > 1 else v, shape).copy())


class zeros_initializer(object):
    def __call__(self, shape, dtype=None):          # noqa: A002
        return torch.zeros(shape, dtype=torch.float64)


class ones_initializer(object):
    def __call__(self, shape, dtype=None):          # noqa: A002
        return torch.ones(shape, dtype=torch.float64)


from . import keras          # noqa: E402,F401
math = math_ns          # noqa: A001  (tf.math — shadows the stdlib module inside this package on purpose, keep last)
