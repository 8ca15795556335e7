"""Small host-side helpers (reference helpers/utils.py:53-75: is_number / is_numeric_type / is_nan)."""
import numpy as np

_NUMERIC = (int, float, bool, np.bool_, np.floating, np.integer)


def is_number(value):
    return isinstance(value, _NUMERIC) and not isinstance(value, (str, bytes))


def is_numeric_type(t):
    try:
        return issubclass(t, _NUMERIC)
    except TypeError:
        return False


def is_nan(value):
    if value is None:
        return True
    return bool(np.isnan(value)) if is_number(value) else False


def format_patch_shape(shape):
    if shape is None:
        return '(?)'
    return '({})'.format(','.join('?' if s is None else str(s) for s in shape))


def join_args(d):
    return ','.join('{}={}'.format(k, v) for k, v in d.items())
