"""Hyper-parameter specification / validation with the semantics of the reference's ParamSpec
(helpers/paramspec.py:33-178): specs are ``name -> (default, dtype, validator)`` where the validator is a
``(min, max)`` tuple, a set of allowed values, a sub-string (for str) or a callable; values are immutable
attributes set only through ``update``; ``to_json`` stringifies non-numeric values."""
import numpy as np

from . import utils


def numbers_in_range(dtype, min_value=None, max_value=None):
    def check(items):
        return all(isinstance(i, dtype) and (min_value is None or i >= min_value) and (max_value is None or i <= max_value)
                   for i in items)
    return check


def item_passes(check):
    return lambda items: all(check(i) for i in items)


class ParamSpec(object):

    def __init__(self, specs):
        self._check_specs(specs)
        object.__setattr__(self, '_specs', dict(specs))
        object.__setattr__(self, '_values', {})

    @staticmethod
    def _check_specs(specs):
        for key, spec in specs.items():
            if not isinstance(spec, tuple) or len(spec) != 3:
                raise ValueError('Invalid parameter specification for key {} - expected tuple of length 3'.format(key))
            _, dtype, rule = spec
            if rule is None:
                continue
            if dtype is str and not (isinstance(rule, (str, set)) or callable(rule)):
                raise ValueError('String data types can be validated by a regex (string), enum (set) or custom function')
            if utils.is_numeric_type(dtype) and not isinstance(rule, (tuple, set)):
                raise ValueError('Numeric data types can be validated by a range (2-elem tuple), or enum (set)')

    def add(self, specs):
        self._check_specs(specs)
        self._specs.update(specs)

    def __getattr__(self, name):
        values, specs = object.__getattribute__(self, '_values'), object.__getattribute__(self, '_specs')
        if name in values:
            return values[name]
        if name in specs:
            return specs[name][0]
        raise KeyError(name)

    def __setattr__(self, key, value):
        raise ValueError('Values cannot be set directly. Use the `update` method.')

    def __contains__(self, item):
        return item in self._specs

    def keys(self):
        return list(self._specs.keys())

    def get_dtype(self, name):
        return self._specs[name][1]

    def get_default(self, name):
        return self._specs[name][0]

    def get_value(self, name):
        return getattr(self, name)

    def _rule(self, name, kind):
        rule = self._specs[name][2]
        return rule if isinstance(rule, kind) else None

    def get_min(self, name):
        r = self._rule(name, tuple)
        return r[0] if r is not None and len(r) == 2 else None

    def get_max(self, name):
        r = self._rule(name, tuple)
        return r[1] if r is not None and len(r) == 2 else None

    def get_enum(self, name):
        r = self._rule(name, set)
        return set(r) if r is not None else None

    def get_regex(self, name):
        return self._rule(name, str)

    def to_dict(self):
        out = {k: spec[0] for k, spec in self._specs.items()}
        out.update(self._values)
        return out

    def to_json(self):
        return {k: (v if utils.is_number(v) else str(v)) for k, v in self.to_dict().items()}

    def changed_params(self):
        return {k: v for k, v in self._values.items() if self._specs[k][0] != v}

    def __repr__(self):
        return '{}({})'.format(type(self).__name__, self.to_dict())

    def update(self, **params):
        for key, value in params.items():
            if key not in self._specs:
                raise ValueError('Unexpected parameter: {}!'.format(key))
            if value is None:
                continue
            _, dtype, rule = self._specs[key]
            if utils.is_number(value) and np.isnan(value):
                raise ValueError('Invalid value {} for attribute {}'.format(value, key))
            cand = value if dtype is None else dtype(value)
            if rule is not None:
                if isinstance(rule, tuple) and len(rule) == 2:
                    if rule[0] is not None and cand < rule[0]:
                        raise ValueError('{}: {} fails minimum validation check >= {}!'.format(key, cand, rule[0]))
                    if rule[1] is not None and cand > rule[1]:
                        raise ValueError('{}: {} fails maximum validation check (<= {})!'.format(key, cand, rule[1]))
                elif isinstance(rule, set):
                    if cand not in rule:
                        raise ValueError('{}: {} is not an allowed value ({})!'.format(key, cand, rule))
                elif isinstance(rule, str) and dtype is str:
                    if rule not in cand:
                        raise ValueError('{}: {} does not match regex ({})!'.format(key, cand, rule))
                elif callable(rule):
                    if not rule(cand):
                        raise ValueError('{}: {} failed custom validation check!'.format(key, cand))
            self._values[key] = cand
