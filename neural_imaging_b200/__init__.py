"""neural_imaging_b200 — B200-native (sm_100a) implementation of the pkorus/neural-imaging end-to-end patch
pipeline behind the reference's own model / operator API.

    import neural_imaging_b200 as ni
    ni.install_aliases()                     # optional: expose `models`, `helpers`, `workflows`, ... top-level names
    from models import jpeg                  # -> neural_imaging_b200.models.jpeg

Sub-packages mirror the reference layout: models/ (jpeg, pipelines, forensics, compression, layers, tfmodel),
helpers/ (tf_helpers, kernels, paramspec, utils, dataset, metrics), compression/ (jpeg_helpers, codec), pyfse/, workflows/, training/.
"""
import importlib
import sys

__version__ = '0.1.0'
_ALIASES = ('models', 'helpers', 'workflows', 'compression', 'training', 'pyfse')


def install_aliases():
    """Register the reference's top-level package names so reference scripts import this implementation."""
    for name in _ALIASES:
        sys.modules.setdefault(name, importlib.import_module(__name__ + '.' + name))


def lib():
    from . import _lib
    return _lib.lib()
