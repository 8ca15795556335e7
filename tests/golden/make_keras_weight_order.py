"""Generate tests/golden/keras_weight_order.json: the Keras variables (name, shape) of the reference's models in `model.weights`
order, obtained by EXECUTING the reference's own model constructors (models/pipelines.py, models/forensics.py,
models/compression.py) through the tests-only TensorFlow stand-in (tests/tf_shim). It pins the variable ORDER and the Keras storage
SHAPES (Dense (in, out), Conv2DTranspose (2, 2, cout, cin)) that `models/tfmodel.py: save_weights_h5 / load_weights_h5` rely on —
Keras' h5 loader matches variables by order. Runs only in the build container (needs /root/reference). TEST INFRASTRUCTURE.

Limits: the stand-in orders the layers of a functional model by its own graph walk; for the models on the path every weighted layer
lies on the longest input -> output chain, where Keras' depth ordering gives the same sequence (DNet's parallel up-sampling branch is
the one place this is an assumption). ClassicISP is a sub-classed model whose sub-layers are built lazily and is left out.

    python tests/golden/make_keras_weight_order.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'tf_shim'))
sys.path.insert(1, '/root/reference')

import scipy.signal  # noqa: E402
import scipy.signal.windows  # noqa: E402

np.bool, np.float, np.int = bool, float, int          # NumPy-1.18 aliases the reference uses
scipy.signal.gaussian = scipy.signal.windows.gaussian

import tensorflow as tf  # noqa: E402,F401  (the shim)
from models import compression, forensics, pipelines  # noqa: E402

CASES = [('UNet', pipelines.UNet, {}), ('INet', pipelines.INet, {}), ('DNet', pipelines.DNet, {}),
         ('FAN', forensics.FAN, dict(n_classes=5)), ('FAN_dense2', forensics.FAN, dict(n_classes=3, n_dense=2)),
         ('TwitterDCN', compression.TwitterDCN, {})]

out = {}
for key, ctor, kw in CASES:
    m = ctor(**kw)
    out[key] = {'kwargs': kw, 'variables': [[v.name.split('/', 1)[1], list(int(s) for s in v.shape)] for v in m._model.weights]}
with open(os.path.join(HERE, 'keras_weight_order.json'), 'w') as f:
    f.write('{\n' + ',\n'.join(' {}: {}'.format(json.dumps(k), json.dumps(v)) for k, v in out.items()) + '\n}\n')
print({k: len(v['variables']) for k, v in out.items()})
