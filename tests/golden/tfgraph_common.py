"""Shared by tests/golden/make_tf_graph_golden.py (the generator, runs the reference through tests/tf_shim) and by the parity tests that
consume tests/golden/tf_graph_golden.npz: deterministic weights in the product's storage layout, layout converters between Keras
variables and that layout, and the sampling scheme for tensors too large to commit. TEST INFRASTRUCTURE."""
import numpy as np

SAMPLES = 512

# parameters whose reference initial value is a constant the reference code computes (helpers/kernels.py, models/layers.py:40 ...):
# golden weights = that constant + seeded noise, and the generator asserts constant == the product's own initial value
CONST_INIT = ('demosaicing/kernel', 'srgb/kernel', 'gamma_d1/kernel', 'gamma_d1/bias', 'gamma_d2/kernel', 'gamma_d2/bias',
              'demosaicing/alpha', 'constrained_conv2d/kernel', 'encoder/discrete_latent/latent_scaling')


def specs_of(model):
    """[(name, shape, trainable, init)] of a product model built with neural_imaging_b200.nn.HOST_ONLY = True (no device needed)."""
    return [(p.name, tuple(p.shape), bool(p.trainable), np.array(p.init, dtype=np.float32).reshape(tuple(p.shape))) for p in model._store.params]


def golden_state(specs, seed, ones_names=()):
    """Deterministic float32 weights in the product layout: Glorot-scaled uniform kernels, biases in (-0.1, 0.1); parameters the
    reference initialises with constants keep that constant plus noise; frozen parameters keep their constant exactly."""
    rs = np.random.RandomState(seed)
    state = {}
    for name, shape, trainable, init in specs:
        if not trainable:
            state[name] = init.copy()
            continue
        if len(shape) >= 2:
            rec = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            lim = np.sqrt(6.0 / (shape[-2] * rec + shape[-1] * rec))
            noise = rs.uniform(-lim, lim, size=shape)
        elif len(shape) == 1:
            noise = rs.uniform(-0.1, 0.1, size=shape)
        else:
            noise = np.asarray(rs.uniform(-0.1, 0.1))
        if name in CONST_INIT or name in ones_names:
            state[name] = (init + 0.3 * noise).astype(np.float32)
        else:
            state[name] = noise.astype(np.float32)
    return state


def keras_to_product(a, product_shape):
    """Keras variable (or its gradient) -> product layout (neural_imaging_b200/nn.py docstring): Dense (in, out) -> (1, 1, in, out);
    Conv2DTranspose 2x2 (a, b, f, ci) -> 1x1 conv (1, 1, ci, (a*2+b)*F + f); everything else unchanged."""
    a = np.asarray(a)
    ps = tuple(product_shape)
    if a.shape == ps:
        return a
    if a.ndim == 2 and ps == (1, 1) + a.shape:
        return a.reshape(ps)
    if a.ndim == 4 and a.shape[:2] == (2, 2) and ps == (1, 1, a.shape[3], 4 * a.shape[2]):
        return np.ascontiguousarray(a.transpose(3, 0, 1, 2)).reshape(ps)
    raise ValueError('no layout rule {} -> {}'.format(a.shape, ps))


def product_to_keras(a, keras_shape):
    a = np.asarray(a)
    ks = tuple(keras_shape)
    if a.shape == ks:
        return a
    if len(ks) == 2 and a.shape == (1, 1) + ks:
        return a.reshape(ks)
    if len(ks) == 4 and ks[:2] == (2, 2) and a.shape == (1, 1, ks[3], 4 * ks[2]):
        return np.ascontiguousarray(a.reshape(ks[3], 2, 2, ks[2]).transpose(1, 2, 3, 0))
    raise ValueError('no layout rule {} -> {}'.format(a.shape, ks))


def sample_index(size, key):
    """Positions kept of a flattened tensor of `size` elements (all of them when small)."""
    if size <= 4 * SAMPLES:
        return np.arange(size)
    seed = (sum(ord(c) * (i + 1) for i, c in enumerate(key)) * 2654435761 + size) % (2 ** 31 - 1)
    return np.sort(np.random.RandomState(seed).choice(size, SAMPLES, replace=False))


def summarize(a, key):
    """What the fixture keeps of tensor `a`: sampled values, L2 norm and sum (float64)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    idx = sample_index(a.size, key)
    return {'v': a[idx], 'n': np.array([np.sqrt(np.sum(a * a)), np.sum(a), float(a.size)])}


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30)) if a.size else 0.0


class Golden(object):
    """Reader of tf_graph_golden.npz: g.get(case, tensor) -> sampled float64 truth; g.compare(case, tensor, got) -> errors."""

    def __init__(self, path):
        with np.load(path, allow_pickle=False) as d:
            self.d = {k: d[k] for k in d.files}

    def has(self, case, name):
        return '{}/{}/v'.format(case, name) in self.d

    def names(self, case, prefix=''):
        pre = '{}/{}'.format(case, prefix)
        return sorted({k[len(case) + 1:-2] for k in self.d if k.startswith(pre) and k.endswith('/v')})

    def get(self, case, name):
        return self.d['{}/{}/v'.format(case, name)]

    def scalar(self, case, name):
        return float(self.d['{}/{}/v'.format(case, name)].reshape(-1)[0])

    def drift(self, case, name, robust=False):
        """scale-relative distance of the reference's float32 run from its float64 run on the kept samples (max, or 98 % quantile)."""
        return float(self.d['{}/{}/d'.format(case, name)][1 if robust else 0])

    def compare(self, case, name, got):
        """(sample error, norm error), both scale-relative, of a full tensor `got` against the fixture."""
        key = '{}/{}'.format(case, name)
        g = np.asarray(got, dtype=np.float64).reshape(-1)
        n = self.d[key + '/n']
        assert g.size == int(n[2]), '{}: size {} vs golden {}'.format(key, g.size, int(n[2]))
        idx = sample_index(g.size, key)
        e = rel(g[idx], self.d[key + '/v'])
        en = abs(np.sqrt(np.sum(g * g)) - n[0]) / max(n[0], 1e-30)
        return e, float(en)

    def check(self, case, name, got, tol=1e-5, slack=4.0, outliers=0.0, loose=None, what=None):
        """got must match the executed-reference float64 truth within tol, or slack x the reference's own float32 drift.
        `outliers` > 0: piecewise-continuous quantities (a LeakyReLU / max-pool / clip / rounding decision can flip on a value that sits
        within float32 noise of its threshold, in ANY float32 evaluation including the reference's own): at most that fraction of the
        kept samples may exceed the bound (then built from the 98 % quantile of the reference's drift), and those must stay within `loose`."""
        key = '{}/{}'.format(case, name)
        g = np.asarray(got, dtype=np.float64).reshape(-1)
        n = self.d[key + '/n']
        assert g.size == int(n[2]), '{}: size {} vs golden {}'.format(key, g.size, int(n[2]))
        ref = self.d[key + '/v']
        scale = max(float(np.max(np.abs(ref))), 1e-30)
        err = np.abs(g[sample_index(g.size, key)] - ref) / scale
        bound = max(tol, slack * self.drift(case, name, robust=outliers > 0.0))
        frac = float(np.mean(err > bound))
        worst = float(err.max()) if err.size else 0.0
        ok = frac <= outliers and (outliers == 0.0 or loose is None or worst <= loose)
        assert ok, '{}{}: {:.2%} of samples beyond {:.1e} (allowed {:.2%}), worst {:.3e}'.format(
            key, ' (' + what + ')' if what else '', frac, bound, outliers, worst)
        REPORT[key] = max(REPORT.get(key, 0.0), worst if outliers == 0.0 else float(np.quantile(err, 1.0 - outliers)) if err.size else 0.0)
        return worst


REPORT = {}          # achieved errors of the current process, written to profiles/parity_report.json by the GPU test-suite
