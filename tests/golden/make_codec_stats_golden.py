"""Generate tests/golden/codec_stats.npz by EXECUTING the reference's own NumPy helpers that the codec statistics and the DCN training loop
use: helpers/stats.py (bin_edges, hist, entropy — compression/codec.py:44, training/compression.py:232) and helpers/image.py
(batch_gamma — training/compression.py:201), function sources extracted with ast. Runs only where /root/reference is mounted."""
import ast
import os

import numpy as np

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


def load_functions(path, names):
    tree = ast.parse(open(path).read())
    ns = {'np': np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, 'exec'), ns)
    return ns


def main():
    st = load_functions(os.path.join(REF, 'helpers/stats.py'), {'bin_edges', 'hist', 'entropy'})
    im = load_functions(os.path.join(REF, 'helpers/image.py'), {'batch_gamma'})
    rs = np.random.RandomState(1234)
    out = {}
    books = [np.arange(-15, 17, dtype=np.float32), np.arange(-7, 9, dtype=np.float32), np.linspace(-1, 1, 7).astype(np.float32)]
    for i, cb in enumerate(books):
        for j, scale in enumerate((0.3, 1.5, 6.0)):
            z = (rs.normal(size=(2, 8, 8, 4)) * scale).astype(np.float32)
            out['z_%d_%d' % (i, j)] = z
            out['cb_%d_%d' % (i, j)] = cb
            out['entropy_%d_%d' % (i, j)] = np.float64(st['entropy'](z, cb))
    x = rs.uniform(size=(3, 4, 4, 3)).astype(np.float32)
    out['gamma_in'] = x
    out['gamma_out_2p0'] = im['batch_gamma'](x, 2.0)
    g = np.array([[[[0.5]]], [[[1.0]]], [[[2.5]]]], dtype=np.float32)
    out['gamma_vec'] = g
    out['gamma_out_vec'] = im['batch_gamma'](x, g)
    np.savez_compressed(os.path.join(OUT, 'codec_stats.npz'), **out)
    print('wrote', len(out), 'arrays')


if __name__ == '__main__':
    main()
