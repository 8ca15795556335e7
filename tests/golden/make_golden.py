"""Generate tests/golden/reference_numpy.npz by EXECUTING the reference's own pure-NumPy functions.

Runs only in the build container (needs /root/reference); the GPU box uses the committed .npz.
TensorFlow is not installable here, so only the TF-free fragments can be executed:
  compression/jpeg_helpers.py : zigzag, jpeg_qtable, jpeg_qf_estimation   (function sources extracted with ast)
  helpers/kernels.py          : whole module (scipy.signal.gaussian was removed from SciPy; shimmed with
                                scipy.signal.windows.gaussian, the same function)
Also crops the reference's default test image (test_jpeg.py:15,107-110) to the 256x256 centre patch.
"""
import ast
import os
import sys
import types

import numpy as np

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


def load_functions(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {'np': np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, 'exec'), ns)
    return ns


def load_kernels():
    import scipy.signal.windows as win
    shim = types.ModuleType('signal_shim')
    shim.gaussian = win.gaussian
    src = open(os.path.join(REF, 'helpers/kernels.py')).read().replace('from scipy import signal', '')
    ns = {'signal': shim}
    exec(compile(src, 'helpers/kernels.py', 'exec'), ns)
    return ns


def main():
    jh = load_functions(os.path.join(REF, 'compression/jpeg_helpers.py'), {'zigzag', 'jpeg_qtable', 'jpeg_qf_estimation'})
    k = load_kernels()
    out = {}
    out['qtable_luma'] = np.stack([jh['jpeg_qtable'](q, 0) for q in range(1, 101)])
    out['qtable_chroma'] = np.stack([jh['jpeg_qtable'](q, 1) for q in range(1, 101)])
    out['zigzag8'] = jh['zigzag'](8)
    out['zigzag4'] = jh['zigzag'](4)
    out['qf_est_luma'] = np.array([jh['jpeg_qf_estimation'](jh['jpeg_qtable'](q, 0), 0) for q in range(1, 101)])
    out['qf_est_chroma'] = np.array([jh['jpeg_qf_estimation'](jh['jpeg_qtable'](q, 1), 1) for q in range(1, 101)])
    for cfa in ('gbrg', 'rggb', 'bggr'):
        out['upk_' + cfa] = k['upsampling_kernel'](cfa)
    for i, a in enumerate(k['gamma_kernels']()):
        out['gamma_%d' % i] = a
    out['bilin3'] = k['bilin_kernel'](3)
    out['bilin5'] = k['bilin_kernel'](5)
    for kl, std in ((5, 0.83), (5, 2.0), (3, 0.5), (7, 1.5)):
        out['gkern_%d_%s' % (kl, str(std).replace('.', 'p'))] = k['gkern'](kl, std)
    f = np.array([[0, 0, 0, 0, 0], [0, -1, -2, -1, 0], [0, -2, 12, -2, 0], [0, -1, -2, -1, 0], [0, 0, 0, 0, 0]])
    out['constrained_init'] = k['repeat_2dfilter'](f, 3)
    out['center_mask_5_3'] = k['center_mask_2dfilter'](5, 3)
    np.savez_compressed(os.path.join(OUT, 'reference_numpy.npz'), **out)

    from PIL import Image
    im = Image.open(os.path.join(REF, 'docs/schematic_overview.png')).convert('RGB')
    w, h = im.size
    x0, y0 = (w - 256) // 2, (h - 256) // 2
    im.crop((x0, y0, x0 + 256, y0 + 256)).save(os.path.join(OUT, 'schematic_crop_256.png'))
    print('wrote', sorted(out.keys()))


if __name__ == '__main__':
    sys.exit(main())
