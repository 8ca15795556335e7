"""Constant filter banks (NumPy, host side). Values follow reference helpers/kernels.py:9-123."""
import numpy as np

_CFA_SLOTS = {  # CFA pattern -> for each of the 4 stacked raw channels, the index of its slot in the 12-feature layout
    'GBRG': (6, 1, 10, 5),
    'RGGB': (0, 4, 7, 11),
    'BGGR': (9, 4, 7, 2),
}


def upsampling_kernel(cfa_pattern='gbrg'):
    """(4, 12) one-hot matrix routing the RGGB-style stack into the depth_to_space(2) RGB layout (kernels.py:9-43)."""
    key = cfa_pattern.upper()
    if key not in _CFA_SLOTS:
        raise ValueError('Unsupported CFA pattern: {}'.format(cfa_pattern))
    upk = np.zeros((4, 12), dtype=np.int64)
    for row, col in enumerate(_CFA_SLOTS[key]):
        upk[row, col] = 1
    return upk


def gamma_kernels():
    """Weights of the toy gamma-correction MLP 3 -> 12 (tanh) -> 3 (kernels.py:46-67)."""
    k1 = np.array([2.9542332, 17.780445, 0.6280197, 0.40384966])
    b1 = np.array([0.4047071, 1.1489044, -0.17624384, 0.47826886])
    k2 = np.array([0.44949612, 0.78081024, 0.97692937, -0.24265033])
    b2 = -0.4702738
    d1k, d1b, d2k, d2b = np.zeros((3, 12)), np.zeros((12,)), np.zeros((12, 3)), np.zeros((3,))
    for c in range(3):
        sl = slice(4 * c, 4 * c + 4)
        d1k[c, sl], d1b[sl], d2k[sl, c], d2b[c] = k1, b1, k2, b2
    return d1k, d1b, d2k, d2b


def bilin_kernel(kernel=3):
    """Bilinear demosaicing filter (k, k, 3, 3), zero-padded 3x3 core (kernels.py:70-91)."""
    g = np.array([[0, .25, 0], [.25, 1, .25], [0, .25, 0]])
    rb = np.array([[.25, .5, .25], [.5, 1, .5], [.25, .5, .25]])
    dmf = np.zeros((3, 3, 3, 3), np.float32)
    dmf[:, :, 0, 0], dmf[:, :, 1, 1], dmf[:, :, 2, 2] = rb, g, rb
    if kernel > 3:
        p = (kernel - 3) // 2
        dmf = np.pad(dmf, ((p, p), (p, p), (0, 0), (0, 0)), 'constant')
    return dmf


def gkern(kernlen=5, std=0.83):
    """Normalised 2-D Gaussian = outer product of scipy.signal.gaussian(kernlen, std) (kernels.py:94-98)."""
    n = np.arange(kernlen, dtype=np.float64) - (kernlen - 1) / 2.0
    g1 = np.exp(-0.5 * (n / std) ** 2)
    g2 = np.outer(g1, g1)
    return g2 / g2.sum()


def repeat_2dfilter(f, channels=3, pad=0):
    """(k, k, C, C) kernel with `f` on the channel diagonal (kernels.py:101-114)."""
    f = np.pad(np.asarray(f, dtype=np.float64), pad, 'constant') if pad else np.asarray(f, dtype=np.float64)
    rf = np.zeros(f.shape + (channels, channels))
    for c in range(channels):
        rf[:, :, c, c] = f
    return rf


def center_mask_2dfilter(f_size, channels):
    ind = np.zeros((f_size, f_size, channels, channels))
    for c in range(channels):
        ind[f_size // 2, f_size // 2, c, c] = 1
    return ind
