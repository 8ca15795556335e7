"""JPEG helpers (host, NumPy). Quantisation tables follow reference compression/jpeg_helpers.py:253-310; `compress_batch`
(:82-114) is the reference's HOST-side libjpeg codec (used for the "final validation" with JPEG(codec='libjpeg')): it is a
per-image file codec by definition (the reference goes through imageio -> Pillow -> libjpeg), not a fallback of a device operator.
The marker-parsing tooling of that file is out of scope."""
import io

import numpy as np

# IJG base tables (ITU-T T.81 Annex K), row-major
_LUMA = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                  14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                  49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], np.float32).reshape(8, 8)
_CHROMA = np.full((8, 8), 99, np.float32)
_CHROMA[:4, :4] = [[17, 18, 24, 47], [18, 21, 26, 66], [24, 26, 56, 99], [47, 66, 99, 99]]


def zigzag(n):
    """Zig-zag scan order as an (n, n) uint16 index matrix (jpeg_helpers.py:253-261)."""
    zz = np.zeros((n, n), dtype=np.uint16)
    idx = 0
    for s in range(2 * n - 1):
        cells = [(x, s - x) for x in range(n) if 0 <= s - x < n]
        if s % 2 == 0:
            cells.reverse()          # even anti-diagonals run bottom-left -> top-right
        for x, y in cells:
            zz[x, y] = idx
            idx += 1
    return zz


def jpeg_qtable(quality, channel=0):
    """IJG-scaled quantisation table for `quality` in [1, 100]; channel 0 = luma, otherwise chroma (jpeg_helpers.py:264-305)."""
    q = min(100, max(1, quality))
    scale = 5000 / q if q < 50 else 200 - q * 2
    t = np.floor(((_LUMA if channel == 0 else _CHROMA) * scale + 50) / 100)
    return np.clip(t, 1, 255).astype(np.float32)


def jpeg_qf_estimation(q_mtx, channel=0):
    """Quality factor whose IJG table is closest (mean absolute difference) to q_mtx (jpeg_helpers.py:308-310)."""
    q_mtx = np.asarray(q_mtx)
    errors = [np.mean(np.abs(jpeg_qtable(qf, channel) - q_mtx)) for qf in range(1, 101)]
    return int(np.argmin(errors)) + 1


_SUBSAMPLING = {'4:4:4': 0, '4:2:2': 1, '4:2:0': 2}


def _libjpeg_roundtrip(img_u8, quality, subsampling):
    from PIL import Image                       # Pillow's libjpeg: what imageio.imsave(format='jpg') calls in the reference
    buf = io.BytesIO()
    Image.fromarray(img_u8).save(buf, format='JPEG', quality=int(quality), subsampling=_SUBSAMPLING[subsampling])
    data = buf.getvalue()
    return np.asarray(Image.open(io.BytesIO(data)).convert('RGB' if img_u8.ndim == 3 else 'L')), len(data)


def compress_batch(batch_x, jpeg_quality, effective=False, subsampling='4:4:4'):
    """Standard JPEG round trip of an (n,h,w,3) / (h,w,3) array in [0,1] (or [0,255]); returns (images in [0,1], sizes in bytes)
    (reference compression/jpeg_helpers.py:82-114). `effective=True` (payload without headers) needs the marker parser, which is out
    of scope."""
    if effective:
        raise NotImplementedError('effective payload size needs the JPEG marker parser (out of scope)')
    batch_x = np.asarray(batch_x)
    if batch_x.max() > 1:
        batch_x = batch_x.astype(np.float32) / (2 ** 8 - 1)
    if batch_x.ndim == 3:
        img, nbytes = _libjpeg_roundtrip((255 * batch_x).astype(np.uint8).squeeze(), jpeg_quality, subsampling)
        return img / (2 ** 8 - 1), nbytes
    if batch_x.ndim == 4:
        out, sizes = np.zeros_like(batch_x), []
        for r in range(batch_x.shape[0]):
            img, nbytes = _libjpeg_roundtrip((255 * batch_x[r]).astype(np.uint8).squeeze(), jpeg_quality, subsampling)
            out[r] = img.astype(np.float32).reshape(out[r].shape) / (2 ** 8 - 1)
            sizes.append(nbytes)
        return out, sizes
    raise ValueError('compress_batch expects an (n,h,w,c) or (h,w,c) array')
