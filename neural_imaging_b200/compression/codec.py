"""The l3ic byte-stream codec of the learned image codec — device mirror of the reference's compression/codec.py:14-265.

Same functions and stream format (`compress(batch_x, model)` -> bytes, `decompress(stream, model)` -> image, `simulate_compression`,
`L3ICError`), but the whole path stays on the GPU: DCN encoder -> code-book indices -> per-layer FSE / run / raw coding -> length table ->
byte stream (csrc/l3ic.cu, one warp per latent layer), and back. `compress_batch` / `decompress_batch` are the forms the device is for:
the reference codes one image per call (codec.py:103-104), here a batch of n images is n * n_latent independent streams in one launch.
Streams are byte-identical to the reference's (tests/test_l3ic.py against oracle/_ref and the committed golden vectors).
"""
import io

import numpy as np
import torch

from .. import _lib
from ..tensor import as_device, device, ptr, stream, wrap

_STATUS = {1: 'FSE coder error', 2: 'latent shape in the stream does not match the model', 3: 'truncated stream',
           4: 'layer does not decode to latent_x * latent_y code-book indices', 5: 'layer-length table cannot be coded / decoded',
           6: 'layer data compresses to a single byte? Something is wrong!'}


class L3ICError(Exception):
    pass


def _codebook(model):
    cb = np.asarray(model.get_codebook(), dtype=np.float32).reshape(-1)
    if len(cb) > 256:
        raise L3ICError('Code-books with more than 256 centers are not supported')
    return cb


def _raise_on(status, what):
    bad = np.flatnonzero(status)
    if len(bad):
        i = int(bad[0])
        raise L3ICError('[l3ic {}] image {}: {}'.format(what, i, _STATUS.get(int(status[i]), 'error {}'.format(int(status[i])))))


def encode_latent(batch_z, code_book):
    """Quantised latents (n,h,w,c) -> list of n byte streams (codec.py:106-186 after `model.compress`)."""
    z = as_device(batch_z)
    if z.dim() == 3:
        z = z.unsqueeze(0)
    n, h, w, c = (int(v) for v in z.shape)
    cb = torch.from_numpy(np.ascontiguousarray(code_book, dtype=np.float32)).to(device())
    hw = h * w
    slot = -(-hw // 16) * 16
    stride = -(-(5 + 2 * c + c * hw) // 16) * 16
    dev = device()
    indices = torch.empty((n * c * hw,), dtype=torch.uint8, device=dev)
    layer_bytes = torch.empty((n * c, slot), dtype=torch.uint8, device=dev)
    layer_len = torch.empty((n * c,), dtype=torch.int32, device=dev)
    streams = torch.empty((n, stride), dtype=torch.uint8, device=dev)
    stream_len = torch.empty((n,), dtype=torch.int32, device=dev)
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    _lib.lib().ni_l3ic_encode(ptr(z), n, h, w, c, ptr(cb), int(cb.numel()), ptr(indices), ptr(layer_bytes), slot, ptr(layer_len), ptr(streams),
                              stride, ptr(stream_len), ptr(status), stream())
    _raise_on(status.cpu().numpy(), 'encoder')
    sizes, data = stream_len.cpu().numpy(), streams.cpu().numpy()
    return [data[i, :sizes[i]].tobytes() for i in range(n)]


def decode_latent(streams, latent_shape, code_book):
    """List of n byte streams -> quantised latents (n,h,w,c) on the device (codec.py:201-255 before `model.decompress`)."""
    streams = [s.getvalue() if isinstance(s, io.BytesIO) else bytes(s) for s in streams]
    h, w, c = (int(v) for v in latent_shape)
    n = len(streams)
    stride = -(-max(max(len(s) for s in streams), 16) // 16) * 16
    host = np.zeros((n, stride), dtype=np.uint8)
    for i, s in enumerate(streams):
        host[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    dev = device()
    cb = torch.from_numpy(np.ascontiguousarray(code_book, dtype=np.float32)).to(dev)
    d_streams = torch.from_numpy(host).to(dev)
    d_len = torch.from_numpy(np.array([len(s) for s in streams], dtype=np.int32)).to(dev)
    slot = -(-(h * w + 4) // 16) * 16
    layer_off = torch.empty((n * c,), dtype=torch.int32, device=dev)
    layer_len = torch.empty((n * c,), dtype=torch.int32, device=dev)
    indices = torch.zeros((n * c, slot), dtype=torch.uint8, device=dev)
    latent = torch.empty((n, h, w, c), dtype=torch.float32, device=dev)
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    _lib.lib().ni_l3ic_decode(ptr(d_streams), stride, ptr(d_len), n, h, w, c, ptr(cb), int(cb.numel()), ptr(layer_off), ptr(layer_len), ptr(indices),
                              slot, ptr(latent), ptr(status), stream())
    _raise_on(status.cpu().numpy(), 'decoder')
    return latent


def compress_batch(batch_x, model):
    """Images (n,H,W,3) -> list of n l3ic byte streams (the batched form of `compress`)."""
    x = as_device(batch_x)
    if x.dim() == 3:
        x = x.unsqueeze(0)
    return encode_latent(model.compress(x), _codebook(model))


def decompress_batch(streams, model):
    """List of l3ic byte streams (all of the model's latent shape) -> decoded images (n,H,W,3), device tensor with .numpy()."""
    streams = list(streams)
    first = streams[0].getvalue() if isinstance(streams[0], io.BytesIO) else bytes(streams[0])
    if len(first) < 3:
        raise L3ICError('[l3ic decoder] truncated stream')
    shape = tuple(int(v) for v in first[:3])
    if model.latent_shape[-1] != shape[2]:
        raise L3ICError('the specified model ({}c) does not match the coded stream ({}c)'.format(model.latent_shape[-1], shape[2]))
    return model.decompress(wrap(decode_latent(streams, shape, _codebook(model))))


def compress(batch_x, model, verbose=False):
    """Serialize ONE image as a bytes sequence; the feature maps are encoded as separate layers (codec.py:87-186)."""
    batch_x = batch_x if isinstance(batch_x, torch.Tensor) else np.asarray(batch_x)
    if batch_x.ndim == 3:
        batch_x = batch_x[None]
    assert batch_x.ndim == 4
    assert batch_x.shape[0] == 1
    out = compress_batch(batch_x, model)[0]
    if verbose:
        print('[l3ic encoder]', 'Code book:', model.get_codebook(), '->', len(out), 'bytes')
    return out


def decompress(stream, model=None, verbose=False):
    """Decompress an image from the given bytes sequence (codec.py:189-265). Returns a numpy array (1,H,W,3)."""
    if type(stream) is bytes:
        pass
    elif type(stream) is io.BytesIO:
        stream = stream.getvalue()
    elif hasattr(stream, 'read'):
        stream = stream.read()
    else:
        raise ValueError('Unsupported stream type!')
    if model is None:
        model = restore('{}c'.format(stream[2]))        # codec.py:228-229: a preset name or a directory (raises ValueError when absent)
    if verbose:
        print('[l3ic decoder]', 'Latent space', tuple(stream[:3]))
    return decompress_batch([stream], model).numpy()


def simulate_compression(batch_x, dcn):
    """The entire compression and decompression through the byte stream: (decompressed image, byte count) (codec.py:18-26)."""
    compressed_image = compress(batch_x, dcn)
    batch_y = decompress(compressed_image, dcn)
    return batch_y, len(compressed_image)


def compress_n_stats(batch_x, dcn):
    """Code every image of the batch through the byte stream and report ssim / psnr / entropy / bytes / bpp per image
    (codec.py:29-55; scalars instead of arrays for a batch of one)."""
    from ..helpers import metrics
    from ..training.compression import latent_entropy
    x = batch_x.numpy() if hasattr(batch_x, 'numpy') and not isinstance(batch_x, np.ndarray) else np.asarray(batch_x)
    if x.ndim == 3:
        x = x[None]
    streams = compress_batch(x, dcn)
    batch_y = decompress_batch(streams, dcn).numpy()
    z = dcn.compress(x).numpy()
    code_book = np.asarray(dcn.get_codebook(), dtype=np.float64).reshape(-1)
    stats = {k: np.zeros((x.shape[0],)) for k in ('ssim', 'psnr', 'entropy', 'bytes', 'bpp')}
    for i in range(x.shape[0]):
        stats['entropy'][i] = latent_entropy(z[i], code_book)          # helpers/stats.py:119-131 on one image's latent
        stats['bytes'][i] = len(streams[i])
        stats['ssim'][i] = metrics.ssim(x[i], batch_y[i])
        stats['psnr'][i] = metrics.psnr(x[i], batch_y[i])
        stats['bpp'][i] = 8 * len(streams[i]) / x[i].shape[0] / x[i].shape[1]
    if x.shape[0] == 1:
        stats = {k: v[0] for k, v in stats.items()}
    return batch_y, stats


def restore(dir_name, patch_size=None, fetch_stats=False):
    """Restore a DCN by directory or preset name — wrapper over tfmodel.restore (codec.py:275-291)."""
    from ..models import compression, tfmodel
    return tfmodel.restore(dir_name, compression, key='codec', patch_size=patch_size, fetch_stats=fetch_stats)
