"""`pyfse.pyfse`: compress / decompress (+ batch forms) on the device. See the package docstring."""
import numpy as np
import torch

from .. import _lib
from ..tensor import device, ptr, stream

_ERRORS = {-1: 'Unspecified error code', -2: 'Destination buffer is too small', -3: 'Unsupported max Log', -4: 'Specified maxSymbolValue is too small',
           -5: 'Src size is incorrect', -6: 'Corrupted block detected'}


class FSEException(Exception):
    pass


class FSENotCompressibleError(FSEException):
    pass


class FSESymbolRepetitionError(FSEException):
    pass


def _pack(strings):
    """list of bytes -> (device uint8 matrix with one string per row, device int32 lengths, row stride)."""
    lens = np.array([len(s) for s in strings], dtype=np.int32)
    stride = max(16, int(-(-max(int(lens.max()), 1) // 16) * 16))
    host = np.zeros((len(strings), stride), dtype=np.uint8)
    for i, s in enumerate(strings):
        host[i, :len(s)] = np.frombuffer(bytes(s), dtype=np.uint8)
    return torch.from_numpy(host).to(device()), torch.from_numpy(lens).to(device()), stride


def compress_batch(strings):
    """FSE-code every byte string of the list in one launch. Returns a list whose entries are `bytes`, or the exception INSTANCE the
    reference would have raised for that string (FSENotCompressibleError / FSESymbolRepetitionError / FSEException)."""
    strings = [bytes(s) for s in strings]
    if not strings:
        return []
    src, lens, stride = _pack(strings)
    dst = torch.empty((len(strings), stride), dtype=torch.uint8, device=device())
    out_len = torch.empty((len(strings),), dtype=torch.int32, device=device())
    _lib.lib().ni_fse_compress_batch(ptr(src), stride, ptr(lens), ptr(dst), stride, ptr(out_len), len(strings), stream())
    sizes, data = out_len.cpu().numpy(), dst.cpu().numpy()
    res = []
    for i, r in enumerate(sizes):
        if r < 0:
            res.append(FSEException('Encoding Error: {}'.format(_ERRORS.get(int(r), 'Unspecified error code'))))
        elif r == 0:
            res.append(FSENotCompressibleError('Encoding Error: data is not compressible'))
        elif r == 1:
            res.append(FSESymbolRepetitionError('Encoding Error: input data is a repetition of a single byte - use RLE encoding instead'))
        else:
            res.append(data[i, :r].tobytes())
    return res


def decompress_batch(strings, max_length=0):
    """FSE-decode every string of the list in one launch (`max_length` as in decompress; 0 -> 10 x the longest input)."""
    strings = [bytes(s) for s in strings]
    if not strings:
        return []
    src, lens, stride = _pack(strings)
    cap = max(1, int(max_length) if max_length else 10 * max(len(s) for s in strings))
    per_string_cap = None if max_length else [10 * len(s) for s in strings]
    dst = torch.empty((len(strings), cap + 16), dtype=torch.uint8, device=device())
    out_len = torch.empty((len(strings),), dtype=torch.int32, device=device())
    if per_string_cap is not None and len(set(per_string_cap)) > 1:
        # the capacity is part of the reference's behaviour (it decides dstSize_tooSmall): run equal-capacity groups separately
        res = [None] * len(strings)
        for c in sorted(set(per_string_cap)):
            idx = [i for i, v in enumerate(per_string_cap) if v == c]
            for i, r in zip(idx, decompress_batch([strings[i] for i in idx], max_length=c)):
                res[i] = r
        return res
    # rows are padded by 16 bytes; the capacity handed to the decoder is `cap` (it decides dstSize_tooSmall, as in the reference)
    _lib.lib().ni_fse_decompress_batch(ptr(src), stride, ptr(lens), ptr(dst), cap + 16, cap, ptr(out_len), len(strings), stream())
    sizes, data = out_len.cpu().numpy(), dst.cpu().numpy()
    res = []
    for i, r in enumerate(sizes):
        if r < 0:
            res.append(FSEException('Decoding Error: {}'.format(_ERRORS.get(int(r), 'Unspecified error code'))))
        else:
            res.append(data[i, :r].tobytes())
    return res


def _one(result):
    if isinstance(result, Exception):
        raise result
    return result


def compress(src):
    """Compress bytes from 'src' and return FSE coded bytes (pyfse.pyx:24-51)."""
    return _one(compress_batch([src])[0])


def decompress(src, max_length=0):
    """Decompress FSE-coded bytes; the output buffer is `max_length` bytes (10 x the input if 0) (pyfse.pyx:53-72)."""
    return _one(decompress_batch([src], max_length)[0])
