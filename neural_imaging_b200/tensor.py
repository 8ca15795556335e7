"""Device-tensor plumbing. PyTorch is used for storage, streams and torch.distributed only — every computation on
the hot path is a kernel of libni_b200.so.

Reference callers invoke ``.numpy()`` on model outputs (training/pipeline.py:40, training/validation.py:35,125,
workflows/manipulation_classification.py:180,187), so results are returned as :class:`NITensor`, a
``torch.Tensor`` subclass whose ``numpy()`` copies to the host.
"""
import numpy as np
import os

import torch


class NITensor(torch.Tensor):
    def numpy(self):
        return torch.Tensor.numpy(self.detach().as_subclass(torch.Tensor).cpu())

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)


def wrap(t):
    return t.as_subclass(NITensor)


def device():
    if not torch.cuda.is_available():
        raise RuntimeError('neural_imaging_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def as_device(x, dtype=torch.float32):
    """numpy array / torch tensor -> contiguous device tensor of `dtype` (host->device copy if needed)."""
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
    elif isinstance(x, torch.Tensor):
        t = x.detach().as_subclass(torch.Tensor)
    else:
        t = torch.as_tensor(np.asarray(x))
    if t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        t = t.to(device(), non_blocking=True)
    return t.contiguous()


def empty(shape, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device=device())


def zeros(shape, dtype=torch.float32):
    return torch.zeros(shape, dtype=dtype, device=device())


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    if t is None:
        return None
    return t.data_ptr()


class Workspace:
    """Named persistent device buffers (re-used across steps: no allocator traffic inside the step, CUDA-graph safe)."""

    def __init__(self):
        self._bufs = {}

    def get(self, name, shape, dtype=torch.float32):
        """Buffers are keyed by (name, shape, dtype): a buffer handed out once is never freed or re-pointed while the workspace
        lives, so a captured CUDA graph that baked its address in stays valid when another batch shape comes through later."""
        key = (name, tuple(int(s) for s in shape), dtype)
        b = self._bufs.get(key)
        if b is None:
            b = self._bufs[key] = torch.empty(key[1], dtype=dtype, device=device())
        return b

    def clear(self):
        self._bufs.clear()


class _Nvtx:
    """NVTX range per pipeline stage (nsys / ncu --nvtx): `with nvtx('fan.backward'): ...`. Host-side markers only (no device work, safe
    inside CUDA-graph capture); switched on with NI_NVTX=1, otherwise a no-op context."""
    enabled = os.environ.get('NI_NVTX', '0') == '1'

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Nvtx.enabled:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if _Nvtx.enabled:
            torch.cuda.nvtx.range_pop()
        return False


def nvtx(name):
    return _Nvtx(name)
