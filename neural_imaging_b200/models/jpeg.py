"""JPEG compression models on the B200 path.

`DifferentiableJPEG` : the differentiable codec (reference models/jpeg.py:45-159) as one fused CUDA kernel forward
                       and one backward (csrc/djpeg.cu) instead of ~40 TensorFlow ops.
`JPEG`               : the toolbox-facing wrapper (reference models/jpeg.py:162-285): quality randomisation, codec
                       selection, `process(batch_x, quality=None, return_entropy=False)`.
`differentiable_jpeg`: lazily-created shared instance (reference models/jpeg.py:38-42).
"""
import numpy as np

from .. import ops
from ..compression import jpeg_helpers
from ..compression.jpeg_helpers import jpeg_qf_estimation, jpeg_qtable
from ..helpers.utils import is_number
from ..tensor import as_device, empty, wrap
from .tfmodel import TFModel

_common_codec = None


def is_valid_quality(quality):
    if is_number(quality) and 1 <= quality <= 100:
        return True
    if hasattr(quality, '__getitem__') and len(quality) > 1 and all((1 <= x <= 100) for x in quality):
        return True
    return False


def differentiable_jpeg(x, quality):
    global _common_codec
    if _common_codec is None:
        _common_codec = JPEG(None, 'soft')
    return _common_codec.process(x, quality)


class DifferentiableJPEG(object):
    """Callable model: ``y, X = model(x)`` with x (N,H,W,3) in [0,1]; X = de-quantised DCT coefficients
    (3*N*H/8*W/8, 8, 8) in the reference block order."""

    def __init__(self, quality=None, rounding_approximation='sin', rounding_approximation_steps=5, trainable=False):
        if quality is not None and not is_valid_quality(quality):
            raise ValueError('Invalid JPEG quality: requires int in [1,100] or an iterable with least 2 such numbers')
        if rounding_approximation is not None and rounding_approximation not in ['sin', 'harmonic', 'soft']:
            raise ValueError('Unsupported rounding approximation: {}'.format(rounding_approximation))
        if rounding_approximation is None:
            # the reference fails inside Quantization(None, ...) (models/layers.py:99-100)
            raise ValueError('Unsupported quantization: {}'.format(rounding_approximation))
        if is_number(quality):
            ql, qc = jpeg_qtable(quality, 0), jpeg_qtable(quality, 1)
        else:
            ql, qc = np.ones((8, 8), dtype=np.float32), np.ones((8, 8), dtype=np.float32)
        self._store = None
        if trainable:
            # the tables are model weights (self.add_weight('Q_mtx_luma', [8, 8]) ..., models/jpeg.py:58-62): kept in a flat parameter
            # store like every other model; the kernels take the tables by value, so each call reads the current 128 floats back
            from .. import nn
            self._store = nn.ParamStore()
            self._pl = self._store.add('Q_mtx_luma', (8, 8), np.asarray(ql, np.float32), True)
            self._pc = self._store.add('Q_mtx_chroma', (8, 8), np.asarray(qc, np.float32), True)
            self._store.finalize()
            self._dq = None
        self._ql_host, self._qc_host = np.asarray(ql, np.float32), np.asarray(qc, np.float32)
        self.quality = quality
        self.trainable = trainable
        self.rounding_approximation = rounding_approximation
        self.rounding_approximation_steps = rounding_approximation_steps

    # the 8x8 tables: host arrays for a fixed codec, the current value of the weights for a trainable one
    @property
    def _q_mtx_luma(self):
        return self._pl.value.detach().cpu().numpy() if self._store is not None else self._ql_host

    @_q_mtx_luma.setter
    def _q_mtx_luma(self, v):
        if self._store is not None:
            self._pl.value.copy_(as_device(np.asarray(v, np.float32)))
        else:
            self._ql_host = np.asarray(v, np.float32)

    @property
    def _q_mtx_chroma(self):
        return self._pc.value.detach().cpu().numpy() if self._store is not None else self._qc_host

    @_q_mtx_chroma.setter
    def _q_mtx_chroma(self, v):
        if self._store is not None:
            self._pc.value.copy_(as_device(np.asarray(v, np.float32)))
        else:
            self._qc_host = np.asarray(v, np.float32)

    @property
    def trainable_weights(self):
        return [wrap(self._pl.value), wrap(self._pc.value)] if self._store is not None else []

    def __call__(self, inputs, want_coeffs=True):
        x = as_device(inputs)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        out = ops.djpeg_fwd(x, self._q_mtx_luma, self._q_mtx_chroma, self.rounding_approximation, want_coeffs=want_coeffs)
        if want_coeffs:
            return wrap(out[0]), wrap(out[1])
        return wrap(out)

    call = __call__

    def forward_into(self, x, y):
        """x, y: device tensors (no conversion); used by the workflow's training step."""
        return ops.djpeg_fwd(x, self._q_mtx_luma, self._q_mtx_chroma, self.rounding_approximation, out=y)

    def backward(self, x, dy, dx=None):
        """dx = J^T dy; for a trainable codec the table gradients land in the parameter store (self._pl.grad, self._pc.grad)."""
        ql, qc = self._q_mtx_luma, self._q_mtx_chroma
        if self._store is not None:
            from ..tensor import ptr, stream
            if self._dq is None:
                self._dq = empty((128,))
            n, h, w = ops._nhw3(x)
            _, pl = ops._f32(ql)
            _, pc = ops._f32(qc)
            from .. import _lib
            _lib.lib().ni_djpeg_bwd_tables(ptr(x), ptr(dy), ptr(self._dq), n, h, w, pl, pc, ops.QUANT_MODES[self.rounding_approximation], stream())
            self._pl.grad.copy_(self._dq[:64].view(8, 8))
            self._pc.grad.copy_(self._dq[64:].view(8, 8))
        return ops.djpeg_bwd(x, dy, ql, qc, self.rounding_approximation, out=dx)


class JPEG(TFModel):

    def __init__(self, quality=None, codec='soft', trainable=False):
        super().__init__()
        if codec is not None and codec not in ['libjpeg', 'soft', 'sin', 'harmonic']:
            raise ValueError('Unsupported codec version: {}'.format(codec))
        self._model = None if codec == 'libjpeg' else DifferentiableJPEG(quality, codec, trainable=trainable)
        self.codec = codec
        self.quality = quality

    @staticmethod
    def loss(target, compressed, sample_weight=None):
        """tf.keras.losses.MeanSquaredError()(a, b, sample_weight). The workflow passes the (NaN) entropy as the third
        positional argument, which Keras interprets as sample_weight => NaN; reproduced."""
        if sample_weight is not None and is_number(sample_weight) and np.isnan(sample_weight):
            return float('nan')
        a, b = as_device(target), as_device(compressed)
        v = ops.image_loss(a, b, 'L2') / (255.0 * 255.0)
        return wrap(v.reshape(())) if sample_weight is None else wrap(v.reshape(()) * float(sample_weight))

    def reset_performance_stats(self):
        self.performance = self._reset_performance(['entropy', 'ssim', 'psnr'])

    @property
    def parameters(self):
        return [] if self._model is None else self._model.trainable_weights

    def count_parameters(self):
        return int(sum(int(np.prod(p.shape)) for p in self.parameters))

    def _draw_quality(self, quality):
        quality = self.quality if quality is None else quality
        if not is_valid_quality(quality):
            raise ValueError('Invalid or unspecified JPEG quality!')
        if hasattr(quality, '__getitem__') and len(quality) > 2:
            return int(np.random.choice(quality))
        if hasattr(quality, '__getitem__') and len(quality) == 2:
            return int(np.random.randint(quality[0], quality[1]))
        if is_number(quality) and 1 <= quality <= 100:
            return int(quality)
        raise ValueError('Invalid quality! {}'.format(quality))

    def process(self, batch_x, quality=None, return_entropy=False):
        quality = self._draw_quality(quality)
        if self._model is None:
            # codec='libjpeg': the reference's host-side file codec for the final validation (models/jpeg.py:227-233); numpy in, numpy out
            batch_x = batch_x if isinstance(batch_x, np.ndarray) else np.asarray(batch_x.numpy() if hasattr(batch_x, 'numpy') else batch_x)
            y = jpeg_helpers.compress_batch(batch_x, quality)[0]
            return (y, np.nan) if return_entropy else y
        y = self._with_quality(quality, lambda: self._model(batch_x, want_coeffs=False))
        return (y, np.nan) if return_entropy else y

    def _with_quality(self, quality, fn):
        """Temporarily swap the tables like the reference (models/jpeg.py:235-243; not re-entrant)."""
        m = self._model
        if quality != self.quality:
            old = m._q_mtx_luma, m._q_mtx_chroma
            m._q_mtx_luma, m._q_mtx_chroma = jpeg_qtable(quality, 0), jpeg_qtable(quality, 1)
            try:
                return fn()
            finally:
                m._q_mtx_luma, m._q_mtx_chroma = old
        return fn()

    def __repr__(self):
        if self._model is not None:
            return 'JPEG(quality={},codec="{}",trainable={})'.format(self.quality, self.codec, self._model.trainable)
        return 'JPEG(quality={},codec="{}")'.format(self.quality, self.codec)

    def summary(self, quality=None):
        return 'JPEG ({}) {}'.format(self.codec, self._quality_mode(quality))

    def summary_compact(self, quality=None):
        return 'JPEG ({}) {}'.format(self.codec, self._quality_mode(quality))

    def estimate_qf(self, channel=0):
        return jpeg_qf_estimation(self._model._q_mtx_luma, channel)

    def _quality_mode(self, quality=None):
        quality = quality or self.quality
        if self._model is not None and self._model.trainable:
            return 'trainable QF~{}/{}'.format(jpeg_qf_estimation(self._model._q_mtx_luma, 0), jpeg_qf_estimation(self._model._q_mtx_chroma, 1))
        if is_number(quality):
            return 'QF={}'.format(quality)
        if hasattr(quality, '__getitem__') and len(quality) == 2:
            return 'QF~[{},{}]'.format(*quality)
        if hasattr(quality, '__getitem__') and len(quality) > 2:
            return 'QF~{{{}}}'.format(','.join(str(x) for x in quality))
        return 'QF=?'
