"""Device mirror of the reference's `pyfse` package (pyfse/pyfse.pyx:17-72, a Cython wrapper of FSE_compress / FSE_decompress): same
functions, exceptions and error behaviour, bit-identical streams — computed by the CUDA kernels of csrc/l3ic.cu (one warp per stream).

`compress` / `decompress` keep the reference's one-string signatures (`from pyfse import pyfse; pyfse.compress(data)`); the batch
variants are what the device is for: thousands of short strings per launch. There is no host implementation behind these functions.
"""
from . import pyfse  # noqa: F401  (the reference is imported as `from pyfse import pyfse`)
