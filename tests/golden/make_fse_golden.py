"""Generates tests/golden/fse_vectors.npz by RUNNING THE REFERENCE'S OWN entropy coder (oracle/_ref/libfse_ref.so, compiled by
oracle/Makefile from /root/reference/pyfse/FiniteStateEntropy/lib) on seeded inputs; the container-level vectors put the restated
compression/codec.py format (oracle/ref_l3ic.py) on top of that library. Run here (the reference is not on the GPU box):

    make -C oracle && python tests/golden/make_fse_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_l3ic as R  # noqa: E402


def inputs(seed=1234):
    rs = np.random.RandomState(seed)
    out = []
    for n in (2, 3, 4, 5, 8, 17, 64, 255, 256, 257, 1024, 4096):
        out.append(np.clip(np.round(rs.normal(15, 1.5, n)), 0, 31).astype(np.uint8))      # what a DCN latent layer looks like
        out.append(rs.randint(0, 256, n).astype(np.uint8))                                # not compressible
    for n in (16, 256, 1000):
        out.append(np.full(n, 9, np.uint8))                                               # one repeated symbol
        a = np.full(n, 200, np.uint8)
        a[rs.randint(0, n, max(1, n // 40))] = rs.randint(0, 256)                         # nearly constant: header run-length paths
        out.append(a)
        out.append((rs.geometric(0.3, n) - 1).clip(0, 255).astype(np.uint8))
        out.append(np.clip(np.round(rs.laplace(128, 20, n)), 0, 255).astype(np.uint8))    # wide alphabet, fallback normalisation territory
        out.append((rs.randint(0, 2, n) * 255).astype(np.uint8))
    out.append(np.array([len(x) % 7 + 40 for x in out] * 2, dtype=np.uint16).view(np.uint8))   # a layer-length table
    return out


def latents(seed=4321):
    rs = np.random.RandomState(seed)
    z = np.clip(np.round(rs.normal(0, 1.2, (3, 16, 16, 32))), -15, 16)
    z[0, :, :, 3] = 2.0                                # constant layer -> run
    z[1, :, :, 5] = rs.randint(-15, 17, (16, 16))      # flat layer -> raw
    z[2, :, :, 7] *= 0.01                              # nearly constant
    return z.astype(np.float32)


def main():
    lib = R.reference_library()
    assert lib is not None, 'build oracle/_ref first: make -C oracle'
    data = {}
    for i, a in enumerate(inputs()):
        r = R.ref_compress(lib, a.tobytes())
        data['in_%02d' % i] = a
        data['out_%02d' % i] = np.frombuffer(r, np.uint8) if isinstance(r, bytes) else np.array([r], dtype=np.int64)
    z = latents()
    code_book = np.arange(-15, 17, dtype=np.float32)
    data['latent'] = z
    data['code_book'] = code_book
    for i in range(z.shape[0]):
        s = R.l3ic_compress(z[i:i + 1], code_book, compress=lambda b: R.ref_compress(lib, b))
        back = R.l3ic_decompress(s, code_book, decompress=lambda b, cap: R.ref_decompress(lib, b, cap))
        assert np.array_equal(back, code_book[R.vq(z[i:i + 1], code_book)].reshape(back.shape))     # values off the code book snap to it
        data['stream_%d' % i] = np.frombuffer(s, np.uint8)
    path = os.path.join(ROOT, 'tests', 'golden', 'fse_vectors.npz')
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), 'bytes', len(inputs()), 'strings')


if __name__ == '__main__':
    main()
