"""Turns the fixtures of the reference's OWN entropy-coder test (pyfse/test_fse.py:11-25: tests/string.txt, all_zeros.dat, binary.dat,
numbers.dat) into tests/golden/pyfse_reference_tests.npz: the four inputs exactly as that test builds them, plus the streams the reference
library (oracle/_ref/libfse_ref.so) produces for them. Run here (the reference is not on the GPU box):

    make -C oracle && python tests/golden/make_pyfse_fixture_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_l3ic as R  # noqa: E402

SRC = '/root/reference/pyfse/tests'


def inputs():
    out = {}
    with open(os.path.join(SRC, 'string.txt')) as f:
        out['ascii'] = f.read().encode('ascii')
    with open(os.path.join(SRC, 'all_zeros.dat')) as f:
        out['all_zeros'] = bytes([int(x) for x in f.read().strip()])
    with open(os.path.join(SRC, 'binary.dat')) as f:
        out['binary'] = bytes([int(x) for x in f.read()])
    with open(os.path.join(SRC, 'numbers.dat')) as f:
        out['numbers'] = bytes([127 + int(x) for x in f.read().split(', ')])
    return out


def main():
    lib = R.reference_library()
    assert lib is not None, 'build oracle/_ref first: make -C oracle'
    data = {}
    for key, raw in inputs().items():
        r = R.ref_compress(lib, raw)
        data['in_' + key] = np.frombuffer(raw, np.uint8)
        data['out_' + key] = np.frombuffer(r, np.uint8) if isinstance(r, bytes) else np.array([r], dtype=np.int64)
        if isinstance(r, bytes):
            assert R.ref_decompress(lib, r, 10 * len(r)) == raw
        print(key, len(raw), 'symbols ->', len(r) if isinstance(r, bytes) else r)
    path = os.path.join(ROOT, 'tests', 'golden', 'pyfse_reference_tests.npz')
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
