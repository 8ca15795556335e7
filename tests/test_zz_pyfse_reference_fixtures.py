"""The reference's OWN entropy-coder test (pyfse/test_fse.py:27-62) on its own fixtures (tests/golden/pyfse_reference_tests.npz, made by
tests/golden/make_pyfse_fixture_golden.py from pyfse/tests/{string.txt, all_zeros.dat, binary.dat, numbers.dat}): round trip, coded size
within [entropy, 1.1 x entropy], FSESymbolRepetitionError for the constant input — plus byte equality with the streams the reference
library produced. CPU: oracle restatement and the host build of the product's coder source; GPU: the kernels through the pyfse mirror.
(Named zz so that it runs after every other file: two of the inputs are 131,072 symbols, beyond what an l3ic layer can hold.)"""
import ctypes
import math
import os
import subprocess
from collections import Counter

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import ref_l3ic as R

KEYS = ('ascii', 'numbers', 'binary')


@pytest.fixture(scope='module')
def fixtures():
    with np.load(os.path.join(GOLDEN, 'pyfse_reference_tests.npz')) as d:
        return {k: d[k] for k in d.files}


def _entropy_bytes(raw):
    n = len(raw)
    return -sum(c / n * math.log2(c / n) for c in Counter(raw).values()) * n / 8


def _check_like_the_reference(raw, coded, decoded):
    limit = _entropy_bytes(raw)
    assert limit <= len(coded) <= 1.1 * limit           # test_fse.py:41-42
    assert decoded == raw                               # test_fse.py:44


def test_fixture_shapes(fixtures):
    assert len(fixtures['in_ascii']) == 2935 and len(fixtures['in_all_zeros']) == 2538
    assert len(fixtures['in_binary']) == 131072 and len(fixtures['in_numbers']) == 131072
    assert int(fixtures['out_all_zeros'][0]) == 1 and set(np.unique(fixtures['in_binary'])) == {0, 1}


@pytest.mark.parametrize('key', KEYS)
def test_oracle_on_reference_fixtures(fixtures, key):
    raw, want = fixtures['in_' + key].tobytes(), fixtures['out_' + key].tobytes()
    coded = R.fse_compress(raw)
    assert coded == want
    _check_like_the_reference(raw, coded, R.fse_decompress(coded, 10 * len(coded)))
    assert R.fse_compress(fixtures['in_all_zeros'].tobytes()) == 1


def test_device_source_on_reference_fixtures(fixtures, tmp_path):
    so = str(tmp_path / 'libfse_host.so')
    subprocess.check_call(['g++', '-O2', '-fPIC', '-shared', '-o', so, os.path.join(ROOT, 'tests', 'fse_host_harness.cpp')])
    lib = ctypes.CDLL(so)
    for key in KEYS:
        raw, want = fixtures['in_' + key].tobytes(), fixtures['out_' + key].tobytes()
        dst = ctypes.create_string_buffer(len(raw) + 8)
        r = lib.fse_host_compress(dst, len(raw), raw, len(raw))
        assert dst.raw[:r] == want
        out = ctypes.create_string_buffer(10 * r + 16)
        n = lib.fse_host_decompress(out, 10 * r, want, r)
        _check_like_the_reference(raw, want, out.raw[:n])
    zeros = fixtures['in_all_zeros'].tobytes()
    assert lib.fse_host_compress(ctypes.create_string_buffer(len(zeros) + 8), len(zeros), zeros, len(zeros)) == 1


@pytest.mark.gpu
def test_kernels_on_reference_fixtures(fixtures):
    from neural_imaging_b200.pyfse import pyfse
    raws = [fixtures['in_' + k].tobytes() for k in KEYS]
    coded = pyfse.compress_batch(raws)
    for key, raw, c in zip(KEYS, raws, coded):
        assert c == fixtures['out_' + key].tobytes(), key
        _check_like_the_reference(raw, c, pyfse.decompress(c))          # default capacity: 10 x the coded size, as in the reference test
    with pytest.raises(pyfse.FSESymbolRepetitionError):                 # test_fse.py:61-62
        pyfse.compress(fixtures['in_all_zeros'].tobytes())


class _ValidationSet:
    """The two members of helpers.dataset.Dataset that the validation helpers use."""

    def __init__(self, y):
        self.y = y
        self.count_validation = len(y)

    def next_validation_batch(self, batch_id, batch_size):
        return self.y[batch_id * batch_size:(batch_id + 1) * batch_size]


@pytest.mark.gpu
def test_validate_jpeg_and_dcn_helpers():
    """training/validation.py:19-93 (validate_jpeg, validate_dcn): the reported numbers equal the same metrics computed from the models'
    outputs with the oracle's formulas."""
    from neural_imaging_b200.models.compression import TwitterDCN
    from neural_imaging_b200.models.jpeg import JPEG
    from neural_imaging_b200.training import validation
    from oracle import ref_ops
    rs = np.random.RandomState(8)
    yy, xx = np.mgrid[0:64, 0:64]
    base = 0.5 + 0.4 * np.sin(yy / 6.0)[None, :, :, None] * np.cos(xx / 9.0)[None, :, :, None]
    y = np.clip(base + 0.05 * rs.normal(size=(4, 64, 64, 3)), 0, 1).astype(np.float32)
    data = _ValidationSet(y)
    codec = JPEG(80, 'soft')
    res = validation.validate_jpeg(codec, data, batch_size=2)
    out = codec.process(y).numpy()
    psnr = np.mean([10 * np.log10(1.0 / np.mean((y[i].astype(np.float64) - out[i]) ** 2)) for i in range(4)])
    ssim = np.mean([ref_ops.ssim_skimage(y[i], out[i]) for i in range(4)])
    assert set(res) == {'psnr', 'ssim', 'entropy'} and np.isnan(res['entropy'])
    assert abs(res['psnr'] - psnr) < 1e-3 and abs(res['ssim'] - ssim) < 1e-4 and res['psnr'] > 25
    with pytest.raises(ValueError):
        validation.validate_jpeg(object(), data)
    dcn = TwitterDCN(patch_size=64, n_features=8, seed=4)
    res = validation.validate_dcn(dcn, data)
    out, ent = dcn.process(y, return_entropy=True)
    out, ent = out.numpy(), float(ent.numpy())
    assert set(res) == {'ssim', 'psnr', 'loss', 'entropy'} and abs(res['entropy'] - ent) < 1e-6 and 0 <= ent <= 5
    assert abs(res['ssim'] - np.mean([ref_ops.ssim_skimage(y[i], out[i]) for i in range(4)])) < 1e-4
    expect_loss = 0.5 * np.sum((y.astype(np.float64) - out) ** 2) + 250 * ent          # l2_loss + entropy_weight * H (models/compression.py:91-94)
    assert abs(res["loss"] - expect_loss) < 1e-3 * expect_loss
    assert validation.validate_dcn(codec, data) is None
