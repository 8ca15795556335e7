"""l3ic byte-stream codec / FSE entropy coder (compression/codec.py:87-265, pyfse/pyfse.pyx:24-72).

Pinning chain — integer / byte work, so everything is BIT-EXACT:
  reference library (oracle/_ref/libfse_ref.so, compiled from /root/reference by oracle/Makefile)
    -> tests/golden/fse_vectors.npz (committed; produced by that library)
    -> oracle/ref_l3ic.py (plain-Python restatement)            [CPU tests]
    -> csrc/fse_core.cuh compiled for the host by g++ (harness) [CPU tests: the very source the kernels instantiate]
    -> the CUDA kernels through the C-ABI / the pyfse + compression.codec mirrors [GPU tests].
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import ref_l3ic as R


@pytest.fixture(scope='module')
def vectors():
    with np.load(os.path.join(GOLDEN, 'fse_vectors.npz')) as d:
        return {k: d[k] for k in d.files}


def _pairs(vectors):
    i = 0
    while 'in_%02d' % i in vectors:
        a, o = vectors['in_%02d' % i], vectors['out_%02d' % i]
        yield a.tobytes(), (o.tobytes() if o.dtype == np.uint8 else int(o[0]))
        i += 1


def _fuzz_strings(seed, count, sizes=(2, 3, 5, 8, 31, 64, 100, 255, 256, 257, 1000, 4096)):
    rs = np.random.RandomState(seed)
    for _ in range(count):
        n = int(rs.choice(sizes))
        kind = rs.randint(0, 7)
        if kind == 0:
            a = rs.randint(0, 256, n)
        elif kind == 1:
            a = np.clip(np.round(rs.normal(15, rs.uniform(0.2, 6), n)), 0, 31)
        elif kind == 2:
            a = np.full(n, rs.randint(0, 256))
        elif kind == 3:
            a = np.clip(np.round(rs.laplace(128, rs.uniform(0.1, 30), n)), 0, 255)
        elif kind == 4:
            a = np.full(n, 7)
            a[rs.randint(0, n, max(1, n // 50))] = rs.randint(0, 256)
        elif kind == 5:
            a = (rs.geometric(rs.uniform(0.05, 0.9), n) - 1).clip(0, 255)
        else:
            a = rs.randint(0, rs.randint(2, 40), n)
        yield a.astype(np.uint8).tobytes()


def _py_decompress(data, cap):
    try:
        return R.fse_decompress(data, cap)
    except R.FSEError:
        return -1


# ------------------------------------------------------------------------------------------------ CPU: oracle and host build of the core
def test_oracle_matches_golden_vectors(vectors):
    n_fse = 0
    for src, ref in _pairs(vectors):
        assert R.fse_compress(src) == ref
        if isinstance(ref, bytes):
            n_fse += 1
            assert R.fse_decompress(ref, len(src)) == src
            assert R.fse_decompress(ref, 4 * len(src)) == src
    assert n_fse >= 15
    cb = vectors['code_book']
    for i in range(vectors['latent'].shape[0]):
        z = vectors['latent'][i:i + 1]
        stream = vectors['stream_%d' % i].tobytes()
        assert R.l3ic_compress(z, cb) == stream
        assert np.array_equal(R.l3ic_decompress(stream, cb), cb[R.vq(z, cb)].reshape(z.shape))
    assert vectors['stream_0'][:3].tolist() == [16, 16, 32]


def test_oracle_matches_reference_library():
    lib = R.reference_library()
    if lib is None:
        pytest.skip('oracle/_ref not built (make -C oracle needs /root/reference)')
    for src in _fuzz_strings(7, 150, sizes=(2, 3, 5, 8, 31, 64, 255, 256, 257, 600)):
        ref = R.ref_compress(lib, src)
        assert R.fse_compress(src) == ref
        if isinstance(ref, bytes):
            for cap in (len(src), 4 * len(src), len(src) - 1):
                assert _py_decompress(ref, cap) == R.ref_decompress(lib, ref, cap)


@pytest.fixture(scope='module')
def host_core(tmp_path_factory):
    """The product's fse_core.cuh compiled by g++ (tests/fse_host_harness.cpp) — test infrastructure, never loaded by the product."""
    so = str(tmp_path_factory.mktemp('fse') / 'libfse_host.so')
    subprocess.check_call(['g++', '-O2', '-fPIC', '-shared', '-o', so, os.path.join(ROOT, 'tests', 'fse_host_harness.cpp')])
    lib = ctypes.CDLL(so)

    def compress(src):
        dst = ctypes.create_string_buffer(max(len(src), 16) + 8)
        r = lib.fse_host_compress(dst, max(len(src), 16), src, len(src))
        return dst.raw[:r] if r > 1 else r

    def decompress(src, cap):
        dst = ctypes.create_string_buffer(cap + 16)
        r = lib.fse_host_decompress(dst, cap, src, len(src))
        return dst.raw[:r] if r >= 0 else -1
    return compress, decompress


def test_device_core_compiled_for_host_matches_golden_and_reference(vectors, host_core):
    compress, decompress = host_core
    for src, ref in _pairs(vectors):
        assert compress(src) == ref
        if isinstance(ref, bytes):
            assert decompress(ref, len(src)) == src
    lib = R.reference_library()
    rs = np.random.RandomState(3)
    n_checked = 0
    for src in _fuzz_strings(11, 400, sizes=(2, 3, 4, 5, 8, 31, 64, 100, 255, 256, 257, 1000, 4096, 16384, 65535)):
        ref = R.ref_compress(lib, src) if lib is not None else (R.fse_compress(src) if len(src) <= 1000 else None)
        if ref is None:
            continue
        assert compress(src) == ref
        n_checked += 1
        if isinstance(ref, bytes) and lib is not None:
            for cap in (len(src), 4 * len(src), len(src) - 1, len(src) + 3):
                assert decompress(ref, cap) == R.ref_decompress(lib, ref, cap)
            bad = bytearray(ref)                       # a corrupted / truncated stream gets the reference's verdict, byte for byte
            bad[rs.randint(0, len(bad))] ^= 1 << rs.randint(0, 8)
            bad = bytes(bad[:rs.randint(1, len(bad) + 1)])
            assert decompress(bad, 4 * len(src)) == R.ref_decompress(lib, bad, 4 * len(src))
    assert n_checked >= 100


def test_c_abi_exports_codec_entry_points():
    from neural_imaging_b200 import _lib
    protos = _lib.parse_header()
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in ('ni_fse_compress_batch', 'ni_fse_decompress_batch', 'ni_l3ic_encode', 'ni_l3ic_decode'):
        assert name in protos and hasattr(dll, name)


# ------------------------------------------------------------------------------------------------ GPU: kernels through the public mirrors
def _expect(ref, pyfse):
    if isinstance(ref, bytes):
        return ref
    return {0: pyfse.FSENotCompressibleError, 1: pyfse.FSESymbolRepetitionError}.get(ref, pyfse.FSEException)


def _same(got, want):
    return got == want if isinstance(want, bytes) else type(got) is want


@pytest.mark.gpu
def test_pyfse_batch_matches_golden_and_oracle(vectors):
    from neural_imaging_b200.pyfse import pyfse
    lib = R.reference_library()
    srcs, refs = zip(*_pairs(vectors))
    extra = list(_fuzz_strings(21, 300, sizes=(2, 3, 4, 5, 8, 31, 64, 100, 255, 256, 257, 1000, 4096, 16384) if lib else (2, 3, 5, 8, 64, 256, 257)))
    extra_refs = [R.ref_compress(lib, s) if lib else R.fse_compress(s) for s in extra]
    srcs, refs = list(srcs) + extra, list(refs) + extra_refs
    got = pyfse.compress_batch(srcs)
    for s, g, r in zip(srcs, got, refs):
        assert _same(g, _expect(r, pyfse)), (len(s), r if not isinstance(r, bytes) else len(r))
    coded = [(s, r) for s, r in zip(srcs, refs) if isinstance(r, bytes)]
    assert len(coded) > 100
    # decoding: exact capacity, the codec's 4x capacity, and a capacity that is one byte short
    for scale, delta in ((1, 0), (4, 0), (1, -1)):
        for (s, r), g in zip(coded, _by_cap(pyfse, coded, scale, delta)):
            cap = scale * len(s) + delta
            want = (R.ref_decompress(lib, r, cap) if lib else _py_decompress(r, cap)) if cap > 0 else None
            if want is None:
                continue
            assert (g == want) if isinstance(want, bytes) else isinstance(g, pyfse.FSEException), (len(s), cap)
    # the one-string API and its exceptions (pyfse.pyx:24-72)
    one = coded[0][0]
    assert pyfse.compress(one) == coded[0][1] and pyfse.decompress(coded[0][1], len(one)) == one and pyfse.decompress(coded[0][1]) == one
    # edges: empty and one-byte strings, and the largest layer the l3ic length table can describe (65,535 symbols: un-staged kernel path)
    big = np.clip(np.round(np.random.RandomState(1).normal(100, 9, 65535)), 0, 255).astype(np.uint8).tobytes()
    edge = pyfse.compress_batch([b'', b'a', big])
    assert isinstance(edge[0], pyfse.FSENotCompressibleError) and isinstance(edge[1], pyfse.FSENotCompressibleError)
    if lib is not None:
        assert edge[2] == R.ref_compress(lib, big)
    assert len(edge[2]) < 50000 and pyfse.decompress(edge[2], 65535) == big
    assert pyfse.compress_batch([]) == [] and pyfse.decompress_batch([]) == []
    with pytest.raises(pyfse.FSESymbolRepetitionError):
        pyfse.compress(b'\x05' * 100)
    with pytest.raises(pyfse.FSENotCompressibleError):
        pyfse.compress(bytes(range(256)))
    with pytest.raises(pyfse.FSEException):
        pyfse.decompress(b'\x00\x01\x02\x03\x04\x05', 100)


def _by_cap(pyfse, coded, scale, delta):
    """decompress_batch takes one capacity per call: group the strings by their capacity."""
    out = [None] * len(coded)
    caps = {}
    for i, (s, _) in enumerate(coded):
        caps.setdefault(scale * len(s) + delta, []).append(i)
    for cap, idx in caps.items():
        if cap <= 0:
            continue
        for i, g in zip(idx, pyfse.decompress_batch([coded[i][1] for i in idx], max_length=cap)):
            out[i] = g
    return out


@pytest.mark.gpu
def test_corrupt_streams_get_the_reference_verdict():
    from neural_imaging_b200.pyfse import pyfse
    lib = R.reference_library()
    if lib is None:
        pytest.skip('needs oracle/_ref (undefined-behaviour-free comparison of corrupt streams is only meaningful against the library)')
    rs = np.random.RandomState(5)
    bad, caps = [], []
    for src in _fuzz_strings(31, 120, sizes=(64, 256, 1000)):
        ref = R.ref_compress(lib, src)
        if not isinstance(ref, bytes):
            continue
        b = bytearray(ref)
        b[rs.randint(0, len(b))] ^= 1 << rs.randint(0, 8)
        bad.append(bytes(b[:rs.randint(1, len(b) + 1)]))
        caps.append(4 * len(src))
    for cap in sorted(set(caps)):
        group = [b for b, c in zip(bad, caps) if c == cap]
        for b, g in zip(group, pyfse.decompress_batch(group, max_length=cap)):
            want = R.ref_decompress(lib, b, cap)
            assert (g == want) if isinstance(want, bytes) else isinstance(g, pyfse.FSEException)


@pytest.mark.gpu
def test_l3ic_streams_match_golden_and_oracle(vectors):
    from neural_imaging_b200.compression import codec
    cb = vectors['code_book']
    z = vectors['latent']
    streams = codec.encode_latent(z, cb)
    for i, s in enumerate(streams):
        assert s == vectors['stream_%d' % i].tobytes()
    back = codec.decode_latent(streams, z.shape[1:], cb).cpu().numpy()
    assert np.array_equal(back, cb[R.vq(z, cb)].reshape(z.shape))
    # other shapes and code books; every image is its own stream, compared with the restated container
    rs = np.random.RandomState(9)
    for (n, h, w, c), book in (((5, 8, 8, 16), np.arange(-15, 17)), ((2, 32, 32, 8), np.arange(-7, 9)), ((3, 4, 6, 3), np.linspace(-1, 1, 7)),
                               ((2, 64, 64, 4), np.arange(-127, 129))):
        book = book.astype(np.float32)
        z = rs.normal(0, rs.uniform(0.3, 3), (n, h, w, c)).astype(np.float32)
        z[0, :, :, 0] = book[1]                                         # run
        z[-1, :, :, -1] = book[rs.randint(0, len(book), (h, w))]        # flat
        streams = codec.encode_latent(z, book)
        for i, s in enumerate(streams):
            assert s == R.l3ic_compress(z[i:i + 1], book), (n, h, w, c, i)
        back = codec.decode_latent(streams, (h, w, c), book).cpu().numpy()
        assert np.array_equal(back, book[R.vq(z, book)].reshape(z.shape))
        assert np.array_equal(R.l3ic_decompress(streams[0], book), back[:1])
    # errors (compression/codec.py raises / numpy raises): truncated stream, wrong model shape, too many code-book entries
    with pytest.raises(codec.L3ICError):
        codec.decode_latent([streams[0][:len(streams[0]) // 2]], (h, w, c), book)
    with pytest.raises(codec.L3ICError):
        codec.decode_latent([streams[0]], (h, w, c + 1), book)
    with pytest.raises(Exception):
        codec.encode_latent(z, np.arange(300, dtype=np.float32))


@pytest.mark.gpu
def test_codec_through_the_dcn_model(tmp_path):
    """codec.compress / decompress / simulate_compression with a TwitterDCN (compression/codec.py:18-26,87-265): the byte stream decodes
    to exactly the image the model itself reconstructs from its quantised latent, for one image and for a batch."""
    from neural_imaging_b200.compression import codec
    from neural_imaging_b200.models.compression import TwitterDCN
    rs = np.random.RandomState(2)
    model = TwitterDCN(patch_size=64, n_features=8, seed=5)
    x = rs.uniform(size=(4, 64, 64, 3)).astype(np.float32)
    z = model.compress(x).numpy()
    streams = codec.compress_batch(x, model)
    assert len(streams) == 4 and all(s[:3] == bytes(model.latent_shape) for s in streams)
    for i, s in enumerate(streams):
        assert s == R.l3ic_compress(z[i:i + 1], model.get_codebook())
        assert codec.compress(x[i], model) == s
    y = codec.decompress_batch(streams, model).numpy()
    assert np.array_equal(y, model.decompress(z).numpy())
    y0, n_bytes = codec.simulate_compression(x[:1], model)
    assert n_bytes == len(streams[0]) and np.array_equal(y0, y[:1])
    with pytest.raises(ValueError):
        codec.decompress(12345, model)
    # per-image statistics through the byte stream (codec.py:29-55)
    y_b, stats = codec.compress_n_stats(x[:2], model)
    assert np.allclose(y_b, y[:2], atol=1e-5) and stats['bytes'].tolist() == [len(s) for s in streams[:2]]
    assert np.allclose(stats['bpp'], [8 * len(s) / 64 / 64 for s in streams[:2]]) and all(0 < v <= 5 for v in stats['entropy'])
    assert all(-1 <= v <= 1 for v in stats['ssim']) and all(v > 0 for v in stats['psnr'])
    # codec.restore / tfmodel.restore from a snapshot directory (models/tfmodel.py:16-83, codec.py:275-291)
    import json
    model.save_model(str(tmp_path))
    with open(os.path.join(str(tmp_path), 'progress.json'), 'w') as f:
        json.dump({'codec': {'model': 'TwitterDCN', 'args': model.get_hyperparameters(), 'performance': {}}}, f)
    restored = codec.restore(str(tmp_path), patch_size=64)
    assert codec.compress(x[0], restored) == streams[0]
    assert np.allclose(codec.decompress(streams[1], restored), y[1:2], atol=1e-5)
    with pytest.raises(ValueError):
        codec.restore('no-such-preset')
    with pytest.raises(ValueError):
        codec.decompress(streams[0])                    # no model: looks for the preset '8c' like the reference, which is absent here


@pytest.mark.gpu
def test_large_batch_round_trip():
    """BASELINE-sized batch (1280 images x 32 layers of 16 x 16 = 40,960 streams in one launch): decode(encode(z)) == z."""
    import torch
    from neural_imaging_b200.compression import codec
    g = torch.Generator(device='cpu').manual_seed(3)
    z = torch.clamp(torch.round(torch.randn((1280, 16, 16, 32), generator=g) * 1.3), -15, 16)
    book = np.arange(-15, 17, dtype=np.float32)
    streams = codec.encode_latent(z, book)
    back = codec.decode_latent(streams, (16, 16, 32), book).cpu()
    assert torch.equal(back, z)
    total = sum(len(s) for s in streams)
    assert total < 0.5 * z.numel()                       # ~2.4 bits per symbol of entropy: well under a byte per latent value
    for i in (0, 777, 1279):
        assert streams[i] == R.l3ic_compress(z[i:i + 1].numpy(), book)


# ------------------------------------------------------------------------------------------------ CPU: the coder's stages on their own
def _count_tables(seed, n_cases):
    """Histograms FSE_compress itself rarely produces: they drive the fallback normalisation through all of its branches ("risk of rounding
    to zero", "all values are pretty poor", "total == 0") and the header writer through long zero runs."""
    rs = np.random.RandomState(seed)
    ones = np.ones(43, dtype=np.uint32)
    yield 42, ones, 6                                                   # 43 symbols seen once, 64 cells: every symbol is "poor"
    for tl in (6, 7):
        for n_once in (43, 50, 86, 100):
            c = np.zeros(128, np.uint32)
            c[:n_once] = 1
            c[127] = 1                                                  # unused symbols in between: the "total == 0" hand-out loop
            yield 127, c, tl
    for _ in range(n_cases):
        m = int(rs.randint(1, 256))
        kind = rs.randint(0, 8)
        if kind == 0:
            count = rs.randint(0, 50, m + 1)
        elif kind == 1:
            count = (rs.rand(m + 1) < 0.3) * rs.randint(1, 4, m + 1)
        elif kind == 2:
            count = np.ones(m + 1, np.int64)
            count[rs.randint(0, m + 1)] = rs.randint(1, 100000)
        elif kind == 3:
            count = rs.geometric(0.01, m + 1) * (rs.rand(m + 1) < 0.5)
        elif kind == 4:
            count = np.full(m + 1, rs.randint(1, 5))
        elif kind == 5:
            count = (rs.zipf(1.3, m + 1) % 5000) * (rs.rand(m + 1) < 0.7)
        else:
            if kind == 7:
                m = int(rs.randint(20, 256))
            count = np.zeros(m + 1, np.int64)
            idx = rs.permutation(m + 1)
            k = rs.randint(8 if kind == 7 else 0, m + 1)
            j = rs.randint(0, 3) if kind == 7 else rs.randint(0, max(1, m + 1 - k))
            nb = 0 if kind == 7 else rs.randint(0, 4)
            a = rs.randint(1, 3)
            count[idx[:k]] = a
            count[idx[k:k + j]] = a * rs.randint(1, 4)
            count[idx[k + j:k + j + nb]] = rs.choice([50, 500, 5000, 50000, 1000000])
        count = np.asarray(count, dtype=np.uint32)
        if count[m] == 0:
            count[m] = 1
        if int(count.sum()) >= 2:
            yield m, count, int(rs.randint(5, 10 if kind == 7 else 13))


def test_normalisation_and_header_stages_match_reference_library(tmp_path):
    lib = R.reference_library()
    if lib is None:
        pytest.skip('oracle/_ref not built (make -C oracle needs /root/reference)')
    so = str(tmp_path / 'libfse_host.so')
    subprocess.check_call(['g++', '-O2', '-fPIC', '-shared', '-o', so, os.path.join(ROOT, 'tests', 'fse_host_harness.cpp')])
    host = ctypes.CDLL(so)
    for fn in (lib.FSE_normalizeCount, lib.FSE_writeNCount, lib.FSE_readNCount):
        fn.restype = ctypes.c_size_t
    agree = rejected = 0
    rs = np.random.RandomState(1)
    for m, count, tl in _count_tables(17, 2500):
        total = int(count.sum())
        c = (ctypes.c_uint32 * 256)(*([int(v) for v in count] + [0] * (255 - m)))
        n0, n1 = (ctypes.c_int16 * 256)(), (ctypes.c_int16 * 256)()
        args = (ctypes.c_uint(tl), c, ctypes.c_size_t(total), ctypes.c_uint(m))
        pid = os.fork()                  # the reference divides by zero on a few out-of-domain tables: probe each one in a child first
        if pid == 0:
            lib.FSE_normalizeCount(n0, *args)
            os._exit(0)
        if os.waitpid(pid, 0)[1] != 0:
            continue
        r0 = lib.FSE_normalizeCount(n0, *args)
        r1 = host.fse_host_normalize(n1, tl, c, total, m)
        failed = bool(lib.FSE_isError(ctypes.c_size_t(r0)))
        assert failed == (r1 < 0), (tl, m, total)
        if failed:
            rejected += 1
            continue
        assert r0 == r1
        if r0 == 0:
            continue
        assert list(n0)[:m + 1] == list(n1)[:m + 1], (tl, m, total)
        agree += 1
        b0, b1 = ctypes.create_string_buffer(600), ctypes.create_string_buffer(600)
        w0 = lib.FSE_writeNCount(b0, ctypes.c_size_t(512), n0, ctypes.c_uint(m), ctypes.c_uint(tl))
        w1 = host.fse_host_write_ncount(b1, 512, n1, m, tl)
        assert not lib.FSE_isError(ctypes.c_size_t(w0)) and w0 == w1 and b0.raw[:w0] == b1.raw[:w1]
        for cut in (w0, w0 + 5, max(1, w0 - 1), 3, 2):          # exact, padded with noise, truncated, shorter than the 4-byte minimum
            hdr = b0.raw[:cut] if cut <= w0 else b0.raw[:w0] + bytes(rs.randint(0, 256, cut - w0).astype(np.uint8))
            a0, a1 = (ctypes.c_int16 * 256)(), (ctypes.c_int16 * 256)()
            ms0, tl0, ms1, tl1 = ctypes.c_uint(255), ctypes.c_uint(0), ctypes.c_uint32(255), ctypes.c_uint32(0)
            q0 = lib.FSE_readNCount(a0, ctypes.byref(ms0), ctypes.byref(tl0), hdr, ctypes.c_size_t(len(hdr)))
            q1 = host.fse_host_read_ncount(a1, ctypes.byref(ms1), ctypes.byref(tl1), hdr, len(hdr))
            bad = bool(lib.FSE_isError(ctypes.c_size_t(q0)))
            assert bad == (q1 < 0), (cut, w0)
            if not bad:
                assert q0 == q1 and ms0.value == ms1.value and tl0.value == tl1.value and list(a0) == list(a1)
    assert agree > 1000 and rejected > 100
