"""SURVEY 8f N4 — Keras .h5 weight-file interchange (reference models/tfmodel.py:150-182) without h5py / libhdf5.

What pins what:
* the READER against a file written by the real HDF5 library: tests/golden/hdf5_matlab73_testdouble.mat (a MATLAB 7.3 MAT-file = HDF5 with
  a 512-byte user block; SciPy's test data `testhdf5_7.4_GLNX86.mat`, whose documented content is theta = pi/4 * arange(9));
* the WRITER against the reader and, byte for byte, against the message encodings found in that libhdf5-written file;
* the Keras layout (variable ORDER and storage SHAPES) against tests/golden/keras_weight_order.json, produced by executing the
  reference's own model constructors (tests/golden/make_keras_weight_order.py).
CPU only: the models are built with nn.HOST_ONLY (no device buffers)."""
import json
import os
import struct

import numpy as np
import pytest

from neural_imaging_b200.helpers import h5lite

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MAT = os.path.join(GOLDEN, 'hdf5_matlab73_testdouble.mat')


def test_reader_on_a_libhdf5_written_file():
    f = h5lite.File(MAT)
    assert (f.O, f.L, f.base) == (8, 8, 512)                       # super-block version 0 behind MATLAB's 512-byte user block
    assert f.keys() == ['testdouble'] and 'testdouble' in f and 'nope' not in f
    d = f['testdouble']
    assert d.shape == (9, 1) and d.dtype == np.dtype('<f8')
    assert np.array_equal(d.read().ravel(), np.pi / 4 * np.arange(9))      # scipy/io/matlab/tests/test_mio.py: theta
    assert d.attrs == {'MATLAB_class': np.bytes_(b'double')}
    with pytest.raises(KeyError):
        f['testdouble/x']
    with pytest.raises(h5lite.H5Error):
        h5lite.File(b'not an hdf5 file at all' * 100)
    with pytest.raises(h5lite.H5Error):                              # truncated file: addresses beyond the end are reported, not read
        h5lite.File(open(MAT, 'rb').read()[:1500])['testdouble'].read()


def test_writer_emits_the_encodings_libhdf5_wrote():
    """Same dataset + attribute written here: data-space, datatype and attribute messages are byte-identical to the library's (the
    attribute's string padding flag aside: NULLPAD as h5py maps numpy 'S', MATLAB used NULLTERM), the group structures have the
    library's sizes and constants."""
    ref = h5lite.File(MAT)['testdouble']
    data = h5lite.write(None, {'testdouble': np.pi / 4 * np.arange(9.).reshape(9, 1)}, {'testdouble': {'MATLAB_class': np.bytes_(b'double')}})
    mine = h5lite.File(data)['testdouble']
    rm = {t: bytes(b) for t, _, b in ref._msgs}
    mm = {t: bytes(b) for t, _, b in mine._msgs}
    assert mm[0x0001] == rm[0x0001] and mm[0x0003] == rm[0x0003]
    a, b = bytearray(mm[0x000C]), bytearray(rm[0x000C])
    assert a[25] == 0x01 and b[25] == 0x00                          # string padding type: null-pad vs null-terminate
    a[25] = b[25]
    assert a == b
    assert np.array_equal(mine.read(), ref.read())
    # structure constants (HDF5 file-format specification, version-0 super-block / version-1 B-tree / symbol node / local heap)
    assert data[:8] == h5lite.SIGNATURE and data[8] == 0 and data[13:15] == b'\x08\x08'
    assert struct.unpack_from('<HH', data, 16) == (4, 16)            # group leaf / internal node K: the library's defaults
    eof, = struct.unpack_from('<Q', data, 40)
    assert eof == len(data)
    for sig, size in ((b'TREE', 24 + 33 * 8 + 32 * 8), (b'SNOD', 8 + 8 * 40), (b'HEAP', 32)):
        at = data.index(sig)
        assert at % 8 == 0 and at + size <= len(data)


def test_roundtrip_groups_attributes_dtypes():
    rs = np.random.RandomState(3)
    tree = {'g%02d' % i: {'w': rs.normal(size=(3, i + 1)).astype(np.float32)} for i in range(40)}     # 40 members: five symbol nodes
    tree['scalar'] = np.float32(2.5)
    tree['ints'] = np.arange(-5, 5, dtype=np.int64).reshape(2, 5)
    tree['bytes'] = np.arange(7, dtype=np.uint8)
    tree['f64'] = rs.normal(size=(2, 2, 2))
    tree['empty'] = {}
    tree['deep'] = {'a': {'b': {'c:0': np.ones((1, 1, 2, 3), np.float32)}}}
    attrs = {'': {'names': np.array([b'alpha', b'be', b'gamma_long_name'], dtype='S'), 'version': np.bytes_(b'2.2.4-tf'), 'n': np.int32(7)},
             'deep/a': {'weight_names': np.array([b'b/c:0'], dtype='S')}, 'ints': {'scale': np.float64(0.5)}}
    f = h5lite.File(h5lite.write(None, tree, attrs))
    assert sorted(f.keys()) == sorted(tree)
    for i in range(40):
        assert np.array_equal(f['g%02d/w' % i].read(), tree['g%02d' % i]['w'])
    assert f['scalar'].shape == () and f['scalar'].read() == np.float32(2.5)
    for k in ('ints', 'bytes', 'f64'):
        a = f[k].read()
        assert a.dtype == tree[k].dtype and np.array_equal(a, tree[k])
    assert f['empty'].keys() == []
    assert np.array_equal(f['deep/a/b/c:0'].read(), tree['deep']['a']['b']['c:0']) and f['deep']['a']['b'].keys() == ['c:0']
    assert list(f.attrs['names']) == [b'alpha', b'be', b'gamma_long_name'] and f.attrs['version'] == b'2.2.4-tf' and f.attrs['n'] == 7
    assert list(f['deep/a'].attrs['weight_names']) == [b'b/c:0'] and f['ints'].attrs['scale'] == 0.5
    seen = []
    f.visit(lambda name, node: seen.append(name))
    assert 'deep/a/b/c:0' in seen and 'g39/w' in seen
    with pytest.raises(h5lite.H5Error):
        h5lite.write(None, {'a/b': np.zeros(1)})


def _host_model(ctor, **kw):
    from neural_imaging_b200 import nn
    old = nn.HOST_ONLY
    nn.HOST_ONLY = True
    try:
        return ctor(**kw)
    finally:
        nn.HOST_ONLY = old


def _cases():
    from neural_imaging_b200.models import compression, forensics, pipelines
    return {'UNet': (pipelines.UNet, {}), 'INet': (pipelines.INet, {}), 'DNet': (pipelines.DNet, {}),
            'FAN': (forensics.FAN, dict(n_classes=5)), 'FAN_dense2': (forensics.FAN, dict(n_classes=3, n_dense=2)),
            'TwitterDCN': (compression.TwitterDCN, {})}


@pytest.mark.parametrize('key', ['UNet', 'INet', 'DNet', 'FAN', 'FAN_dense2', 'TwitterDCN'])
def test_keras_weight_file_layout_matches_the_executed_reference(key, tmp_path):
    """save_model writes <dir>/<scoped name>/<class>.h5 whose variables — walked the way Keras' load_weights_from_hdf5_group walks
    them — have the order and the storage shapes of the reference model's own `model.weights`; load_model restores them into a
    differently initialised model."""
    from neural_imaging_b200.models.tfmodel import _to_keras
    with open(os.path.join(GOLDEN, 'keras_weight_order.json')) as fh:
        expected = json.load(fh)[key]['variables']
    ctor, kw = _cases()[key]
    m = _host_model(ctor, seed=11, **kw) if 'seed' in ctor.__init__.__code__.co_varnames else _host_model(ctor, **kw)
    m.save_model(str(tmp_path))
    path = os.path.join(str(tmp_path), m.scoped_name, m.class_name.lower() + '.h5')
    assert os.path.isfile(path)
    f = h5lite.File(path)
    assert f.attrs['backend'] == b'tensorflow' and b'tf' in bytes(f.attrs['keras_version'])
    found = []
    for layer in f.attrs['layer_names']:
        g = f[layer.decode()]
        for wn in g.attrs['weight_names']:
            assert wn.decode().startswith(layer.decode() + '/') and wn.decode().endswith(':0')
            found.append([wn.decode().rsplit('/', 1)[1], list(g[wn.decode()].shape)])
    assert found == expected
    # a model with other initial values takes the file's values (order-based, layouts converted back)
    m2 = _host_model(ctor, seed=12, **kw) if 'seed' in ctor.__init__.__code__.co_varnames else _host_model(ctor, **kw)
    for p in m2._store.params:
        if p.trainable:
            p.init = p.init + 1.0
    m2.load_model(str(tmp_path))
    for p, q in zip(m._store.params, m2._store.params):
        assert p.name == q.name and (p.keras == 'internal' or np.array_equal(p.init, q.init)), p.name
    # the stored arrays ARE the Keras-layout views of the product's parameters
    vals = [f[l.decode()][w.decode()].read() for l in f.attrs['layer_names'] for w in f[l.decode()].attrs['weight_names']]
    for p, a in zip([p for p in m._store.params if p.keras != 'internal'], vals):
        assert np.array_equal(_to_keras(p, p.init.reshape(p.shape)), a)


def test_keras_layout_converters_and_errors(tmp_path):
    from neural_imaging_b200.models import forensics, tfmodel
    m = _host_model(forensics.FAN, n_classes=5)
    # Conv2DTranspose rule: W1x1[ci, (a*2+b)*F + f] = K[a, b, f, ci]   (nn.py docstring; golden-tested in tfgraph_common)
    class P(object):
        keras, name = 'conv2d_transpose', 'dct/kernel'
        shape = (1, 1, 6, 4 * 5)
    k = np.random.RandomState(0).normal(size=(2, 2, 5, 6)).astype(np.float32)
    w = tfmodel._from_keras(P, k)
    for a in range(2):
        for b in range(2):
            assert np.array_equal(w[0, 0, :, (a * 2 + b) * 5:(a * 2 + b) * 5 + 5], k[a, b].T)
    assert np.array_equal(tfmodel._to_keras(P, w), k)
    # a weight file of another architecture is rejected by count or by shape
    m.save_model(str(tmp_path))
    other = _host_model(forensics.FAN, n_classes=4)
    with pytest.raises(ValueError):
        other.load_model(str(tmp_path))
    deeper = _host_model(forensics.FAN, n_classes=5, n_dense=1)
    with pytest.raises(ValueError):
        deeper.load_model(str(tmp_path))
    with pytest.raises(FileNotFoundError):
        m.load_model(os.path.join(str(tmp_path), 'missing'))
    # snapshots of earlier versions of this stack (.npz) are still read
    legacy = os.path.join(str(tmp_path), 'legacy', 'fan')
    os.makedirs(legacy)
    np.savez(os.path.join(legacy, 'fan.npz'), **{p.name: p.init + 2.0 for p in m._store.params})
    m3 = _host_model(forensics.FAN, n_classes=5)
    m3.load_model(os.path.join(str(tmp_path), 'legacy'))
    assert all(np.array_equal(p.init + 2.0, q.init) for p, q in zip(m._store.params, m3._store.params))


def test_loader_accepts_what_keras_writes_around_the_weights(tmp_path):
    """Keras writes a group for EVERY layer (weightless ones carry an empty float64 `weight_names` attribute), `model.save()` nests the
    same structure under /model_weights, and layer names differ from this stack's: the loader must skip the former, find the latter and
    match purely by order."""
    from neural_imaging_b200.models import forensics
    m = _host_model(forensics.FAN, n_classes=5)
    m.save_model(str(tmp_path))
    f = h5lite.File(os.path.join(str(tmp_path), 'fan', 'fan.h5'))
    tree, attrs, names = {}, {}, []
    k = 0
    for layer in f.attrs['layer_names']:
        g = f[layer.decode()]
        new = 'renamed_layer_%d' % k                                          # Keras' automatic names depend on the session
        k += 1
        wn = []
        tree[new] = {new: {}}
        for w in g.attrs['weight_names']:
            leaf = w.decode().rsplit('/', 1)[1]
            tree[new][new][leaf] = g[w.decode()].read()
            wn.append('{}/{}'.format(new, leaf).encode())
        names.append(new.encode())
        attrs['model_weights/' + new] = {'weight_names': np.array(wn, dtype='S')}
        pool = 'max_pooling2d_%d' % k                                          # a weightless layer after every weighted one
        tree[pool] = {}
        names.append(pool.encode())
        attrs['model_weights/' + pool] = {'weight_names': np.array([])}
    attrs['model_weights'] = {'layer_names': np.array(names, dtype='S'), 'backend': np.bytes_(b'tensorflow'), 'keras_version': np.bytes_(b'2.2.4-tf')}
    attrs[''] = {'keras_version': np.bytes_(b'2.2.4-tf'), 'model_config': np.bytes_(b'{}')}
    out = os.path.join(str(tmp_path), 'saved', 'fan')
    os.makedirs(out)
    h5lite.write(os.path.join(out, 'fan.h5'), {'model_weights': tree}, attrs)
    m2 = _host_model(forensics.FAN, n_classes=5, seed=99)
    m2.load_model(os.path.join(str(tmp_path), 'saved'))
    assert all(np.array_equal(p.init, q.init) for p, q in zip(m._store.params, m2._store.params))
