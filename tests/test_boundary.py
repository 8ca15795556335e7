"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares, the
product never touches the oracle, and the host-side API mirrors the reference's error behaviour (no GPU needed)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'neural_imaging_b200')


def test_library_exports_every_header_symbol():
    from neural_imaging_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 40
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    dll = ctypes.CDLL(_lib.LIB_PATH)
    missing = [name for name in protos if not hasattr(dll, name)]
    assert not missing, 'header declares symbols the library does not export: {}'.format(missing)
    dll.ni_version.restype = ctypes.c_int
    assert dll.ni_version() >= 100
    assert _lib.lib().ni_last_error() is not None


def test_conv_desc_layout_matches_header():
    from neural_imaging_b200 import _lib
    src = open(_lib.HEADER).read()
    body = re.search(r'typedef struct ni_conv_desc \{(.*?)\} ni_conv_desc;', src, flags=re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = []
    for decl in body.split(';'):
        decl = decl.strip()
        if decl:
            names += [n.strip() for n in decl.split(None, 1)[1].split(',')]
    assert names == [f[0] for f in _lib.ConvDesc._fields_]
    csrc = open(os.path.join(PKG, 'csrc', 'conv_desc.h')).read()
    cbody = re.search(r'typedef struct ni_conv_desc \{(.*?)\} ni_conv_desc;', csrc, flags=re.S).group(1)
    cbody = re.sub(r'//[^\n]*', '', cbody)
    cnames = []
    for decl in cbody.split(';'):
        decl = decl.strip()
        if decl:
            cnames += [n.strip() for n in decl.split(None, 1)[1].split(',')]
    assert cnames == names


def test_product_never_imports_oracle():
    pat = re.compile(r'^\s*(from|import)\s+oracle\b|/oracle/|oracle\.', re.M)
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), '{} references the oracle'.format(os.path.join(dirpath, f))


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from neural_imaging_b200.models import jpeg
    codec = jpeg.JPEG(50, 'soft')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        codec.process(np.zeros((1, 8, 8, 3), np.float32))


def test_jpeg_api_errors():
    from neural_imaging_b200.models import jpeg
    with pytest.raises(ValueError):
        jpeg.JPEG(50, 'bogus')
    with pytest.raises(ValueError):
        jpeg.DifferentiableJPEG(quality=0)
    with pytest.raises(ValueError):
        jpeg.DifferentiableJPEG(50, 'round')
    with pytest.raises(ValueError, match='Invalid or unspecified JPEG quality'):
        jpeg.JPEG(None, 'soft').process(np.zeros((1, 8, 8, 3), np.float32))
    assert jpeg.is_valid_quality(50) and jpeg.is_valid_quality((50, 90)) and not jpeg.is_valid_quality(101)
    c = jpeg.JPEG(75, 'sin')
    assert repr(c) == 'JPEG(quality=75,codec="sin",trainable=False)'
    assert c.summary() == 'JPEG (sin) QF=75' and c.estimate_qf() == 75
    assert jpeg.JPEG((50, 90))._quality_mode() == 'QF~[50,90]'
    assert np.isnan(jpeg.JPEG.loss(None, None, float('nan')))          # the workflow's NaN entropy lands in sample_weight


def test_paramspec_semantics():
    from neural_imaging_b200.helpers.paramspec import ParamSpec
    h = ParamSpec({'n': (5, int, (2, 6)), 'act': ('leaky_relu', str, {'leaky_relu', 'relu'}), 'flag': (False, bool, None)})
    h.update(n=3, act='relu')
    assert h.n == 3 and h.act == 'relu' and h.flag is False
    assert h.to_json() == {'n': 3, 'act': 'relu', 'flag': False} and h.changed_params() == {'n': 3, 'act': 'relu'}
    for bad in ({'n': 7}, {'act': 'gelu'}, {'bogus': 1}, {'n': float('nan')}):
        with pytest.raises(ValueError):
            h.update(**bad)
    with pytest.raises(ValueError):
        h.n = 4
    with pytest.raises(ValueError):
        ParamSpec({'x': (1, int)})


def test_alias_install():
    code = ('import sys; sys.path.insert(0, %r); import neural_imaging_b200 as ni; ni.install_aliases(); '
            'from models import jpeg; from helpers import kernels; from compression import jpeg_helpers; '
            'from workflows import manipulation_classification as mc; '
            'print(jpeg.JPEG.__module__, mc.ManipulationClassification.__name__)') % ROOT
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert 'neural_imaging_b200.models.jpeg ManipulationClassification' in out.stdout


def test_reference_import_lines_for_the_codec_resolve():
    """`from pyfse import pyfse` and `from compression import codec` (compression/codec.py:9-10, test_dcn.py) land on the device mirrors."""
    code = ('import sys; sys.path.insert(0, %r); import neural_imaging_b200 as ni; ni.install_aliases(); '
            'from pyfse import pyfse; from compression import codec; from training import validation; '
            'print(pyfse.__name__, codec.L3ICError.__name__, issubclass(pyfse.FSESymbolRepetitionError, pyfse.FSEException), '
            'hasattr(validation, "validate_dcn") and hasattr(validation, "validate_jpeg"))') % ROOT
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert 'neural_imaging_b200.pyfse.pyfse L3ICError True True' in out.stdout


def test_tfmodel_restore_host_logic(tmp_path, monkeypatch):
    """models/tfmodel.py:16-83 `restore(dir_name, module, key, patch_size, restore_perf, fetch_stats)`: training-log discovery, key
    lookup, tuple strings, presets and their errors — with a stand-in model class (no device needed)."""
    import json
    import types
    from neural_imaging_b200.models import tfmodel

    class Dummy:
        def __init__(self, **kw):
            self.kw, self.loaded = kw, None
            self.performance = {'loss': {'training': [], 'validation': []}}

        def load_model(self, dirname):
            self.loaded = dirname
    module = types.ModuleType('pkg.compression')
    module.Dummy = Dummy
    d = tmp_path / 'run' / 'dummy'
    d.mkdir(parents=True)
    log = {'codec': {'model': 'Dummy', 'args': {'n_features': 8, 'shape': '(4, 4)', 'name': 'x'},
                     'performance': {'loss': {'training': [3.0, 2.0], 'validation': [2.5, 1.23456]}, 'ssim': {'training': [0.5], 'validation': []}}}}
    (d / 'progress.json').write_text(json.dumps(log))
    m = tfmodel.restore(str(tmp_path / 'run'), module, key='codec', patch_size=64)
    assert isinstance(m, Dummy) and m.kw == {'n_features': 8, 'shape': (4, 4), 'name': 'x', 'patch_size': 64} and m.loaded == str(tmp_path / 'run')
    m, stats = tfmodel.restore(str(tmp_path / 'run'), module, key='codec', restore_perf=True, fetch_stats=True)
    assert m.kw['patch_size'] is None and stats == {'loss': 1.235, 'ssim': 0.5}
    _, stats = tfmodel.restore(str(tmp_path / 'run'), module, key='codec', fetch_stats=True)
    assert stats == {}                                       # without restore_perf the statistics are those of the fresh model
    with pytest.raises(ValueError):
        tfmodel.restore(None, module)
    with pytest.raises(KeyError):
        tfmodel.restore(str(tmp_path / 'run'), module, key='nip')
    empty = tmp_path / 'empty'
    empty.mkdir()
    with pytest.raises(FileNotFoundError):
        tfmodel.restore(str(empty), module)
    monkeypatch.chdir(tmp_path)
    with pytest.raises(ValueError, match='presets not available'):
        tfmodel.restore('32c', module)
    (tmp_path / 'config' / 'presets').mkdir(parents=True)
    (tmp_path / 'config' / 'presets' / 'compression.json').write_text(json.dumps({'32c': str(tmp_path / 'run')}))
    assert isinstance(tfmodel.restore('32c', module, key='codec'), Dummy)
    with pytest.raises(ValueError, match='key not found in presets'):
        tfmodel.restore('64c', module)


def test_header_is_plain_c():
    """The boundary is a C ABI: include/ni_b200.h must compile on its own as C and as C++ (no torch / CUDA types in the signatures)."""
    header = os.path.join(ROOT, 'include', 'ni_b200.h')
    for lang, cc in (('c', 'gcc'), ('c++', 'g++')):
        out = subprocess.run([cc, '-fsyntax-only', '-Wall', '-Werror', '-x', lang, header], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
    src = open(header).read()
    assert 'torch' not in src and 'at::Tensor' not in src and '#include <cuda' not in src


def test_no_undefined_names_in_bench_and_package():
    """Static check (symtable): a function of bench.py / the package that reads a name which is neither local, nor module-level, nor a
    builtin fails only when its configuration runs — on the GPU box, where a NameError costs a bench line (it did once: `clk` in run_c2)."""
    import builtins
    import symtable
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    paths = [os.path.join(root, 'bench.py'), os.path.join(root, '__graft_entry__.py')]
    for base, _, files in os.walk(os.path.join(root, 'neural_imaging_b200')):
        paths += [os.path.join(base, f) for f in files if f.endswith('.py')]
    bad = []
    for path in paths:
        with open(path) as fh:
            top = symtable.symtable(fh.read(), path, 'exec')
        module_names = {s.get_name() for s in top.get_symbols()}

        def walk(table):
            for ch in table.get_children():
                if ch.get_type() == 'function':
                    for s in ch.get_symbols():
                        if s.is_global() and not s.is_declared_global() and s.get_name() not in module_names and not hasattr(builtins, s.get_name()):
                            bad.append((os.path.relpath(path, root), ch.get_name(), s.get_name()))
                walk(ch)
        walk(top)
    assert not bad, bad
