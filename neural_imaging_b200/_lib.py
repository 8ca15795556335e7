"""ctypes binding of libni_b200.so — generated from include/ni_b200.h so that header and binding cannot drift.

The product path has NO fallback: if the shared library is missing or a symbol declared in the header is not
exported, importing this module (or calling the missing symbol) raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), 'include', 'ni_b200.h')
# NI_B200_LIB selects a development variant of the same library (tools/: profiling build); never a different implementation
LIB_PATH = os.environ.get('NI_B200_LIB') or os.path.join(_HERE, 'libni_b200.so')


class ConvDesc(ctypes.Structure):
    """Mirror of ni_conv_desc (include/ni_b200.h)."""
    _fields_ = [(k, ctypes.c_int) for k in
                ('n', 'h', 'w', 'cin', 'cout', 'kh', 'kw', 'stride', 'pad_t', 'pad_l', 'oh', 'ow', 'in_pitch', 'in_coff',
                 'out_pitch', 'out_coff', 'in_mode', 'out_mode', 'act')] + \
               [('act_alpha', ctypes.c_float), ('accumulate', ctypes.c_int), ('bias_mod', ctypes.c_int), ('pad_mode', ctypes.c_int)]


ACT_NONE, ACT_LEAKY_RELU, ACT_RELU, ACT_TANH, ACT_SIGMOID, ACT_CLIP01 = range(6)
MODE_PLAIN, MODE_BLOCK2 = 0, 1
PAD_ZERO, PAD_SYMMETRIC, PAD_REFLECT = 0, 1, 2

_CTYPES = {
    'int': ctypes.c_int, 'float': ctypes.c_float, 'double': ctypes.c_double, 'long long': ctypes.c_longlong,
    'unsigned long long': ctypes.c_ulonglong, 'long long*': ctypes.c_void_p, 'ni_stream_t': ctypes.c_void_p, 'void': None,
}


def _ctype(decl):
    decl = decl.strip()
    if '*' in decl:
        return ctypes.c_char_p if decl.replace(' ', '') == 'constchar*' else ctypes.c_void_p
    decl = re.sub(r'\bconst\b', '', decl).strip()
    return _CTYPES[decl]


def parse_header(path=HEADER):
    """Return {name: (restype, [argtypes], [argnames])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    src = re.sub(r'^[ \t]*#.*$', ' ', src, flags=re.M)            # preprocessor lines
    src = re.sub(r'typedef struct ni_conv_desc \{.*?\} ni_conv_desc;', ' ', src, flags=re.S)
    src = re.sub(r'enum \{.*?\};', ' ', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'([A-Za-z_][\w\s\*]*?)\b(ni_\w+)\s*\(([^;{}]*?)\)\s*;', src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes, argnames = [], []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                mm = re.match(r'(.*?)(\w+)$', a)
                argtypes.append(_ctype(mm.group(1)))
                argnames.append(mm.group(2))
        protos[name] = (_ctype(ret), argtypes, argnames)
    return protos


class NIError(RuntimeError):
    pass


# Optional per-call profiler (bench.py installs one): an object with before(name, args) / after(name, args), used to
# bracket every C-ABI call with CUDA events. None on the normal path.
PROFILER = None


class _Lib:
    def __init__(self):
        if not os.path.isfile(LIB_PATH):
            raise ImportError('{} not found: build it with `python -m neural_imaging_b200.build` or '
                              '__graft_entry__.build() — there is no CPU/PyTorch fallback'.format(LIB_PATH))
        self._dll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        dev_header = os.path.join(os.path.dirname(HEADER), 'ni_b200_dev.h')
        if os.environ.get('NI_B200_LIB') and os.path.isfile(dev_header):       # development variant: probes / profiler of csrc/dev
            self.protos.update({k: v for k, v in parse_header(dev_header).items() if hasattr(self._dll, k)})
        for name, (ret, argtypes, _) in self.protos.items():
            fn = getattr(self._dll, name)   # AttributeError if the header declares a symbol the library lacks
            fn.restype = ret
            fn.argtypes = argtypes
        self._dll.ni_last_error.restype = ctypes.c_char_p

    def __getattr__(self, name):
        protos = object.__getattribute__(self, 'protos')
        if name not in protos:
            raise AttributeError(name)
        fn = getattr(self._dll, name)
        if protos[name][0] is not ctypes.c_int or name.endswith('_supported') or name in ('ni_version', 'ni_device_arch', 'ni_tma_probe'):
            return fn

        def checked(*args):
            prof = PROFILER
            if prof is not None:
                prof.before(name, args)
            rc = fn(*args)
            if rc != 0:
                raise NIError('{} failed ({}): {}'.format(name, rc, self._dll.ni_last_error().decode()))
            if prof is not None:
                prof.after(name, args)
            return rc
        checked.__name__ = name
        setattr(self, name, checked)
        return checked


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
