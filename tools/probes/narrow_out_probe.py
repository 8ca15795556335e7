import ctypes, sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from neural_imaging_b200 import _lib, nn
from neural_imaging_b200._lib import MODE_BLOCK2
from neural_imaging_b200.tensor import empty, ptr, stream
L = _lib.lib(); rs = np.random.RandomState(0)
n, h, w = 256, 128, 128
st = nn.ParamStore(); conv = nn.Conv2D(st, 'c', 3, 32, 12, activation='clip01', rng=rs); st.finalize()
x = torch.randn((n, h, w, 32), device='cuda')
for mode, shape in ((0, (n, h, w, 12)), (MODE_BLOCK2, (n, 2 * h, 2 * w, 3))):
    d = conv.desc(n, h, w, out_mode=mode) if mode else conv.desc(n, h, w)
    y = empty(shape)
    for name, fn in (('dispatch', L.ni_conv2d_fprop), ('tc', L.ni_conv2d_fprop_tc), ('direct', L.ni_conv2d_fprop_direct)):
        f = lambda: fn(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y), stream())
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); e1.synchronize()
        print('mode', mode, name, round(e0.elapsed_time(e1) / 10, 4), 'ms', 'supported', L.ni_conv2d_tc_supported(ctypes.byref(d), 0))
