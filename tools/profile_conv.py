"""Run single conv layers of the real step through the C-ABI (for ncu captures and CUDA-event timing)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neural_imaging_b200 import _lib, nn
from neural_imaging_b200.tensor import as_device, empty, ptr, stream

L = _lib.lib()
rs = np.random.RandomState(0)
shapes = [(256, 32, 32, 128, 128, 3), (256, 128, 128, 32, 32, 3), (1280, 64, 64, 32, 64, 5), (256, 16, 16, 256, 256, 3),
          (1280, 128, 128, 3, 32, 5), (256, 128, 128, 32, 12, 3),      # 4, 5: direct FP32 kernels (through the dispatcher)
          (256, 64, 64, 64, 64, 3), (256, 64, 64, 128, 64, 3)]          # 6, 7: N = 64 tiles in both directions / fprop
which = [int(a) for a in sys.argv[1:]] or list(range(len(shapes)))
res = {}
for i in which:
    n, h, w, cin, cout, k = shapes[i]
    st = nn.ParamStore()
    conv = nn.Conv2D(st, 'c', k, cin, cout, activation='leaky_relu', rng=rs)
    st.finalize()
    d = conv.desc(n, h, w)
    x = torch.randn((n, h, w, cin), device='cuda')
    dy = torch.randn((n, h, w, cout), device='cuda')
    y, dx, dw = empty((n, h, w, cout)), empty((n, h, w, cin)), empty((k, k, cin, cout))
    flop = 2.0 * n * h * w * cin * cout * k * k
    tc = cin % 32 == 0 and cout % 32 == 0
    fprop, dgrad, wgrad = (L.ni_conv2d_fprop_tc, L.ni_conv2d_dgrad_tc, L.ni_conv2d_wgrad_tc) if tc else (L.ni_conv2d_fprop, L.ni_conv2d_dgrad, L.ni_conv2d_wgrad)
    for name, fn in (('fprop', lambda: fprop(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y), stream())),
                     ('dgrad', lambda: dgrad(ctypes.byref(d), ptr(dy), ptr(conv.w.value), ptr(dx), stream())),
                     ('wgrad', lambda: wgrad(ctypes.byref(d), ptr(x), ptr(dy), ptr(dw), stream()))):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res['%s n%d %dx%d c%d->%d k%d' % (name, n, h, w, cin, cout, k)] = {'ms': ms, 'tflops': flop / (ms * 1e-3) / 1e12}
print(json.dumps(res, indent=1))
