"""Per-role clock64 spans of the persistent tcgen05 gemm (library built with NI_NVCC_EXTRA=-DNI_TC_PROFILE), CTA 0 only."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neural_imaging_b200 import _lib, nn
from neural_imaging_b200.tensor import empty, ptr, stream

L = _lib.lib()
rs = np.random.RandomState(0)
NAMES = {1: 'mma: wait B full', 2: 'mma: wait A slot ready', 3: 'mma: wait acc free', 4: 'mma: issue 12 MMA + commits', 6: 'conv: wait A halo',
         7: 'conv: LDS + split', 8: 'conv: wait slot free', 9: 'conv: STTM + wait::st + arrive', 10: 'epi: wait acc full', 11: 'epi: drain + store',
         12: 'Aprod: wait stage free', 13: 'Bprod: wait stage free',
         14: 'epi: tmem loads', 15: 'epi: bias + act', 16: 'epi: stores', 17: 'mma: 12 MMA issue only'}
shapes = [(256, 32, 32, 128, 128, 3), (256, 128, 128, 32, 32, 3), (1280, 64, 64, 32, 64, 5), (256, 16, 16, 256, 256, 3)]
buf = (ctypes.c_longlong * 32)()
for n, h, w, cin, cout, k in shapes:
    st = nn.ParamStore()
    conv = nn.Conv2D(st, 'c', k, cin, cout, activation='leaky_relu', rng=rs)
    st.finalize()
    d = conv.desc(n, h, w)
    x = torch.randn((n, h, w, cin), device='cuda')
    y = empty((n, h, w, cout))
    fn = lambda: L.ni_conv2d_fprop_tc(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y), stream())
    for _ in range(2):
        fn()
    L.ni_tc_prof_read(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    L.ni_tc_prof_read(buf, 1)
    tiles = n * h * w // 128 * (cout // (128 if cout % 128 == 0 else (64 if cout % 64 == 0 else 32)))
    per_cta = -(-tiles // 148)
    iters = k * k * cin // 32
    print('fprop n%d %dx%d c%d->%d k%d: %.3f ms, %d tiles/CTA x %d iters' % (n, h, w, cin, cout, k, ms, per_cta, iters))
    for i in sorted(NAMES):
        print('   %-34s %10d clk total  %8.1f clk / iteration' % (NAMES[i], buf[i], buf[i] / (per_cta * iters)))
