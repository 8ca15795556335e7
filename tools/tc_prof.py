"""Per-role clock64 spans of the tcgen05 kernels (library built with NI_NVCC_EXTRA=-DNI_TC_PROFILE), CTA 0 only.

Slots 0-31: persistent gemm (fprop / dgrad); slots 32-63: wgrad. One lane per role records; numbers are cycles of that lane.
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neural_imaging_b200 import _lib, nn
from neural_imaging_b200.tensor import empty, ptr, stream

L = _lib.lib()
rs = np.random.RandomState(0)
GEMM = {1: 'mma: wait B full', 2: 'mma: wait A slot ready', 3: 'mma: wait acc free', 4: 'mma: commits', 17: 'mma: 12 MMA issue only',
        6: 'conv: wait A halo', 7: 'conv: LDS + split', 9: 'conv: wait::st + arrive (deferred)', 8: 'conv: wait slot free',
        10: 'epi: wait acc full', 14: 'epi: tmem loads', 15: 'epi: bias + act', 16: 'epi: stores', 11: 'epi: rest',
        12: 'Aprod: wait stage free', 13: 'Bprod: wait stage free'}
WGRAD = {32: 'prod: wait stage free', 33: 'prod: issue TMA boxes', 34: 'mma: wait A ready', 35: 'mma: wait B ready', 36: 'mma: issue + commit',
         37: 'Aconv: wait tile landed', 38: 'Aconv: 32 LDS + split', 39: 'Aconv: STTM + wait::st + arrive', 40: 'Btr: wait tile landed',
         41: 'Btr: transpose + fence + arrive'}
shapes = [(256, 32, 32, 128, 128, 3), (256, 128, 128, 32, 32, 3), (1280, 64, 64, 32, 64, 5), (256, 16, 16, 256, 256, 3), (256, 128, 128, 64, 32, 3)]
which = [int(a) for a in sys.argv[1:]] or list(range(len(shapes)))
buf = (ctypes.c_longlong * 64)()


def run(fn):
    for _ in range(2):
        fn()
    L.ni_tc_prof_read(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); e1.synchronize()
    L.ni_tc_prof_read(buf, 1)
    return e0.elapsed_time(e1)


def pick(nn_):
    return 128 if nn_ % 128 == 0 else (64 if nn_ % 64 == 0 else 32)


for i in which:
    n, h, w, cin, cout, k = shapes[i]
    st = nn.ParamStore()
    conv = nn.Conv2D(st, 'c', k, cin, cout, activation='leaky_relu', rng=rs)
    st.finalize()
    d = conv.desc(n, h, w)
    x = torch.randn((n, h, w, cin), device='cuda')
    dy = torch.randn((n, h, w, cout), device='cuda')
    y, dx, dw = empty((n, h, w, cout)), empty((n, h, w, cin)), empty((k, k, cin, cout))
    for name, fn, N, K in (('fprop', lambda: L.ni_conv2d_fprop_tc(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y), stream()), cout, cin),
                           ('dgrad', lambda: L.ni_conv2d_dgrad_tc(ctypes.byref(d), ptr(dy), ptr(conv.w.value), ptr(dx), stream()), cin, cout)):
        ms = run(fn)
        tiles = n * h * w // 128 * (N // pick(N))
        per_cta = -(-tiles // 148)
        iters = k * k * K // 32
        print('%s n%d %dx%d c%d->%d k%d (N tile %d): %.3f ms, %d tiles/CTA x %d iters, %.0f clk/iter at 1.965 GHz' % (
            name, n, h, w, cin, cout, k, pick(N), ms, per_cta, iters, ms * 1.965e6 / (per_cta * iters)))
        for j in GEMM:
            print('   %-36s %12d clk total  %8.1f clk / iteration' % (GEMM[j], buf[j], buf[j] / (per_cta * iters)))
    ms = run(lambda: L.ni_conv2d_wgrad_tc(ctypes.byref(d), ptr(x), ptr(dy), ptr(dw), stream()))
    tot = sum(buf[j] for j in (37, 38, 39))
    steps = max(1, round(tot / max(1.0, ms * 1.965e6))) if tot else 1
    print('wgrad n%d %dx%d c%d->%d k%d (N tile %d): %.3f ms = %.0f clk; spans below are totals of CTA (0,0,0)' % (n, h, w, cin, cout, k, pick(cout), ms, ms * 1.965e6))
    for j in WGRAD:
        print('   %-36s %12d clk total' % (WGRAD[j], buf[j]))
