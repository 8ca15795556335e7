"""Repro of tests/test_conv_gpu.py::test_conv_fwd_bwd[dispatch-case13] with error localisation."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from neural_imaging_b200 import _lib, nn
from neural_imaging_b200.tensor import as_device, empty, ptr, stream

L = _lib.lib()
n, h, w, cin, cout, k = 2, 32, 32, 128, 128, 3
rs = np.random.RandomState(0)
x = rs.normal(size=(n, h, w, cin)).astype(np.float32)
wgt = (rs.normal(size=(k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
b = rs.normal(size=(cout,)).astype(np.float32)
dy0 = np.random.RandomState(1).normal(size=(n, h, w, cout)).astype(np.float32)


def locate(name, a, ref):
    d = (a.double() - ref.double()).abs()
    print('%s: max-rel %.3e, bad(>1e-3*max) %d' % (name, float(d.max() / ref.abs().max()), int((d > 1e-3 * ref.abs().max()).sum())))
    bad = (d > 1e-3 * float(ref.abs().max())).nonzero()
    if len(bad):
        print('   first bad', bad[:6].tolist())
        print('   n-hist', torch.bincount(bad[:, 0], minlength=n).tolist(), ' y-hist', torch.bincount(bad[:, 1], minlength=h).tolist())
        print('   x-hist', torch.bincount(bad[:, 2], minlength=w).tolist())
        print('   c-hist(by 32)', torch.bincount(bad[:, 3] // 32, minlength=cin // 32).tolist())
        i = tuple(bad[0].tolist())
        print('   got %.5f ref %.5f' % (float(a[i]), float(ref[i])))


for act in (None, 'leaky_relu'):
    st = nn.ParamStore()
    conv = nn.Conv2D(st, 'c', k, cin, cout, activation=act, kernel_init=wgt, bias_init=b)
    st.finalize()
    d = conv.desc(n, h, w)
    xd = as_device(x)
    y = empty((n, h, w, cout))
    L.ni_conv2d_fprop(ctypes.byref(d), ptr(xd), ptr(conv.w.value), ptr(conv.b.value), ptr(y), stream())
    for variant in ('dgrad only', 'actbwd+dgrad', 'full bprop'):
        dyd = as_device(dy0.copy())
        dx_tc, dx_si = empty(x.shape), empty(x.shape)
        if variant == 'dgrad only':
            L.ni_conv2d_dgrad_tc(ctypes.byref(d), ptr(dyd), ptr(conv.w.value), ptr(dx_tc), stream())
        elif variant == 'actbwd+dgrad':
            L.ni_act_bwd_bias(ptr(y), ptr(dyd), ptr(conv.b.grad), n, h, w, cout, cout, 0, 0, cout, 0, 0, d.act, d.act_alpha, 0, stream())
            L.ni_conv2d_dgrad_tc(ctypes.byref(d), ptr(dyd), ptr(conv.w.value), ptr(dx_tc), stream())
        else:
            conv.bprop(xd, y, dyd, dx_tc, d)
        torch.cuda.synchronize()
        L.ni_conv2d_set_force_simt(1)
        L.ni_conv2d_dgrad(ctypes.byref(d), ptr(dyd), ptr(conv.w.value), ptr(dx_si), stream())
        L.ni_conv2d_set_force_simt(-1)
        locate('act=%s %s' % (act, variant), dx_tc.cpu(), dx_si.cpu())
        print('   dy stats: max %.3f, nan %d; dx_tc nan %d' % (float(dyd.abs().max()), int(torch.isnan(dyd).sum()), int(torch.isnan(dx_tc).sum())))
