"""Debug: fprop of (1,256,256,32->32,k3) through the tcgen05 path vs the SIMT path, error per 128-pixel tile."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neural_imaging_b200 import _lib, nn
from neural_imaging_b200.tensor import as_device, empty, ptr, stream
L = _lib.lib()
rs = np.random.RandomState(0)
for (n, h, w, cin, cout, k) in [(1, 256, 256, 32, 32, 3), (1, 256, 256, 64, 64, 3), (16, 64, 64, 128, 128, 3), (1, 256, 256, 32, 64, 5), (1, 256, 256, 64, 32, 3)]:
    st = nn.ParamStore()
    conv = nn.Conv2D(st, 'c', k, cin, cout, activation='relu', rng=rs)
    st.finalize()
    d = conv.desc(n, h, w)
    x = as_device(rs.normal(size=(n, h, w, cin)).astype(np.float32))
    y_si = empty((n, h, w, cout))
    L.ni_conv2d_fprop_simt(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y_si), stream())
    for rep in range(4):
        y = torch.full((n, h, w, cout), 777.0, device='cuda')
        L.ni_conv2d_fprop_tc(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y), stream())
        torch.cuda.synchronize()
        err = (y - y_si).abs().amax(dim=3)[0]                       # (h, w) of image 0
        tiles = err.reshape(h // 8, 8, w // 16, 16).amax(dim=(1, 3))  # 16x8 pixel tiles
        bad = (tiles > 1e-3).nonzero()
        print('shape', (n, h, w, cin, cout, k), 'rep', rep, 'max err %.3e' % float(err.max()), 'bad tiles', int(bad.shape[0]), 'of', tiles.numel(),
              'untouched', int((y == 777.0).sum()))
        if bad.shape[0] and n == 1:
            idx = (bad[:, 0] * (w // 16) + bad[:, 1]).tolist()
            print('   bad tile ids (first 40):', idx[:40], ' -> id % 148:', sorted(set(i % 148 for i in idx))[:40], ' id // 148:', sorted(set(i // 148 for i in idx)))
