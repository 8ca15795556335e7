"""Diagnostics for the tcgen05 conv kernels: TC vs SIMT on the device, error statistics and locations."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from neural_imaging_b200 import _lib, nn
from neural_imaging_b200.tensor import as_device, empty, ptr, stream

L = _lib.lib()
tc_sup = L._dll.ni_conv2d_tc_supported
rs = np.random.RandomState(0)


def stats(name, a, b):
    a64, b64 = a.double(), b.double()
    d = (a64 - b64).abs()
    mx = float(d.max() / b64.abs().max())
    l2 = float(torch.sqrt((d * d).sum()) / torch.sqrt((b64 * b64).sum()))
    nbad = int((d > 1e-3 * b64.abs().max()).sum())
    idx = np.unravel_index(int(d.argmax()), tuple(a.shape))
    print('  %-6s max-rel %.3e  l2-rel %.3e  |a|max %.3e |b|max %.3e  bad(>1e-3) %d/%d  argmax %s' % (
        name, mx, l2, float(a64.abs().max()), float(b64.abs().max()), nbad, a.numel(), idx))
    return d


# ---- UMMA layout self-test (K-major and MN-major)
A = torch.randint(-4, 5, (128, 32), device='cuda').float()
B = torch.randint(-4, 5, (64, 32), device='cuda').float()
for mn in (0, 1, 2, 3):
    D = torch.full((128, 64), 7.0, device='cuda')
    if mn == 3:   # does the tensor core truncate or round FP32 -> TF32 ? (A with full mantissas, mode 0)
        A = (torch.randint(-4, 5, (128, 32), device='cuda').float() * (1 + 2.0 ** -11 + 2.0 ** -12 + 2.0 ** -20))
    L.ni_tc_selftest(ptr(A), ptr(B), ptr(D), 0 if mn == 3 else mn, stream())
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    if mn == 3:
        At = torch.from_numpy((A.cpu().numpy().view('uint32') & 0xFFFFE000).view('float32')).cuda()
        print('  vs truncated-A product: max diff %.6f ; vs exact product: max diff %.6f' % (float((D - At.double() @ B.double().t()).abs().max()), float((D - ref).abs().max())))
        continue
    print('selftest mn_major=%d: max|D-ref| = %.3f, |D|max %.3f, |ref|max %.3f' % (mn, float((D - ref).abs().max()), float(D.abs().max()), float(ref.abs().max())))
    if float((D - ref).abs().max()) > 0:
        print('  D[0,:8]  ', D[0, :8].tolist()); print('  ref[0,:8]', ref[0, :8].tolist())
        print('  D[:8,0]  ', D[:8, 0].tolist()); print('  ref[:8,0]', ref[:8, 0].tolist())
        print('  (A@B.T).T?', float((D[:64, :64] - ref[:64, :64].t()).abs().max()))

shapes = [(2, 16, 16, 32, 32, 3), (2, 32, 32, 128, 128, 3), (1, 256, 256, 32, 32, 3), (2, 64, 64, 64, 64, 3), (4, 8, 8, 512, 512, 3),
          (2, 32, 32, 64, 128, 5)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (n, h, w, cin, cout, k) in shapes:
    print('shape n=%d %dx%d cin=%d cout=%d k=%d' % (n, h, w, cin, cout, k))
    st = nn.ParamStore()
    conv = nn.Conv2D(st, 'c', k, cin, cout, activation=None, rng=rs)
    st.finalize()
    d = conv.desc(n, h, w)
    print('  supported:', [tc_sup(ctypes.byref(d), op) for op in (0, 1, 2)])
    x = as_device(rs.normal(size=(n, h, w, cin)).astype(np.float32))
    dy = as_device(rs.normal(size=(n, h, w, cout)).astype(np.float32))
    xt = torch.tensor(x.cpu().numpy(), dtype=torch.float64).permute(0, 3, 1, 2)
    wt = conv.w.value.cpu().double().permute(3, 2, 0, 1)
    y64 = torch.nn.functional.conv2d(xt, wt, conv.b.value.cpu().double(), padding=k // 2).permute(0, 2, 3, 1)
    y_tc, y_si = empty((n, h, w, cout)), empty((n, h, w, cout))
    L.ni_conv2d_fprop_simt(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y_si), stream())
    for rep in range(2):
        y_tc.fill_(7.0)
        L.ni_conv2d_fprop_tc(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y_tc), stream())
        stats('fprop', y_tc.cpu(), y_si.cpu())
    stats('f-tc64', y_tc.cpu(), y64)
    stats('f-si64', y_si.cpu(), y64)
    dx_tc, dx_si = empty((n, h, w, cin)), empty((n, h, w, cin))
    L.ni_conv2d_set_force_simt(1)
    L.ni_conv2d_dgrad(ctypes.byref(d), ptr(dy), ptr(conv.w.value), ptr(dx_si), stream())
    L.ni_conv2d_set_force_simt(-1)
    for rep in range(3):
        dx_tc.fill_(7.0)
        L.ni_conv2d_dgrad_tc(ctypes.byref(d), ptr(dy), ptr(conv.w.value), ptr(dx_tc), stream())
        dd = stats('dgrad', dx_tc.cpu(), dx_si.cpu())
    if float(dd.max()) > 1e-3:
        bad = (dd > 1e-3).nonzero()
        print('   bad locations (first 8):', bad[:8].tolist(), ' y-hist:', torch.bincount(bad[:, 1], minlength=h)[:40].tolist())
    dw_tc, dw_si = empty((k, k, cin, cout)), empty((k, k, cin, cout))
    L.ni_conv2d_wgrad_simt(ctypes.byref(d), ptr(x), ptr(dy), ptr(dw_si), stream())
    L.ni_conv2d_wgrad_tc(ctypes.byref(d), ptr(x), ptr(dy), ptr(dw_tc), stream())
    stats('wgrad', dw_tc.cpu(), dw_si.cpu())
    print('   dw_tc[0,0,:2,:6]', dw_tc[0, 0, :2, :6].cpu().numpy().round(3).tolist())
    print('   dw_si[0,0,:2,:6]', dw_si[0, 0, :2, :6].cpu().numpy().round(3).tolist())
    print('   dw_tc[1,1,:2,:6]', dw_tc[1, 1, :2, :6].cpu().numpy().round(3).tolist())
    print('   dw_si[1,1,:2,:6]', dw_si[1, 1, :2, :6].cpu().numpy().round(3).tolist())
    torch.cuda.synchronize()
