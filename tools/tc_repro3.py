"""Debug: which data do the corrupted tiles of the persistent gemm (N tile 32) contain?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neural_imaging_b200 import _lib, nn
from neural_imaging_b200.tensor import as_device, empty, ptr, stream
L = _lib.lib()
rs = np.random.RandomState(0)
n, h, w, cin, cout, k = 1, 256, 256, 32, 32, 3
st = nn.ParamStore()
conv = nn.Conv2D(st, 'c', k, cin, cout, activation=None, rng=rs)
st.finalize()
d = conv.desc(n, h, w)
x = as_device(rs.normal(size=(n, h, w, cin)).astype(np.float32))
y_si = empty((n, h, w, cout))
L.ni_conv2d_fprop_simt(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y_si), stream())
def tiles_of(t):
    return t[0].reshape(h // 8, 8, w // 16, 16, cout).permute(0, 2, 1, 3, 4).reshape(-1, 128, cout)     # (tile id, pixel, channel)
ref = tiles_of(y_si)
# per-tap contributions (plain torch, debug only): C[t] = conv(x, w restricted to tap t), to see WHICH taps are wrong in a corrupted row
xt = x.permute(0, 3, 1, 2)
wt = conv.w.value.reshape(k, k, cin, cout)
C = []
for t in range(k * k):
    wm = torch.zeros_like(wt); wm[t // k, t % k] = wt[t // k, t % k]
    C.append(tiles_of(torch.nn.functional.conv2d(xt, wm.permute(3, 2, 0, 1), padding=k // 2).permute(0, 2, 3, 1).contiguous()))
C = torch.stack(C)          # (taps, tiles, 128, cout)
bias = conv.b.value.reshape(1, 1, cout)
nbad_total = 0
for rep in range(6):
    y = torch.full((n, h, w, cout), 777.0, device='cuda')
    L.ni_conv2d_fprop_tc(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y), stream())
    torch.cuda.synchronize()
    got = tiles_of(y)
    err = (got - ref).abs().amax(dim=(1, 2))
    bad = (err > 1e-3).nonzero().flatten().tolist()
    nbad_total += len(bad)
    print('rep', rep, 'bad tiles', len(bad), bad[:12])
    for T in bad[:4]:
        g = got[T] - bias
        line = 'tile %d: rows wrong %d/128, cols wrong %d/32;' % (T, int(((got[T] - ref[T]).abs().amax(dim=1) > 1e-3).sum()), int(((got[T] - ref[T]).abs().amax(dim=0) > 1e-3).sum()))
        for name, U in (('own', T), ('prev(-148)', T - 148), ('next(+148)', T + 148), ('next2(+296)', T + 296)):
            if 0 <= U < ref.shape[0]:
                r = ref[U] - bias
                a = float((g * r).sum() / (r * r).sum())          # least-squares fraction of that tile's result present in the output
                line += ' %s: %.3f' % (name, a)
        print('   ', line)
        wrong_rows = ((got[T] - ref[T]).abs().amax(dim=1) > 1e-3).nonzero().flatten().tolist()
        print('    wrong rows:', wrong_rows[:40])
        for r_ in wrong_rows[:6]:
            resid = (got[T, r_] - ref[T, r_]).double()
            A = C[:, T, r_, :].double().t()            # (cout, taps)
            sol = torch.linalg.lstsq(A, resid.unsqueeze(1)).solution.flatten()
            fit = float((A @ sol - resid).norm() / resid.norm())
            print('      row %3d: residual = sum_t alpha_t * tap_t, alpha =' % r_, ['%.2f' % float(a) for a in sol], 'misfit %.2f' % fit)
print('total bad', nbad_total)
