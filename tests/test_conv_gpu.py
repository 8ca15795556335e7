"""GPU parity of the convolution entry points (fprop / dgrad / wgrad) against torch-CPU float64 convolutions."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import assert_parity, rel_err
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu

CASES = [
    # n, h, w, cin, cout, k, stride, padding, act
    (2, 16, 16, 4, 32, 3, 1, 'SAME', 'leaky_relu'),
    (2, 16, 16, 32, 32, 3, 1, 'SAME', 'leaky_relu'),
    (1, 12, 20, 3, 32, 5, 1, 'SAME', 'leaky_relu'),
    (2, 16, 16, 32, 64, 5, 1, 'SAME', 'relu'),
    (2, 8, 8, 64, 128, 3, 1, 'SAME', None),
    (3, 16, 16, 3, 64, 5, 2, 'SAME', 'leaky_relu'),
    (2, 16, 16, 64, 128, 5, 2, 'SAME', None),
    (2, 9, 9, 16, 20, 3, 1, 'VALID', 'tanh'),
    (4, 1, 1, 256, 5, 1, 1, 'VALID', None),
    (2, 8, 8, 256, 256, 1, 1, 'VALID', 'sigmoid'),
    (1, 16, 16, 32, 12, 3, 1, 'SAME', None),
    # shapes of the real UNet / FAN layers (tcgen05 path)
    (2, 128, 128, 32, 32, 3, 1, 'SAME', 'leaky_relu'),
    (2, 64, 64, 64, 64, 3, 1, 'SAME', 'leaky_relu'),
    (2, 32, 32, 128, 128, 3, 1, 'SAME', 'leaky_relu'),
    (2, 16, 16, 256, 256, 3, 1, 'SAME', 'leaky_relu'),
    (4, 8, 8, 512, 512, 3, 1, 'SAME', 'leaky_relu'),
    (3, 8, 8, 256, 512, 3, 1, 'SAME', None),
    (2, 64, 64, 32, 64, 5, 1, 'SAME', 'leaky_relu'),
    (2, 32, 32, 64, 128, 5, 1, 'SAME', 'leaky_relu'),
    (2, 16, 16, 128, 256, 5, 1, 'SAME', 'leaky_relu'),
    (1, 256, 256, 32, 32, 3, 1, 'SAME', 'relu'),
    # register-blocked direct kernels (csrc/conv_direct.cu): several tiles, ragged right / bottom edges
    (2, 40, 72, 3, 32, 5, 1, 'SAME', 'leaky_relu'),
    (1, 24, 40, 4, 32, 3, 1, 'SAME', 'leaky_relu'),
    (2, 20, 36, 32, 12, 3, 1, 'SAME', None),
    (2, 40, 40, 3, 3, 5, 1, 'SAME', None),
    (1, 128, 128, 3, 32, 5, 1, 'SAME', 'leaky_relu'),
    (1, 64, 128, 32, 12, 3, 1, 'SAME', None),          # 12 outputs as three groups of 4 through the persistent column-blocked kernel
    (2, 64, 64, 32, 3, 5, 1, 'SAME', None),
]


def _mk(n, h, w, cin, cout, k, seed=0):
    rs = np.random.RandomState(seed)
    x = rs.normal(size=(n, h, w, cin)).astype(np.float32)
    wgt = (rs.normal(size=(k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
    b = rs.normal(size=(cout,)).astype(np.float32)
    return x, wgt, b


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('impl', ['dispatch', 'simt'])
def test_conv_fwd_bwd(case, impl):
    from neural_imaging_b200 import _lib, nn
    from neural_imaging_b200.tensor import as_device, empty, ptr, stream
    n, h, w, cin, cout, k, stride, padding, act = case
    x, wgt, b = _mk(n, h, w, cin, cout, k)
    st = nn.ParamStore()
    conv = nn.Conv2D(st, 'c', k, cin, cout, stride=stride, padding=padding, activation=act, kernel_init=wgt, bias_init=b)
    st.finalize()
    L = _lib.lib()
    d = conv.desc(n, h, w)
    xd = as_device(x)
    y = empty((n, d.oh, d.ow, cout))
    fprop = L.ni_conv2d_fprop if impl == 'dispatch' else L.ni_conv2d_fprop_simt
    fprop(ctypes.byref(d), ptr(xd), ptr(conv.w.value), ptr(conv.b.value), ptr(y), stream())
    rs = np.random.RandomState(1)
    dy = rs.normal(size=tuple(y.shape)).astype(np.float32)
    yg = y.cpu().numpy().astype(np.float64)
    # d(act)/d(pre-activation) evaluated at the DEVICE output: the activation derivative is discontinuous at 0, and an
    # output within rounding of 0 may legitimately land on either side (seen once: 1 element in 262144 -> 3.5e-2 in dx)
    dact = {None: np.ones_like(yg), 'leaky_relu': np.where(yg > 0, 1.0, 0.2), 'relu': (yg > 0).astype(np.float64),
            'tanh': 1 - yg ** 2, 'sigmoid': yg * (1 - yg)}[act]
    res = {}
    for dt in (torch.float64, torch.float32):
        xt = torch.tensor(x, dtype=dt, requires_grad=True)
        wt = torch.tensor(wgt, dtype=dt, requires_grad=True)
        bt = torch.tensor(b, dtype=dt, requires_grad=True)
        pre = R.conv2d(xt, wt, bt, stride, padding)
        gx, gw, gb = torch.autograd.grad(pre, (xt, wt, bt), torch.tensor(dy * dact, dtype=dt))
        res[dt] = [t.detach().numpy() for t in (R.ACT[act](pre), gx, gw, gb)]
    r64, r32 = res[torch.float64], res[torch.float32]
    # tcgen05 path: FP32-accurate 3xTF32; the in-TMEM accumulation truncates, so deep contractions sit at 2-4e-6
    # (measured, tools/tc_diag.py) instead of the SIMT path's ~1e-6: same 1e-5 bar, 2e-5 for the weight gradients
    assert_parity(y.cpu().numpy(), r64[0], r32[0], tol=1e-5, what='y')
    if impl == 'simt':
        return
    dyd = as_device(dy)
    dx = empty(x.shape)
    conv.bprop(xd, y, dyd, dx, d)
    assert_parity(dx.cpu().numpy(), r64[1], r32[1], tol=1e-5, what='dx')
    assert_parity(conv.w.grad.cpu().numpy(), r64[2], r32[2], tol=2e-5, what='dw')
    assert_parity(conv.b.grad.cpu().numpy(), r64[3], r32[3], tol=2e-5, what='db')


def test_pitch_offset_and_block2_addressing():
    """Concat-free skip connections and Conv2DTranspose-as-1x1-conv + depth_to_space."""
    from neural_imaging_b200 import _lib, nn
    from neural_imaging_b200._lib import MODE_BLOCK2
    from neural_imaging_b200.tensor import as_device, empty, zeros
    n, h, w, cin, c = 2, 8, 8, 64, 32
    rs = np.random.RandomState(2)
    x = rs.normal(size=(n, h, w, cin)).astype(np.float32)
    skip = rs.normal(size=(n, 2 * h, 2 * w, c)).astype(np.float32)
    st = nn.ParamStore()
    up = nn.Conv2D(st, 'up', 1, cin, 4 * c, padding='VALID', bias_mod=c,
                   kernel_init=(rs.normal(size=(1, 1, cin, 4 * c)) / 8).astype(np.float32), bias_init=rs.normal(size=(c,)).astype(np.float32))
    conv = nn.Conv2D(st, 'dc', 3, 2 * c, c, activation='leaky_relu', rng=rs)
    st.finalize()
    xd = as_device(x)
    cat = zeros((n, 2 * h, 2 * w, 2 * c))
    cat[..., c:] = as_device(skip)
    du = up.desc(n, h, w, out_pitch=2 * c, out_coff=0, out_mode=MODE_BLOCK2)
    up.fprop(xd, cat, du)
    dc = conv.desc(n, 2 * h, 2 * w)
    y = conv.fprop(cat, empty((n, 2 * h, 2 * w, c)), dc)
    # oracle
    P = {k_: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k_, v in st.state_dict().items()}
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    upt = R.conv2d_transpose_2x2(xt, P['up/kernel'], P['up/bias'])
    catt = torch.cat((upt, torch.tensor(skip, dtype=torch.float64)), dim=3)
    yt = R.leaky_relu(R.conv2d(catt, P['dc/kernel'], P['dc/bias']))
    assert_parity(cat.cpu().numpy(), catt.detach().numpy(), tol=1e-5, what='parity')
    assert_parity(y.cpu().numpy(), yt.detach().numpy(), tol=1e-5, what='parity')
    dy = rs.normal(size=tuple(y.shape)).astype(np.float32)
    grads = torch.autograd.grad(yt, [xt] + list(P.values()), torch.tensor(dy, dtype=torch.float64))
    gref = dict(zip(['x'] + list(P.keys()), [g.numpy() for g in grads]))
    dcat = empty(tuple(cat.shape))
    conv.bprop(cat, y, as_device(dy), dcat, dc)
    dx = empty(x.shape)
    up.bprop(xd, None, dcat, dx, du, dy_addr=(2 * c, 0, MODE_BLOCK2))
    assert_parity(dx.cpu().numpy(), gref['x'], tol=1e-5, what='parity')
    for p in st.params:
        assert_parity(p.grad.cpu().numpy(), gref[p.name], tol=2e-5, what='parity')


@pytest.mark.parametrize('act', ['leaky_relu', None])
def test_subpixel_conv3_with_depth_to_space_backward(act):
    """3x3 convolution + depth_to_space(2) epilogue (the DCN decoder's up-sampling layers, models/compression.py:250-262): the backward
    converts the depth_to_space-addressed gradient to the logical layout once and runs wgrad / dgrad on the dense tcgen05 path."""
    from neural_imaging_b200 import nn
    from neural_imaging_b200._lib import MODE_BLOCK2
    from neural_imaging_b200.tensor import as_device, empty
    n, h, w, cin, f = 3, 16, 16, 32, 32
    rs = np.random.RandomState(11)
    x = rs.normal(size=(n, h, w, cin)).astype(np.float32)
    st = nn.ParamStore()
    conv = nn.Conv2D(st, 'up', 3, cin, 4 * f, activation=act, rng=rs, bias_init=rs.normal(size=(4 * f,)).astype(np.float32))
    st.finalize()
    d = conv.desc(n, h, w, out_mode=MODE_BLOCK2)
    xd = as_device(x)
    y = conv.fprop(xd, empty((n, 2 * h, 2 * w, f)), d)
    dy = rs.normal(size=(n, 2 * h, 2 * w, f)).astype(np.float32)

    def d2s(t):          # TensorFlow block-major depth_to_space(2) on NHWC
        return t.reshape(n, h, w, 2, 2, f).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * h, 2 * w, f)
    res = {}
    for dt in (torch.float64, torch.float32):
        P = {k_: torch.tensor(v, dtype=dt, requires_grad=True) for k_, v in st.state_dict().items()}
        xt = torch.tensor(x, dtype=dt, requires_grad=True)
        yt = d2s(R.ACT[act](R.conv2d(xt, P['up/kernel'], P['up/bias'])))
        g = torch.autograd.grad(yt, [xt, P['up/kernel'], P['up/bias']], torch.tensor(dy, dtype=dt))
        res[dt] = [yt.detach().numpy()] + [t.numpy() for t in g]
    r64, r32 = res[torch.float64], res[torch.float32]
    assert_parity(y.cpu().numpy(), r64[0], r32[0], tol=1e-5, what='y')
    dx = empty(x.shape)
    conv.bprop(xd, y, as_device(dy), dx, d)
    assert_parity(dx.cpu().numpy(), r64[1], r32[1], tol=1e-5, slack=6.0, what='dx')
    assert_parity(conv.w.grad.cpu().numpy(), r64[2], r32[2], tol=2e-5, slack=6.0, what='dw')
    assert_parity(conv.b.grad.cpu().numpy(), r64[3], r32[3], tol=2e-5, slack=6.0, what='db')


def test_space_to_depth2_round_trip_any_channel_count():
    from neural_imaging_b200 import _lib
    from neural_imaging_b200.tensor import as_device, empty, ptr, stream
    L = _lib.lib()
    rs = np.random.RandomState(3)
    for c in (3, 8):
        n, h2, w2 = 2, 5, 7
        x = rs.normal(size=(n, 2 * h2, 2 * w2, c)).astype(np.float32)
        ref = x.reshape(n, h2, 2, w2, 2, c).transpose(0, 1, 3, 2, 4, 5).reshape(n, h2, w2, 4 * c)      # tf.nn.space_to_depth(x, 2)
        xd, deep = as_device(x), empty((n, h2, w2, 4 * c))
        L.ni_space_to_depth2(ptr(xd), ptr(deep), n, h2, w2, c, 0, 0, stream())
        assert np.array_equal(deep.cpu().numpy(), ref)
        back = as_device(np.ones_like(x))
        L.ni_space_to_depth2(ptr(back), ptr(deep), n, h2, w2, c, 1, 1, stream())
        assert np.array_equal(back.cpu().numpy(), x + 1.0)


def test_mirrored_pad_conv_and_fold():
    """ConstrainedConv2D-style conv: SYMMETRIC pad folded into the addressing; dgrad via padded-domain + ni_pad_fold."""
    from neural_imaging_b200 import _lib, nn
    from neural_imaging_b200._lib import PAD_REFLECT, PAD_SYMMETRIC, PAD_ZERO
    from neural_imaging_b200.tensor import as_device, empty, ptr, stream
    L = _lib.lib()
    for mode, name in ((PAD_SYMMETRIC, 'SYMMETRIC'), (PAD_REFLECT, 'REFLECT')):
        n, h, w, c, k = 2, 12, 10, 3, 5
        rs = np.random.RandomState(3)
        x = rs.normal(size=(n, h, w, c)).astype(np.float32)
        wgt = rs.normal(size=(k, k, c, c)).astype(np.float32)
        st = nn.ParamStore()
        conv = nn.Conv2D(st, 'c', k, c, c, padding='VALID', use_bias=False, pad_mode=mode, explicit_pad=k // 2, kernel_init=wgt)
        st.finalize()
        d = conv.desc(n, h, w)
        xd = as_device(x)
        y = conv.fprop(xd, empty((n, h, w, c)), d)
        xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        wt = torch.tensor(wgt, dtype=torch.float64, requires_grad=True)
        yt = R.conv2d(R.tf_pad(xt, k // 2, name), wt, padding='VALID')
        assert_parity(y.cpu().numpy(), yt.detach().numpy(), tol=1e-5, what='parity')
        dy = rs.normal(size=(n, h, w, c)).astype(np.float32)
        gx, gw = torch.autograd.grad(yt, (xt, wt), torch.tensor(dy, dtype=torch.float64))
        dyd = as_device(dy)
        L.ni_conv2d_wgrad(ctypes.byref(d), ptr(xd), ptr(dyd), ptr(conv.w.grad), stream())
        assert_parity(conv.w.grad.cpu().numpy(), gw.numpy(), tol=1e-5, what='parity')
        p = k // 2
        dd = conv.desc(n, h + 2 * p, w + 2 * p)
        dd.pad_t = dd.pad_l = 0
        dd.oh, dd.ow, dd.pad_mode = h, w, PAD_ZERO
        dpad = empty((n, h + 2 * p, w + 2 * p, c))
        L.ni_conv2d_dgrad(ctypes.byref(dd), ptr(dyd), ptr(conv.w.value), ptr(dpad), stream())
        dx = empty((n, h, w, c))
        L.ni_pad_fold(ptr(dpad), ptr(dx), n, h, w, c, p, mode, 0, stream())
        assert_parity(dx.cpu().numpy(), gx.numpy(), tol=1e-5, what='parity')


def test_maxpool_gap_softmax_adam():
    from neural_imaging_b200 import _lib
    from neural_imaging_b200.tensor import as_device, empty, ptr, stream, zeros
    L = _lib.lib()
    rs = np.random.RandomState(4)
    for same, (h, w) in ((1, (8, 8)), (0, (8, 8)), (1, (7, 9)), (0, (7, 9))):
        n, c = 2, 6
        x = rs.normal(size=(n, h, w, c)).astype(np.float32)
        oh, ow = ((h + 1) // 2, (w + 1) // 2) if same else (h // 2, w // 2)
        xd, y = as_device(x), empty((n, oh, ow, c))
        L.ni_maxpool2_fwd(ptr(xd), ptr(y), n, h, w, c, same, c, 0, c, 0, stream())
        xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        yt = R.max_pool(xt, same)
        assert np.array_equal(y.cpu().numpy(), yt.detach().numpy().astype(np.float32))
        dy = rs.normal(size=(n, oh, ow, c)).astype(np.float32)
        add = rs.normal(size=(n, h, w, c)).astype(np.float32)
        g, = torch.autograd.grad(yt, xt, torch.tensor(dy, dtype=torch.float64))
        dx, dyd, addd = empty((n, h, w, c)), as_device(dy), as_device(add)
        L.ni_maxpool2_bwd(ptr(xd), ptr(dyd), ptr(addd), ptr(dx), n, h, w, c, same, c, 0, c, 0, c, 0, c, 0, stream())
        assert_parity(dx.cpu().numpy(), g.numpy() + add, tol=1e-6, what='maxpool bwd')
    # softmax + Keras sparse CE (+ gradient)
    m, c = 37, 5
    z = (rs.normal(size=(m, c)) * 4).astype(np.float32)
    z[0] = [40, 0, 0, 0, -40]          # saturated probabilities exercise the 1e-7 clip
    lab = rs.randint(0, c, size=(m,)).astype(np.int32)
    zd, probs, loss, dz, labd = as_device(z), empty((m, c)), zeros((1,)), empty((m, c)), as_device(lab, torch.int32)
    L.ni_softmax_ce(ptr(zd), ptr(labd), ptr(probs), ptr(loss), ptr(dz), m, c, 1.0 / m, stream())
    zt = torch.tensor(z, dtype=torch.float64, requires_grad=True)
    pt = torch.softmax(zt, dim=1)
    lt = R.sparse_categorical_crossentropy(lab, pt)
    g, = torch.autograd.grad(lt, zt)
    assert_parity(probs.cpu().numpy(), pt.detach().numpy(), tol=1e-6, what='parity')
    assert abs(float(loss.item()) / m - float(lt)) < 1e-5 * max(1.0, abs(float(lt)))
    assert_parity(dz.cpu().numpy(), g.numpy(), tol=1e-4, what='parity')
    # Keras Adam, 3 steps
    nparam = 1003
    p0 = rs.normal(size=(nparam,)).astype(np.float32)
    pd, md, vd, flag = as_device(p0.copy()), zeros((nparam,)), zeros((nparam,)), zeros((1,), torch.int32)
    pt_, mt, vt = [torch.tensor(p0, dtype=torch.float64)], [torch.zeros(nparam, dtype=torch.float64)], [torch.zeros(nparam, dtype=torch.float64)]
    for t in range(1, 4):
        g = rs.normal(size=(nparam,)).astype(np.float32)
        gd = as_device(g)
        L.ni_adam_keras(ptr(pd), ptr(gd), ptr(md), ptr(vd), nparam, 1e-3, 0.9, 0.999, 1e-7, t, 1.0, ptr(flag), stream())
        R.adam_keras_step(pt_, [torch.tensor(g, dtype=torch.float64)], mt, vt, t, 1e-3)
    assert rel_err(pd.cpu().numpy(), pt_[0].numpy()) < 1e-6 and int(flag.item()) == 0
    gbad = np.zeros((nparam,), np.float32); gbad[5] = np.nan
    gd = as_device(gbad)
    L.ni_adam_keras(ptr(pd), ptr(gd), ptr(md), ptr(vd), nparam, 1e-3, 0.9, 0.999, 1e-7, 4, 1.0, ptr(flag), stream())
    assert int(flag.item()) == 1


def test_tc_path_is_taken_and_matches_simt():
    """The dispatcher must route dense layers to the tcgen05 kernels (not silently to SIMT), and both must agree."""
    from neural_imaging_b200 import _lib, nn
    from neural_imaging_b200.tensor import as_device, empty, ptr, stream
    L = _lib.lib()
    rs = np.random.RandomState(5)
    for (n, h, w, cin, cout, k) in [(8, 64, 64, 64, 64, 3), (4, 32, 32, 128, 128, 5), (16, 8, 8, 512, 512, 3), (8, 128, 128, 32, 32, 3)]:
        st = nn.ParamStore()
        conv = nn.Conv2D(st, 'c', k, cin, cout, activation='leaky_relu', rng=rs)
        st.finalize()
        d = conv.desc(n, h, w)
        for op in (0, 1, 2):
            assert L.ni_conv2d_tc_supported(ctypes.byref(d), op) == 1
        x = as_device(rs.normal(size=(n, h, w, cin)).astype(np.float32))
        dy = as_device(rs.normal(size=(n, h, w, cout)).astype(np.float32))
        y_tc, y_si = empty((n, h, w, cout)), empty((n, h, w, cout))
        L.ni_conv2d_fprop_tc(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y_tc), stream())
        L.ni_conv2d_fprop_simt(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y_si), stream())
        assert_parity(y_tc.cpu().numpy(), y_si.cpu().numpy(), tol=1e-5, what='fprop tc vs simt')
        dx_tc, dx_si = empty((n, h, w, cin)), empty((n, h, w, cin))
        L.ni_conv2d_dgrad_tc(ctypes.byref(d), ptr(dy), ptr(conv.w.value), ptr(dx_tc), stream())
        L.ni_conv2d_set_force_simt(1)
        L.ni_conv2d_dgrad(ctypes.byref(d), ptr(dy), ptr(conv.w.value), ptr(dx_si), stream())
        L.ni_conv2d_set_force_simt(-1)
        assert_parity(dx_tc.cpu().numpy(), dx_si.cpu().numpy(), tol=1e-5, what='dgrad tc vs simt')
        dw_tc, dw_si = empty((k, k, cin, cout)), empty((k, k, cin, cout))
        L.ni_conv2d_wgrad_tc(ctypes.byref(d), ptr(x), ptr(dy), ptr(dw_tc), stream())
        L.ni_conv2d_wgrad_simt(ctypes.byref(d), ptr(x), ptr(dy), ptr(dw_si), stream())
        assert_parity(dw_tc.cpu().numpy(), dw_si.cpu().numpy(), tol=2e-5, what='wgrad tc vs simt')


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(1, 256, 256, 32, 32, 3), (1, 256, 256, 64, 64, 3), (1, 256, 256, 32, 64, 5), (1, 256, 256, 64, 32, 3),
                                   (16, 64, 64, 128, 128, 3), (8, 128, 128, 32, 32, 3), (20, 64, 64, 32, 64, 5)])
def test_persistent_gemm_many_tiles_per_cta(shape):
    """More pixel tiles than SMs: every persistent CTA walks several tiles (stage / slot / accumulator-set recycling, resident and
    streamed weights). Cold launches into fresh output buffers, repeated: a stage handed back to the TMA producer too early showed
    up as corrupted rows in the SECOND tile of a CTA on the first launches of a process only."""
    from neural_imaging_b200 import _lib, nn
    from neural_imaging_b200.tensor import as_device, ptr, stream
    L = _lib.lib()
    n, h, w, cin, cout, k = shape
    rs = np.random.RandomState(11)
    st = nn.ParamStore()
    conv = nn.Conv2D(st, 'c', k, cin, cout, activation='relu', rng=rs)
    st.finalize()
    d = conv.desc(n, h, w)
    x = as_device(rs.normal(size=(n, h, w, cin)).astype(np.float32))
    dy = as_device(rs.normal(size=(n, h, w, cout)).astype(np.float32))
    y_si = torch.empty((n, h, w, cout), device='cuda')
    dx_si = torch.empty((n, h, w, cin), device='cuda')
    L.ni_conv2d_fprop_simt(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y_si), stream())
    L.ni_conv2d_set_force_simt(1)
    L.ni_conv2d_dgrad(ctypes.byref(d), ptr(dy), ptr(conv.w.value), ptr(dx_si), stream())
    L.ni_conv2d_set_force_simt(-1)
    scale_y, scale_dx = float(y_si.abs().max()), float(dx_si.abs().max())
    keep = []
    for rep in range(3):
        y = torch.full((n, h, w, cout), 777.0, device='cuda')          # fresh buffers on purpose
        dx = torch.full((n, h, w, cin), 777.0, device='cuda')
        keep.append((y, dx))
        L.ni_conv2d_fprop_tc(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y), stream())
        L.ni_conv2d_dgrad_tc(ctypes.byref(d), ptr(dy), ptr(conv.w.value), ptr(dx), stream())
        ey = float((y - y_si).abs().max()) / scale_y
        edx = float((dx - dx_si).abs().max()) / scale_dx
        assert ey <= 1e-5, 'fprop rep {}: relative error {:.3e}'.format(rep, ey)
        assert edx <= 1e-5, 'dgrad rep {}: relative error {:.3e}'.format(rep, edx)
