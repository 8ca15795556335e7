"""Stand-alone layer mirrors (neural_imaging_b200/models/layers.py) against the oracle restatement of reference models/layers.py."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import ref_models as M
from oracle import ref_ops as R


@pytest.mark.gpu
@pytest.mark.parametrize('rounding', ['round', 'sin', 'soft', 'harmonic', 'identity'])
def test_quantization_scalar_modes(rounding):
    from neural_imaging_b200.models.layers import Quantization
    rs = np.random.RandomState(0)
    x = (rs.normal(size=(3, 16, 16, 8)) * 20).astype(np.float32)
    got = Quantization(rounding).call(x).numpy()
    ref = R.quantization(torch.tensor(x, dtype=torch.float64), rounding).numpy()
    if rounding in ('round', 'soft', 'identity'):
        ties = np.abs(np.abs(x - np.floor(x)) - 0.5) < 1e-6
        assert np.array_equal(got[~ties], ref[~ties].astype(np.float32))
    else:
        # float32 sine after an exact period-1 reduction vs the float64 oracle: 1e-6 for the sine term, the rounding of the float32
        # RESULT itself (half an ulp of |ref|: 3.8e-6 beyond +-64), and the phase the oracle inherits from TF's float32 constant 2*pi
        # (1.75e-7 * |x| radians -> up to 6e-8 * |x| in the output)
        tol = 1e-6 + 6e-8 * np.abs(x) + 0.5 * np.spacing(np.abs(ref).astype(np.float32))
        assert np.all(np.abs(got - ref) <= tol)


@pytest.mark.gpu
def test_quantization_soft_codebook_and_discrete_latent():
    from neural_imaging_b200.models.layers import DiscreteLatent, Quantization
    rs = np.random.RandomState(1)
    z = (rs.normal(size=(2, 8, 8, 16)) * 4).astype(np.float32)
    q = Quantization('soft-codebook', latent_bpf=5)
    assert q.codebook.shape == (1, 32) and q.codebook[0, 0] == -15 and q.codebook[0, -1] == 16
    cb = torch.tensor(q.codebook.reshape(-1))
    got = q(z).numpy()
    ref = R.soft_codebook_quantization(torch.tensor(z), cb).numpy()
    assert np.max(np.abs(got - ref)) < 1e-6
    dl = DiscreteLatent('soft-codebook', latent_bpf=5)
    lat, ent = dl(z)
    rq, rent = R.discrete_latent(torch.tensor(z), torch.tensor(1.0), cb)
    assert np.max(np.abs(lat.numpy() - rq.numpy())) < 1e-6
    assert abs(float(ent.numpy()) - float(rent)) < 1e-5 * max(1.0, abs(float(rent)))
    with pytest.raises(ValueError):
        Quantization('no-such-mode')
    with pytest.raises(ValueError):
        DiscreteLatent('no-such-mode')
    # the scalar rounding modes of the latent quantiser (models/layers.py:118-136) with the same entropy estimate
    for mode in ('sin', 'soft', 'identity'):
        layer = DiscreteLatent(mode, latent_bpf=5)
        lat, ent = layer(z)
        rq, rent = R.discrete_latent(torch.tensor(z), torch.ones(()), M.dcn_codebook(5), rounding=mode)
        assert np.mean(np.abs(lat.numpy() - rq.numpy()) > 1e-5) < 1e-3, mode          # 'soft': a value within float32 noise of k + 1/2 may round the other way
        assert abs(float(ent.numpy()) - float(rent)) < 1e-4 * max(1.0, abs(float(rent))), mode


@pytest.mark.gpu
def test_constrained_conv2d():
    from neural_imaging_b200.models.layers import ConstrainedConv2D
    rs = np.random.RandomState(2)
    x = rs.uniform(size=(2, 24, 40, 3)).astype(np.float32)
    layer = ConstrainedConv2D()
    k = layer.kernel.numpy()
    assert k.shape == (5, 5, 3, 3) and k[2, 2, 0, 0] == 12 and k[2, 2, 0, 1] == 0
    nf = M.constrained_filter(torch.tensor(k, dtype=torch.float64))
    ref = R.conv2d(R.tf_pad(torch.tensor(x, dtype=torch.float64), 2, 'SYMMETRIC'), nf, None, 1, 'VALID').numpy()
    got = layer(x).numpy()
    assert got.shape == x.shape
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-5
    assert np.max(np.abs(layer.normalized_kernel().numpy() - nf.numpy())) < 1e-4
    with pytest.raises(ValueError):
        layer(np.zeros((1, 8, 8, 4), np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize('n,h,w', [(2, 32, 32), (1, 128, 128), (3, 40, 72), (2, 12, 20)])
def test_constrained_conv_kernels_forward_data_and_filter_gradient(n, h, w):
    """ni_cconv5_fwd / _bwd_data / _bwd_filter (SYMMETRIC pad 2 + VALID 5x5 conv 3 -> 3 with the pad's transpose folded into the input
    gradient) against autograd on the float64 oracle (models/layers.py:45-57); sizes with partial tiles and borders inside one tile."""
    from neural_imaging_b200 import _lib
    from neural_imaging_b200.tensor import as_device, empty, ptr, stream
    rs = np.random.RandomState(n * 100 + h)
    x = rs.uniform(size=(n, h, w, 3)).astype(np.float32)
    f = rs.normal(size=(5, 5, 3, 3)).astype(np.float32)
    dy = rs.normal(size=(n, h, w, 3)).astype(np.float32)
    L = _lib.lib()
    xd, fd, dyd = as_device(x), as_device(f), as_device(dy)
    y, dx, df = empty((n, h, w, 3)), empty((n, h, w, 3)), empty((5, 5, 3, 3))
    L.ni_cconv5_fwd(ptr(xd), ptr(fd), ptr(y), n, h, w, stream())
    L.ni_cconv5_bwd_data(ptr(dyd), ptr(fd), ptr(dx), n, h, w, 0, stream())
    L.ni_cconv5_bwd_filter(ptr(xd), ptr(dyd), ptr(df), n, h, w, stream())
    dx2 = as_device(x.copy())
    L.ni_cconv5_bwd_data(ptr(dyd), ptr(fd), ptr(dx2), n, h, w, 1, stream())          # accumulate
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    ft = torch.tensor(f, dtype=torch.float64, requires_grad=True)
    yt = R.conv2d(R.tf_pad(xt, 2, 'SYMMETRIC'), ft, padding='VALID')
    gx, gf = torch.autograd.grad(yt, [xt, ft], torch.tensor(dy, dtype=torch.float64))
    assert rel_err(y.cpu().numpy(), yt.detach().numpy()) < 1e-5
    assert rel_err(dx.cpu().numpy(), gx.numpy()) < 1e-5
    assert rel_err(dx2.cpu().numpy() - x, gx.numpy()) < 1e-5
    assert rel_err(df.cpu().numpy(), gf.numpy()) < 2e-5
