"""Training-data feed: host logic against the oracle restatement of helpers/dataset.py / helpers/loading.py,
device kernels (ni_feed_gather / ni_feed_convert) bit-exact against the reference's float64-division batches."""
import numpy as np
import pytest
import torch

from oracle import ref_data as RD


def _images(n=6, h=96, w=128, seed=3):
    rs = np.random.RandomState(seed)
    y = np.zeros((n, h, w, 3), dtype=np.uint8)
    for i in range(n):
        base = rs.randint(0, 256, size=(h // 8, w // 8, 3))
        img = np.kron(base, np.ones((8, 8, 1)))                      # blocky texture: mixed flat / textured patches
        if i % 3 == 0:
            img[:, : w // 2] = 40 + 3 * i                           # a flat half: the discard policies have something to reject
        if i % 3 == 1:
            img = img * 0.02 + 200                                   # nearly flat, bright: exercises 'dark-n-textured' acceptance
        y[i] = np.clip(img + rs.randint(0, 3, size=img.shape), 0, 255).astype(np.uint8)
    x = rs.randint(0, 65536, size=(n, h // 2, w // 2, 4)).astype(np.uint16)
    return x, y


def test_division_matches_float64_path_exhaustively():
    """float32(v) / float32(d) (what the kernels compute with IEEE division) == float32(float64(v) / d) (the reference) for
    every uint16 / uint8 value."""
    v16 = np.arange(65536, dtype=np.float64)
    assert np.array_equal((v16 / 65535).astype(np.float32), v16.astype(np.float32) / np.float32(65535))
    v8 = np.arange(256, dtype=np.float64)
    assert np.array_equal((v8 / 255).astype(np.float32), v8.astype(np.float32) / np.float32(255))


def test_patch_stats_match_numpy():
    from neural_imaging_b200.helpers.dataset import PatchStats
    _, y = _images()
    st = PatchStats(y[0])
    rs = np.random.RandomState(0)
    for _ in range(50):
        p = int(rs.choice([16, 32, 64]))
        yy, xx = int(rs.randint(0, y.shape[1] - p)), int(rs.randint(0, y.shape[2] - p))
        patch = y[0][yy:yy + p, xx:xx + p].astype(np.float64) / 255
        m, v = st.mean_var(yy, xx, p)
        assert abs(m - patch.mean()) < 1e-12 and abs(v - patch.var()) < 1e-12


@pytest.mark.parametrize('discard', [None, 'flat', 'flat-aggressive', 'dark-n-textured'])
@pytest.mark.parametrize('fast', [False, True])
def test_positions_and_host_batches_match_the_reference_restatement(discard, fast):
    from neural_imaging_b200.helpers.dataset import Dataset
    x, y = _images()
    ds = Dataset.from_arrays(x=x, y=y, fast_stats=fast)
    for batch_id in (0, 1):
        np.random.seed(77 + batch_id)
        bx, by = ds.next_training_batch(batch_id, 3, 32, discard, max_attempts=6)
        np.random.seed(77 + batch_id)
        rx, ry, pos = RD.next_training_batch(x, y, batch_id, 3, 32, discard, max_attempts=6)
        assert np.array_equal(bx, rx) and np.array_equal(by, ry)
        assert bx.dtype == np.float32 and by.dtype == np.float32
        assert (pos[:, 1:] % 2 == 0).all()                          # Bayer alignment


def test_dataset_errors_mirror_the_reference():
    from neural_imaging_b200.helpers.dataset import Dataset, sample_patch
    x, y = _images()
    with pytest.raises(ValueError):
        Dataset({'y': y}, load='z')
    ds = Dataset.from_arrays(x=x, y=y)
    with pytest.raises(ValueError):
        ds.next_training_batch(5, 3, 32, 'flat')                     # not enough images
    with pytest.raises(ValueError):
        sample_patch(y[0], 32, 'no-such-mode')
    with pytest.raises(ValueError):
        Dataset.from_arrays(x=x).next_training_batch(0, 2, 32, 'flat')   # discard needs RGB
    with pytest.raises(KeyError):
        ds['test']
    assert sample_patch(y[0][:32, :32], 32, 'flat') == (0, 0)        # no room to move: (0, 0) without touching the RNG


@pytest.mark.gpu
@pytest.mark.parametrize('discard', [None, 'flat-aggressive'])
def test_device_batches_are_bit_identical(discard):
    from neural_imaging_b200.helpers.dataset import Dataset
    x, y = _images(n=8, h=160, w=192)
    ds = Dataset.from_arrays(x=x, y=y)
    for batch_id in (0, 1):
        np.random.seed(5 + batch_id)
        dx, dy = ds.next_training_batch_device(batch_id, 4, 64, discard)
        np.random.seed(5 + batch_id)
        rx, ry, _ = RD.next_training_batch(x, y, batch_id, 4, 64, discard)
        assert np.array_equal(dx.numpy(), rx)
        assert np.array_equal(dy.numpy(), ry)


@pytest.mark.gpu
def test_gather_rejects_out_of_range_and_handles_ragged_sizes():
    from neural_imaging_b200 import _lib
    from neural_imaging_b200.tensor import ptr, stream
    L = _lib.lib()
    rs = np.random.RandomState(0)
    img = rs.randint(0, 256, size=(3, 37, 53, 3)).astype(np.uint8)           # odd sizes: unaligned row starts
    d = torch.from_numpy(img).cuda()
    coords = np.array([[0, 0, 0], [2, 37 - 9, 53 - 11], [1, 5, 7], [9, 0, 0], [1, 30, 0]], dtype=np.int32)   # [3]: bad image, [4]: y + ph > H
    out = torch.full((5, 9, 11, 3), -1.0, device='cuda')
    L.ni_feed_gather(ptr(d), 1, 3, 37, 53, 3, ptr(torch.from_numpy(coords).cuda()), 5, 9, 11, 255.0, ptr(out), stream())
    o = out.cpu().numpy()
    for b in (0, 1, 2):
        i, yy, xx = coords[b]
        assert np.array_equal(o[b], (img[i, yy:yy + 9, xx:xx + 11].astype(np.float64) / 255).astype(np.float32))
    assert (o[3] == -1).all() and (o[4] == -1).all()                          # invalid triples leave the output untouched
    with pytest.raises(_lib.NIError):
        L.ni_feed_gather(ptr(d), 4, 3, 37, 53, 3, ptr(torch.from_numpy(coords).cuda()), 5, 9, 11, 255.0, ptr(out), stream())


@pytest.mark.gpu
@pytest.mark.parametrize('n', [0, 1, 7, 16, 4099, 1 << 20])
def test_convert_matches_numpy(n):
    from neural_imaging_b200 import _lib
    from neural_imaging_b200.tensor import ptr, stream
    L = _lib.lib()
    rs = np.random.RandomState(n)
    for dt, denom in ((np.uint8, 255.0), (np.uint16, 65535.0)):
        a = rs.randint(0, np.iinfo(dt).max + 1, size=(n,)).astype(dt)
        src = torch.from_numpy(a.view(np.int16) if dt == np.uint16 else a).cuda() if n else torch.empty(16, dtype=torch.uint8, device='cuda')
        dst = torch.full((max(n, 1),), -1.0, device='cuda')
        L.ni_feed_convert(ptr(src), a.dtype.itemsize, ptr(dst), n, denom, stream())
        assert np.array_equal(dst.cpu().numpy()[:n], (a.astype(np.float64) / denom).astype(np.float32))


@pytest.mark.gpu
def test_device_feed_double_buffers_and_converts():
    from neural_imaging_b200.helpers.dataset import DeviceFeed
    rs = np.random.RandomState(1)
    feed = DeviceFeed()
    batches = [(rs.randint(0, 65536, size=(4, 16, 16, 4)).astype(np.uint16), rs.randint(0, 256, size=(4, 32, 32, 3)).astype(np.uint8)) for _ in range(5)]
    pinned = [(torch.from_numpy(x.view(np.int16)).pin_memory(), torch.from_numpy(y).pin_memory()) for x, y in batches]
    feed.submit(*pinned[0])
    for i in range(5):
        x, y = feed.next()
        if i + 1 < 5:
            feed.submit(*pinned[i + 1])
        got_x, got_y = x.clone(), y.clone()              # "the step": consumes the batch on the compute stream
        feed.release()
        assert np.array_equal(got_x.cpu().numpy(), (batches[i][0].astype(np.float64) / 65535).astype(np.float32))
        assert np.array_equal(got_y.cpu().numpy(), (batches[i][1].astype(np.float64) / 255).astype(np.float32))
    f32 = torch.from_numpy(rs.uniform(size=(2, 8, 8, 4)).astype(np.float32)).pin_memory()
    feed.submit(f32)
    assert torch.equal(feed.next().cpu(), f32)
    with pytest.raises(RuntimeError):
        feed.next()
