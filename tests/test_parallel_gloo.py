"""CPU tests (gloo, world_size 2) of the data-parallel host logic: batch sharding, the gradient all-reduce hook with the 1/W
scale folded into the optimiser, parameter broadcast — and, with the oracle, that sharding the raw batch by rank and averaging
gradients reproduces the single-process step (mean-type losses, class-major labels per local batch)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Store:
    def __init__(self, n, rank):
        g = torch.Generator().manual_seed(100 + rank)
        self.flat = torch.randn(n, generator=g)
        self.gflat = torch.randn(n, generator=g)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from neural_imaging_b200.parallel import GradSync, broadcast_parameters, shard_batch
    stores = [_Store(1000, rank), _Store(37, rank)]
    expect = [sum(_Store(n, r).gflat for r in range(world)) for n in (1000, 37)]
    sync = GradSync()
    assert sync.world == world and abs(sync.gscale - 1.0 / world) < 1e-12
    sync(stores)
    ok = all(torch.allclose(s.gflat, e) for s, e in zip(stores, expect))
    broadcast_parameters(stores, src=0)
    ok = ok and all(torch.equal(s.flat, _Store(n, 0).flat) for s, n in zip(stores, (1000, 37)))
    batch = torch.arange(8 * 3).reshape(8, 3)
    mine = shard_batch(batch, rank, world)
    ok = ok and torch.equal(mine, batch[rank * 4:(rank + 1) * 4])
    try:
        shard_batch(torch.zeros(7, 3), rank, world)
        ok = False
    except ValueError:
        pass
    # gradient buffers laid out in one arena (nn.unify_gradients): the hook issues ONE collective over the whole bucket
    a, b = _Store(1000, rank), _Store(40, rank)
    arena = torch.cat((a.gflat, b.gflat))
    a.gflat, b.gflat = arena[:1000], arena[1000:]
    a.arena = b.arena = arena
    calls = []
    real = dist.all_reduce
    dist.all_reduce = lambda t, **kw: (calls.append(t.numel()), real(t, **kw))[1]
    try:
        sync([a, b])
    finally:
        dist.all_reduce = real
    ok = ok and calls == [1040]
    ok = ok and torch.allclose(a.gflat, sum(_Store(1000, r).gflat for r in range(world))) \
        and torch.allclose(b.gflat, sum(_Store(40, r).gflat for r in range(world)))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gradsync_shard_broadcast_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29731, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


def test_batch_sharded_gradients_equal_full_batch_oracle():
    """mean(CE) + lambda*mean(MSE) over the full batch == average over ranks of the per-shard losses' gradients."""
    from oracle import ref_models as M
    rs = np.random.RandomState(0)
    B, ps, world = 4, 16, 2
    import bench
    from neural_imaging_b200.models import forensics, pipelines
    s_nip = bench._host_state(pipelines.UNet, dict(patch_size=ps, seed=1))
    s_fan = bench._host_state(forensics.FAN, dict(n_classes=3, patch_size=2 * ps, seed=1, n_filters=8, n_convolutions=2))
    x = rs.uniform(size=(B, ps, ps, 4)).astype(np.float32)
    t = rs.uniform(size=(B, 2 * ps, 2 * ps, 3)).astype(np.float32)
    names = ('gaussian', 'jpeg')

    def grads(xb, tb):
        Pn, Pf = M.to_params(s_nip, torch.float64), M.to_params(s_fan, torch.float64)
        Y = M.unet_forward(Pn, torch.tensor(xb, dtype=torch.float64))
        m = M.run_manipulations(Y, names)
        probs = M.fan_forward(Pf, m, n_convolutions=2)
        loss = M.R.sparse_categorical_crossentropy(M.batch_labels(xb.shape[0], 3), probs) + 0.1 * M.R.mse(torch.tensor(tb, dtype=torch.float64), Y)
        ps_ = list(Pn.values()) + list(Pf.values())
        return [g.numpy() for g in torch.autograd.grad(loss, ps_)]
    full = grads(x, t)
    parts = [grads(x[r * B // world:(r + 1) * B // world], t[r * B // world:(r + 1) * B // world]) for r in range(world)]
    for gf, *gs in zip(full, *parts):
        avg = sum(gs) / world
        assert np.max(np.abs(avg - gf)) <= 1e-9 * max(1.0, np.max(np.abs(gf)))


def _entropy_worker(rank, world, port, out):
    """Each rank holds a shard of the quantised latent; the 32-bin soft-histogram partial sums are all-reduced BEFORE the log
    (product: DCN.set_data_parallel -> one 32-double all-reduce), which must reproduce the single-process entropy and its
    gradient exactly — the per-shard entropy would not (it is smaller, by concavity)."""
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import ref_ops as R
    g = torch.Generator().manual_seed(7)
    z = (torch.randn((4, 6, 6, 8), generator=g, dtype=torch.float64) * torch.tensor([0.5, 1.0, 2.0, 4.0], dtype=torch.float64).view(4, 1, 1, 1))
    cb = torch.arange(-15, 17, dtype=torch.float64)
    full = z.clone().requires_grad_(True)
    h_full, _ = R.entropy(full, cb)
    g_full, = torch.autograd.grad(h_full.double(), full)
    per = z.shape[0] // world
    mine = z[rank * per:(rank + 1) * per].clone().requires_grad_(True)
    w = R._codebook_weights(mine, cb)
    partial = w.sum(dim=0)
    total = partial.detach().clone()
    dist.all_reduce(total)                                           # the one data-path collective of the DCN forward
    n_total = z.numel()
    hist = (total / n_total).clamp(1e-9, float(np.finfo(np.float32).max))
    hist = hist / hist.sum()
    h_global = float(-(hist * torch.log(hist)).sum() / 0.6931)
    # backward re-uses the GLOBAL histogram: dH/dhist at the global point, chained through this rank's own weights
    hg = hist.clone().requires_grad_(True)
    dh, = torch.autograd.grad(-(hg * torch.log(hg)).sum() / 0.6931, hg)
    s = (total / n_total).sum()
    dpartial = (dh - (dh * hist).sum()) / s / n_total                # through hist / hist.sum() and the mean over all values
    g_mine, = torch.autograd.grad(partial, mine, grad_outputs=dpartial)
    h_shard, _ = R.entropy(mine.detach(), cb)
    ok = abs(h_global - float(h_full)) < 1e-6 and float((g_mine - g_full[rank * per:(rank + 1) * per]).abs().max()) < 1e-9 * max(1.0, float(g_full.abs().max()))
    out[rank] = (bool(ok), float(h_shard), float(h_full))
    dist.destroy_process_group()


def test_dcn_entropy_from_allreduced_histogram_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_entropy_worker, args=(world, 29741, out), nprocs=world, join=True)
    assert all(out[r][0] for r in range(world)), dict(out)
    assert any(out[r][1] < out[r][2] - 1e-3 for r in range(world))      # a per-shard entropy is NOT the global one
