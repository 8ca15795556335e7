#!/usr/bin/env python
"""Benchmark of the hot path: the joint UNet + manipulations + dJPEG(50) + FAN training step (BASELINE.json config 4).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation (restated; TF absent)

Prints ONE JSON line (rank 0). `value` = raw 128x128 patches / s with inputs resident in HBM; `e2e` = the same through
helpers.dataset.DeviceFeed + ManipulationClassification.training_step_device with pinned-host inputs (H2D of every
step inside the timed region, overlapped with the previous step's compute) and a D2H loss read.
Global batch is fixed at 256 raw patches (strong scaling): each rank processes 256/N patches = 1280/N codec/FAN images.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GLOBAL_BATCH = 256
RAW = 128
LAMBDA_NIP = 0.1
LR = 1e-4
METRIC = 'patches_per_sec_unet_djpeg50_fan_train_step'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'], 'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


class EventProfiler:
    """Brackets every C-ABI call with CUDA events; aggregates device time, algorithmic FLOPs and bytes per entry point."""

    def __init__(self):
        self.records = []
        self._open = None

    def before(self, name, args):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self._open = e

    def after(self, name, args):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        key = name
        if name.startswith('ni_conv2d_'):
            d = args[0]._obj
            key = '%s n%d %dx%d c%d->%d k%d s%d' % (name[10:], d.n, d.h, d.w, d.cin, d.cout, d.kh, d.stride)
        self.records.append((name, self._open, e, self._work(name, args), key))

    def layer_table(self):
        torch.cuda.synchronize()
        agg = {}
        for name, a, b, (kind, work), key in self.records:
            r = agg.setdefault(key, {'calls': 0, 'ms': 0.0, 'flop': 0.0})
            r['calls'] += 1
            r['ms'] += a.elapsed_time(b)
            r['flop'] += work if kind == 'flop' else 0.0
        rows = []
        for k, r in sorted(agg.items(), key=lambda kv: -kv[1]['ms']):
            rows.append({'key': k, 'calls': r['calls'], 'ms': r['ms'], 'tflops': (r['flop'] / (r['ms'] * 1e-3) / 1e12) if r['ms'] > 0 and r['flop'] else None})
        return rows

    @staticmethod
    def _work(name, args):
        if name.startswith('ni_conv2d_'):
            d = args[0]._obj
            return ('flop', 2.0 * d.n * d.oh * d.ow * d.cout * d.cin * d.kh * d.kw)
        if name == 'ni_djpeg_fwd':
            n, h, w = args[3], args[4], args[5]
            return ('byte', 24.0 * n * h * w)
        if name == 'ni_djpeg_bwd':
            n, h, w = args[3], args[4], args[5]
            return ('byte', 36.0 * n * h * w)
        return ('none', 0.0)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, a, b, (kind, work), _key in self.records:
            r = agg.setdefault(name, {'calls': 0, 'ms': 0.0, 'kind': kind, 'work': 0.0})
            r['calls'] += 1
            r['ms'] += a.elapsed_time(b)
            r['work'] += work
        return agg


def make_inputs(b, seed):
    rs = np.random.RandomState(seed)
    x = rs.uniform(size=(b, RAW, RAW, 4)).astype(np.float32)
    y = rs.uniform(size=(b, 2 * RAW, 2 * RAW, 3)).astype(np.float32)
    return x, y


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from neural_imaging_b200 import _lib
    from neural_imaging_b200.parallel import GradSync, broadcast_parameters
    from neural_imaging_b200.workflows.manipulation_classification import ManipulationClassification
    dev = torch.device('cuda', local_rank)
    L = _lib.lib()
    gb = args.batch
    assert gb % world == 0
    bl = gb // world
    flow = ManipulationClassification('UNet', trainable={'nip'}, raw_patch_size=RAW, seed=1234)
    sync = GradSync() if world > 1 else None
    if not args.no_graph:
        flow.enable_cuda_graph()       # static step (augment=False, fixed quality): two captured graphs instead of ~220 launches
    if world > 1:
        broadcast_parameters(flow._stores)
    xh, yh = make_inputs(gb, 1234)
    xh, yh = xh[rank * bl:(rank + 1) * bl], yh[rank * bl:(rank + 1) * bl]
    xp, yp = torch.from_numpy(xh).pin_memory(), torch.from_numpy(yh).pin_memory()
    xd, yd = xp.to(dev), yp.to(dev)

    def step_resident():
        return flow.training_step_device(xd, yd, LAMBDA_NIP, 0, False, LR, grad_sync=sync)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {}

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        host_ms[fn.__name__] = (time.perf_counter() - t0) * 1e3 / steps      # host time to ENQUEUE a step (no sync inside)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        loss, parts = step_resident()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    L.ni_reset_launch_count()
    ms = timed(step_resident, args.steps)
    launches = int(L.ni_launch_count())
    if flow.graph_launches_per_step:        # graph replays do not pass through the host-side launch counter
        launches = flow.graph_launches_per_step * args.steps
    clk = clocks.stop() if rank == 0 else None

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H loss, every step, through the public API a training loop uses:
    # helpers.dataset.DeviceFeed (double-buffered: the H2D copy of batch i + 1 is enqueued on the copy stream before step i is
    # launched, so it runs under that step's compute) + ManipulationClassification.training_step_device + loss.numpy()
    from neural_imaging_b200.helpers.dataset import DeviceFeed

    def make_e2e(hx, hy):
        feed = DeviceFeed()
        feed.submit(hx, hy)

        def step():
            xe, ye = feed.next()
            feed.submit(hx, hy)                   # next step's inputs: same pinned buffers, a full H2D copy every step
            loss, _ = flow.training_step_device(xe, ye, LAMBDA_NIP, 0, False, LR, grad_sync=sync)
            feed.release()
            return float(loss.numpy())            # device -> host read of the step's loss (synchronises)
        return step, feed
    step_e2e, feed32 = make_e2e(xp, yp)
    for _ in range(2):
        last = step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    # same with the batches in their stored integer form (uint16 RGGB stacks, uint8 RGB: helpers/loading.py:69-71): 4x fewer
    # bytes over PCIe, converted on the device by ni_feed_convert (bit-identical to the reference's host conversion)
    xi = torch.from_numpy(np.round(xh * 65535).astype(np.uint16).view(np.int16)).pin_memory()
    yi = torch.from_numpy(np.round(yh * 255).astype(np.uint8)).pin_memory()
    step_e2e_int, feed_int = make_e2e(xi, yi)
    for _ in range(2):
        step_e2e_int()
    ms_e2e_int = timed(step_e2e_int, args.steps)
    if flow._optimizer.nonfinite():
        raise RuntimeError('non-finite gradients during the benchmark')

    # ---- per-kernel roofline: same steps again with every C-ABI call bracketed by CUDA events (own pass so that the
    # event records do not perturb `value`)
    flow.enable_cuda_graph(False)
    prof = EventProfiler()
    _lib.PROFILER = prof
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step_resident()
    t1.record()
    _lib.PROFILER = None
    agg = prof.summary()
    if args.layer_report and rank == 0:
        with open(args.layer_report, 'w') as f:
            json.dump({'steps': args.steps, 'batch': args.batch, 'layers': prof.layer_table()}, f, indent=1)
    prof_ms = t0.elapsed_time(t1)
    if rank != 0:
        return
    pk = peaks()
    kernels = []
    for name, r in sorted(agg.items(), key=lambda kv: -kv[1]['ms']):
        e = {'entry': name, 'calls_per_step': r['calls'] / args.steps, 'ms_per_step': r['ms'] / args.steps,
             'share': r['ms'] / max(prof_ms, 1e-9)}
        if r['kind'] == 'flop' and r['ms'] > 0:
            e['tflops'] = r['work'] / (r['ms'] * 1e-3) / 1e12
        if r['kind'] == 'byte' and r['ms'] > 0:
            e['gbs'] = r['work'] / (r['ms'] * 1e-3) / 1e9
            e['frac_of_hbm_peak'] = e['gbs'] / pk['hbm_gbs']
        kernels.append(e)
    conv = [k for k in kernels if k['entry'].startswith('ni_conv2d_')]
    conv_ms = sum(k['ms_per_step'] for k in conv)
    conv_flop = sum(agg[k['entry']]['work'] for k in conv) / args.steps
    dj = next((k for k in kernels if k['entry'] == 'ni_djpeg_fwd'), None)
    conv_tf = conv_flop / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    # DRAM traffic per launch comes from the committed `ncu --set full` captures (profiles/r1_ncu_traffic.json, written by
    # tools/ncu_traffic.py from the .ncu-rep files): it cannot be measured live without a profiler attached.
    traffic = {}
    tpath = os.path.join(ROOT, 'profiles', 'r1_ncu_traffic.json')
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath))
    n_conv_launches = sum(k['calls_per_step'] for k in conv)
    roofline = {'kernel': 'conv2d tcgen05 3xTF32 implicit GEMM + direct FP32 stencils (fprop+dgrad+wgrad, all %d conv launches of the step)' % int(n_conv_launches),
                'bound': 'tensor', 'achieved': conv_tf, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s', 'frac': conv_tf / pk['bf16_tflops_sustained'],
                'traffic': traffic.get('conv_top_launch', {}).get('dram_bytes'), 'traffic_launch': traffic.get('conv_top_launch'),
                'flop_per_launch_avg': conv_flop / max(n_conv_launches, 1), 'ms_per_launch_avg': conv_ms / max(n_conv_launches, 1),
                'share_of_step': conv_ms / max(prof_ms / args.steps, 1e-9), 'peak_source': pk['source'],
                # measured with tools/hw_probes.py (profiles/r1_hw_probes.txt): a kind::tf32 MMA (M 128, K 8, A in tensor memory) takes
                # 20.5 + 0.42 N cycles; a 32-deep 3xTF32 k-iteration is 12 of them => 2*128*N*32 / (12 * (20.5 + 0.42 N)) flop/clk/SM
                'ceiling_3xtf32_tflops_at_n128': 2 * 128 * 128 * 32 / (12 * (20.5 + 0.42 * 128)) * 148 * 1.965e9 / 1e12,
                'frac_of_3xtf32_ceiling': conv_tf / (2 * 128 * 128 * 32 / (12 * (20.5 + 0.42 * 128)) * 148 * 1.965e9 / 1e12),
                'note': 'FP32 results (1e-5 parity) => 3xTF32 (three tensor-core passes per product) / FP32 SIMT; the denominator is the dense bf16 cuBLAS peak, '
                        'so 1/6 of it is the ceiling of an ideal 3xTF32 kernel; traffic = DRAM bytes of the single most expensive launch (ncu), see traffic_launch'}
    roofline_djpeg = None
    if dj is not None:
        dj_bytes = agg['ni_djpeg_fwd']['work'] / max(agg['ni_djpeg_fwd']['calls'], 1)
        roofline_djpeg = {'kernel': 'djpeg_fwd3_kernel (fused colour+DCT+quant+IDCT+colour)', 'bound': 'hbm', 'achieved': dj['gbs'],
                          'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': dj['gbs'] / pk['hbm_gbs'],
                          'traffic': traffic.get('djpeg_fwd', {}).get('dram_bytes'), 'traffic_launch': traffic.get('djpeg_fwd'),
                          'algorithmic_bytes_per_launch_avg': dj_bytes,
                          'bytes_per_launch_basis': '24 B/pixel (read x + write y); the two launches of the step (256 x 256x256 at q=80, 1280 x 128x128 at q=50) averaged; '
                                                    'timed inside the step (inputs partly L2-resident); tools/profile_djpeg.py times it alone with L2 flushed',
                          'peak_source': pk['source']}
    out = {
        'metric': METRIC, 'value': gb / (ms * 1e-3), 'unit': 'patches/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'BASELINE config 4: UNet(128x128x4 raw) -> [native,sharpen,resample,gaussian,jpeg80] -> avgpool2 -> dJPEG(50,soft) -> FAN(5 classes); fwd+bwd+Adam, trainable {fan,nip}, lambda_nip=0.1',
                   'global_batch': gb, 'per_gpu_batch': bl, 'codec_fan_images_per_step': 5 * gb, 'parallelism': 'dp%d' % world,
                   'l2': 'working set (multi-GB activations per step) >> 126 MB L2; no explicit flush needed'},
        'e2e': {'value': gb / (ms_e2e * 1e-3), 'unit': 'patches/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': int(xp.numel() * 4 + yp.numel() * 4) * world, 'd2h_bytes_per_step': 4 * world,
                'api': 'helpers.dataset.DeviceFeed (float32 host batches, H2D of step i+1 under the compute of step i) -> training_step_device -> loss.numpy()'},
        'e2e_integer_feed': {'value': gb / (ms_e2e_int * 1e-3), 'unit': 'patches/s', 'ms_per_step': ms_e2e_int,
                             'h2d_bytes_per_step': int(xi.numel() * 2 + yi.numel()) * world, 'd2h_bytes_per_step': 4 * world,
                             'api': 'as e2e, host batches as stored (uint16 RAW / uint8 RGB), converted on the device (ni_feed_convert)'},
        'gpu_launches': launches, 'gpu_launches_per_step': launches / args.steps, 'cuda_graph': not args.no_graph,
        'host_enqueue_ms_per_step': host_ms.get('step_resident'), 'clocks': clk, 'roofline': roofline, 'roofline_djpeg': roofline_djpeg, 'kernels': kernels[:12],
        'images_per_sec_codec_fan': 5 * gb / (ms * 1e-3), 'loss': float(loss.numpy()),
    }
    if world == 1 and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_baseline(bounded_batch=args.cpu_batch, steps=args.cpu_steps)
    print(json.dumps(out), flush=True)


def _host_state(cls, kwargs):
    """Initial (Keras-default, seeded) weights of a product model WITHOUT touching the GPU: parameter specs only."""
    from neural_imaging_b200 import nn
    nn.HOST_ONLY = True
    try:
        m = cls(**kwargs)
    finally:
        nn.HOST_ONLY = False
    return {p.name: p.init for p in m._store.params}


def cpu_step_factory(b):
    """The restated reference (oracle) joint training step on host cores; b raw patches of config 4."""
    from neural_imaging_b200.models import forensics, pipelines
    from oracle import ref_models as M
    state_nip = _host_state(pipelines.UNet, dict(patch_size=RAW, seed=1234))
    state_fan = _host_state(forensics.FAN, dict(n_classes=5, patch_size=RAW, seed=1234))
    Pn, Pf = M.to_params(state_nip), M.to_params(state_fan)
    x, y = make_inputs(b, 1234)
    xt, yt = torch.tensor(x), torch.tensor(y)
    opt = {'t': 0, 'm': {}, 'v': {}}
    return lambda: M.training_step(Pn, Pf, opt, xt, yt, lambda_nip=LAMBDA_NIP, lr=LR, train_nip=True)


def cpu_baseline(bounded_batch=32, steps=4):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_factory(bounded_batch)
    step()                                   # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {'value': bounded_batch / dt, 'unit': 'patches/s', 'cores': cores, 'kind': 'port',
            'sample': '{} raw patches/step ({} codec/FAN images), {} timed steps of the restated reference (PyTorch-CPU float32 op-for-op oracle; TensorFlow unavailable)'.format(
                bounded_batch, 5 * bounded_batch, steps), 's_per_step': dt}


def run_reference(args, rank, world):
    if rank != 0:
        return
    b = args.cpu_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_factory(b)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    k = max(1, min(args.steps, args.cpu_steps))
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    dt = (time.perf_counter() - t0) / k
    v = b / dt
    sample = '{} raw patches/step ({} codec/FAN images) of config 4, {} timed steps; restated reference (oracle/, PyTorch-CPU float32, {} threads) because TensorFlow 2.1 is not installable here'.format(b, 5 * b, k, cores)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'patches/s', 'n_gpus': world, 'steps': k, 'warmup': 1, 'ms_per_step': dt * 1e3,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'BASELINE config 4 (bounded sample): UNet -> manipulations -> avgpool2 -> dJPEG(50) -> FAN train step', 'global_batch': b, 'parallelism': 'cpu'},
        'cpu_baseline': {'value': v, 'unit': 'patches/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=GLOBAL_BATCH, help='global raw batch (BASELINE config 4: 256)')
    ap.add_argument('--cpu-batch', type=int, default=32, help='raw patches per step of the CPU legs (bounded sample of config 4: ~10 s of host work)')
    ap.add_argument('--cpu-steps', type=int, default=4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch the step kernel by kernel instead of replaying the captured CUDA graphs')
    ap.add_argument('--layer-report', default=None, help='write a per-layer (per conv shape) timing table to this JSON file')
    args = ap.parse_args()
    rank, world, local_rank = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        return run_reference(args, rank, world)
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
