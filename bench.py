#!/usr/bin/env python
"""Benchmark of the hot path: the joint UNet + manipulations + dJPEG(50) + FAN training step (BASELINE.json config 4).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation (restated; TF absent)
    python bench.py --config c1|c2|c3|c5 ...                 # the other BASELINE.json configurations (c4 = default = headline):
        c1 dJPEG round trip q=50 (1280 x 128x128x3 roofline run + the 256x256 image of test_jpeg.py), c2 UNet pretrain B=32,
        c3 TwitterDCN-32C pretrain B=64 (1 -> 8 GPUs, batch-global entropy via a 32-double histogram all-reduce),
        c5 UNet + TwitterDCN + FAN end to end, trainable {fan, nip, dcn} (--batch, or --sweep for 64 .. 1024)

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
LAMBDA_DCN = 0.1          # config/tests/framework.json:55


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
        # algorithmic bytes of the other HBM-bound entry points (float32 NHWC; DESIGN.md section 4 states each formula)
        if name == 'ni_maxpool2_act_bwd_bias':          # read y (full), read d(pooled), write dy (full)
            n, h, w, c = args[5], args[6], args[7], args[8]
            return ('byte', 4.0 * n * h * w * c * (1 + 0.25 + 1))
        if name == 'ni_maxpool2_fwd':                   # read x, write pooled
            n, h, w, c = args[2], args[3], args[4], args[5]
            return ('byte', 4.0 * n * h * w * c * 1.25)
        if name == 'ni_act_bwd_bias':                   # read y (when the activation needs it), read + write dy
            n, h, w, c = args[3], args[4], args[5], args[6]
            return ('byte', 4.0 * n * h * w * c * (3 if args[0] else 2))
        if name == 'ni_manip_stack_pool2_fwd':          # read Y (12 B/px), write 4 pooled slots (4 x 3 B/px) + 1 mask byte
            b, h, w = args[3], args[4], args[5]
            return ('byte', 25.0 * b * h * w)
        if name == 'ni_manip_stack_pool2_bwd':          # read 3 pooled gradient slots (9 B/px) + mask, read + write dY (24 B/px)
            b, h, w = args[3], args[4], args[5]
            return ('byte', 34.0 * b * h * w)
        if name == 'ni_maxpool2_code_bwd_bias':         # read the code byte + d(pooled), write dy (full resolution)
            n, oh, ow, c = args[4], args[5], args[6], args[7]
            return ('byte', n * oh * ow * c * (1.0 + 4.0 + 16.0))
        if name in ('ni_cconv5_fwd', 'ni_cconv5_bwd_data', 'ni_cconv5_bwd_filter'):      # 225 MAC per pixel: FP32-FMA bound, not HBM (AI 18.8 flop/B)
            n, h, w = args[3], args[4], args[5]
            return ('flop', 2.0 * 225 * n * h * w)
        if name in ('ni_avgpool_fwd', 'ni_avgpool_bwd'):
            n, h, w = args[2], args[3], args[4]
            return ('byte', 15.0 * n * h * w)
        if name == 'ni_adam_keras' or name == 'ni_adam_keras_dev':
            return ('byte', 28.0 * args[4])
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


def param_checksum(stores):
    """Sum and L2 norm (float64) of every trained parameter after the timed steps: equal step counts at N = 1 / 2 / 4 / 8 must land on
    (almost) the same numbers if the data-parallel step equals the single-GPU step (differences: summation order + rounding ties)."""
    tot, sq = 0.0, 0.0
    for s in stores:
        f = s.flat.double()
        tot += float(f.sum().item())
        sq += float((f * f).sum().item())
    return {'sum': tot, 'l2': sq ** 0.5}


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
    c5 = args.config == 'c5'
    lam_dcn = LAMBDA_DCN if c5 else 0.0
    if c5:      # TwitterDCN-32C constructed directly instead of codec.restore: no pre-trained model files exist offline
        flow = ManipulationClassification('UNet', trainable={'nip', 'dcn'}, raw_patch_size=RAW, seed=1234,
                                          distribution={'downsampling': 'pool:2', 'compression': 'dcn', 'compression_params': {'patch_size': RAW}})
        if world > 1:
            flow.codec.set_data_parallel(world)
    else:
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
        return flow.training_step_device(xd, yd, LAMBDA_NIP, lam_dcn, False, LR, grad_sync=sync)

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
        for i in range(steps):
            fn()
            if i == 0:
                # host time to ENQUEUE one step into an empty queue (later steps block on the driver's launch queue once the host has run
                # far enough ahead of the device, which measures the device, not the host)
                host_ms[fn.__name__] = (time.perf_counter() - t0) * 1e3
        e1.record()
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
            loss, _ = flow.training_step_device(xe, ye, LAMBDA_NIP, lam_dcn, False, LR, grad_sync=sync)
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
    checksum = param_checksum(flow._stores)
    checksum['optimizer_steps'] = int(flow._optimizer.iterations)

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
        if _bound_of(name):
            e['bound'] = _bound_of(name)
        kernels.append(e)
    conv = [k for k in kernels if k['entry'].startswith(('ni_conv2d_', 'ni_cconv5_'))]       # every convolution launch, incl. the constrained 5x5 3->3 filter
    conv_ms = sum(k['ms_per_step'] for k in conv)
    conv_flop = sum(agg[k['entry']]['work'] for k in conv) / args.steps
    dj = next((k for k in kernels if k['entry'] == 'ni_djpeg_fwd'), None)
    conv_tf = conv_flop / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    # DRAM traffic per launch comes from the committed `ncu --set full` captures (profiles/r1_ncu_traffic.json, written by
    # tools/ncu_traffic.py from the .ncu-rep files): it cannot be measured live without a profiler attached.
    traffic = {}
    for tname in ('r2_ncu_traffic.json', 'r1_ncu_traffic.json'):
        tpath = os.path.join(ROOT, 'profiles', tname)
        if os.path.isfile(tpath):
            traffic = json.load(open(tpath))
            traffic['source'] = 'profiles/' + tname
            break
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
        roofline_djpeg = {'kernel': 'djpeg_fwd4_kernel (persistent CTAs, TMA-fed stage; fused colour+DCT+quant+IDCT+colour)', 'bound': 'hbm', 'achieved': dj['gbs'],
                          'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': dj['gbs'] / pk['hbm_gbs'],
                          'traffic': traffic.get('djpeg_fwd', {}).get('dram_bytes'), 'traffic_launch': traffic.get('djpeg_fwd'),
                          'algorithmic_bytes_per_launch_avg': dj_bytes,
                          'bytes_per_launch_basis': '24 B/pixel (read x + write y); the two launches of the step (256 x 256x256 at q=80, 1280 x 128x128 at q=50) averaged; '
                                                    'timed inside the step (inputs partly L2-resident); tools/profile_djpeg.py times it alone with L2 flushed',
                          'peak_source': pk['source']}
    lat = next((k for k in kernels if k['entry'] in ('ni_latent_quantise_fwd', 'ni_latent_softcodebook_fwd')), None)
    metric = METRIC if not c5 else 'patches_per_sec_unet_twitterdcn_fan_train_step'
    workload = ('BASELINE config 4: UNet(128x128x4 raw) -> [native,sharpen,resample,gaussian,jpeg80] -> avgpool2 -> dJPEG(50,soft) -> FAN(5 classes); '
                'fwd+bwd+Adam, trainable {fan,nip}, lambda_nip=0.1') if not c5 else \
               ('BASELINE config 5: UNet(128x128x4 raw) -> [native,sharpen,resample,gaussian,jpeg80] -> avgpool2 -> TwitterDCN-32C (soft-codebook, 5 bpf, '
                'float64 latent path) -> FAN(5 classes); fwd+bwd+Adam, trainable {fan,nip,dcn}, lambda_nip=0.1, lambda_dcn=0.1')
    out = {
        'metric': metric, 'value': gb / (ms * 1e-3), 'unit': 'patches/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload, 'name': args.config,
                   'global_batch': gb, 'per_gpu_batch': bl, 'codec_fan_images_per_step': 5 * gb, 'parallelism': 'dp%d' % world,
                   'l2': 'working set (multi-GB activations per step) >> 126 MB L2; no explicit flush needed'},
        'e2e': {'value': gb / (ms_e2e * 1e-3), 'unit': 'patches/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': int(xp.numel() * 4 + yp.numel() * 4) * world, 'd2h_bytes_per_step': 4 * world,
                'api': 'helpers.dataset.DeviceFeed (float32 host batches, H2D of step i+1 under the compute of step i) -> training_step_device -> loss.numpy()'},
        'e2e_integer_feed': {'value': gb / (ms_e2e_int * 1e-3), 'unit': 'patches/s', 'ms_per_step': ms_e2e_int,
                             'h2d_bytes_per_step': int(xi.numel() * 2 + yi.numel()) * world, 'd2h_bytes_per_step': 4 * world,
                             'api': 'as e2e, host batches as stored (uint16 RAW / uint8 RGB), converted on the device (ni_feed_convert)'},
        'gpu_launches': launches, 'gpu_launches_per_step': launches / args.steps, 'cuda_graph': not args.no_graph,
        'host_enqueue_ms_per_step': host_ms.get('step_resident'), 'clocks': clk, 'roofline': roofline, 'roofline_djpeg': roofline_djpeg, 'kernels': kernels[:16],
        'images_per_sec_codec_fan': 5 * gb / (ms * 1e-3), 'loss': float(loss.numpy()), 'param_checksum': checksum,
    }
    if lat is not None:
        out['roofline_latent'] = {'kernel': 'latent_softcodebook_fwd_kernel (scale, 32-entry soft code book in float64, hard value, soft histogram)',
                                  'bound': 'hbm', 'calls_per_step': lat['calls_per_step'], 'ms_per_step': lat['ms_per_step'],
                                  'note': 'bytes: 4 B read + 4 B written per latent value (M x 16 x 16 x 32 values); the float64 kernel weights of '
                                          'the reference (M x 8192 x 32 x 8 B, twice) never exist in memory'}
    if args.sweep:
        sweep = []
        for b in (64, 128, 256, 512, 1024):
            if b % world:
                continue
            torch.cuda.synchronize()
            flow.release_workspaces()          # buffers are kept per shape (graph safety): free the previous batch size first
            xs, ys = make_inputs(b // world, 4321 + rank)
            xs, ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
            fn = lambda: flow.training_step_device(xs, ys, LAMBDA_NIP, lam_dcn, False, LR, grad_sync=sync)
            try:
                for _ in range(3):
                    fn()
                m = timed(fn, max(3, args.steps // 2))
                sweep.append({'global_batch': b, 'ms_per_step': m, 'value': b / (m * 1e-3),
                              'peak_hbm_gb': torch.cuda.max_memory_allocated() / 1e9})
            except torch.OutOfMemoryError:
                sweep.append({'global_batch': b, 'ms_per_step': None, 'value': None, 'note': 'does not fit 180 GB on %d GPU(s)' % world})
            torch.cuda.reset_peak_memory_stats()
            del xs, ys
        out['sweep'] = sweep
    if world == 1 and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_baseline(args, bounded_batch=args.cpu_batch or (gb if args.config == 'c4' else min(gb, 64)), steps=args.cpu_steps)
    print(json.dumps(out), flush=True)


# ================================================================================================ configs 1 - 3 (stand-alone paths)
def _harness(world, dev):
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())
    return timed


def _profiled(step, steps):
    """Same steps with every C-ABI call bracketed by CUDA events -> (per-entry aggregate, wall ms of the pass)."""
    from neural_imaging_b200 import _lib
    prof = EventProfiler()
    _lib.PROFILER = prof
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    _lib.PROFILER = None
    return prof.summary(), t0.elapsed_time(t1)


def _conv_roofline(agg, steps, prof_ms, pk):
    conv = {k: r for k, r in agg.items() if k.startswith(('ni_conv2d_', 'ni_cconv5_'))}
    ms = sum(r['ms'] for r in conv.values()) / steps
    flop = sum(r['work'] for r in conv.values()) / steps
    n = sum(r['calls'] for r in conv.values()) / steps
    tf = flop / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    return {'kernel': 'conv2d tcgen05 3xTF32 implicit GEMM + direct FP32 stencils (all %d conv launches of the step)' % int(n), 'bound': 'tensor',
            'achieved': tf, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s', 'frac': tf / pk['bf16_tflops_sustained'], 'traffic': None,
            'flop_per_launch_avg': flop / max(n, 1), 'ms_per_launch_avg': ms / max(n, 1), 'share_of_step': ms / max(prof_ms / steps, 1e-9),
            'peak_source': pk['source'], 'note': 'FP32 results (1e-5 parity) => 3xTF32: 1/6 of the bf16 peak is the ceiling of an ideal kernel'}


# What bounds each entry point (ncu evidence: profiles/r2_*_ncu.txt). The fused stencil kernels move so few bytes per pixel BY DESIGN
# (fusion removed the intermediate tensors) that instruction issue, not HBM, is their limit: their gbs figure is informational.
BOUND = {'ni_manip_stack_pool2_fwd': 'fp32-issue (four fused stencils + pool per pass; issue slots 81 % busy, DRAM = algorithmic)',
         'ni_manip_stack_pool2_bwd': 'fp32-issue / latency (issue slots 57 % busy, 59 registers)',
         'ni_djpeg_fwd': 'hbm', 'ni_djpeg_bwd': 'fp32-issue (recomputes the forward chain: ~190 thread-instructions per pixel)',
         'ni_maxpool2_act_bwd_bias': 'hbm', 'ni_maxpool2_fwd': 'hbm', 'ni_act_bwd_bias': 'hbm', 'ni_maxpool2_code_bwd_bias': 'hbm',
         'ni_avgpool_fwd': 'hbm', 'ni_avgpool_bwd': 'hbm', 'ni_adam_keras': 'hbm', 'ni_adam_keras_dev': 'hbm',
         'ni_cconv5_fwd': 'fp32-fma', 'ni_cconv5_bwd_data': 'fp32-fma', 'ni_cconv5_bwd_filter': 'fp32-fma', 'ni_conv2d_pool2_fwd': 'fp32-fma'}


def _bound_of(name):
    if name in BOUND:
        return BOUND[name]
    if name.startswith('ni_conv2d_'):
        return 'tensor (3xTF32) / fp32-fma for the 3-, 4- and 12-channel layers'
    return None


def _kernel_list(agg, steps, prof_ms, pk):
    out = []
    for name, r in sorted(agg.items(), key=lambda kv: -kv[1]['ms'])[:10]:
        e = {'entry': name, 'calls_per_step': r['calls'] / steps, 'ms_per_step': r['ms'] / steps, 'share': r['ms'] / max(prof_ms, 1e-9)}
        if r['kind'] == 'flop' and r['ms'] > 0:
            e['tflops'] = r['work'] / (r['ms'] * 1e-3) / 1e12
        if r['kind'] == 'byte' and r['ms'] > 0:
            e['gbs'] = r['work'] / (r['ms'] * 1e-3) / 1e9
            e['frac_of_hbm_peak'] = e['gbs'] / pk['hbm_gbs']
        if _bound_of(name):
            e['bound'] = _bound_of(name)
        out.append(e)
    return out


def _emit(args, world, rank, ms, ms_e2e, units, h2d, d2h, launches, clk, extra):
    if rank != 0:
        return
    out = {'metric': METRICS[args.config], 'value': units / (ms * 1e-3), 'unit': UNITS[args.config], 'n_gpus': world, 'steps': args.steps,
           'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
           'data': 'synthetic',
           'e2e': {'value': units / (ms_e2e * 1e-3), 'unit': UNITS[args.config], 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': int(h2d) * world,
                   'd2h_bytes_per_step': int(d2h) * world},
           'gpu_launches': launches, 'gpu_launches_per_step': launches / args.steps, 'clocks': clk}
    out.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_baseline(args, bounded_batch=args.cpu_batch or min(args.batch, 256), steps=args.cpu_steps)
    print(json.dumps(out), flush=True)


def run_c1(args, rank, world, local_rank):
    """config 1: differentiable JPEG round trip, q = 50. The roofline run is the batched tensor of the codec stage of config 4
    (global batch images of 128x128x3, forward = JPEG.process); the 256x256 single image of test_jpeg.py:105-110 is timed beside it."""
    from neural_imaging_b200 import _lib, ops
    from neural_imaging_b200.compression.jpeg_helpers import jpeg_qtable
    from neural_imaging_b200.models import jpeg
    dev = torch.device('cuda', local_rank)
    L = _lib.lib()
    timed = _harness(world, dev)
    n = args.batch // world
    rs = np.random.RandomState(1234 + rank)
    xh = torch.from_numpy(rs.uniform(size=(n, RAW, RAW, 3)).astype(np.float32)).pin_memory()
    yh = torch.empty_like(xh).pin_memory()
    xd, dyd = xh.to(dev), torch.from_numpy(rs.normal(size=tuple(xh.shape)).astype(np.float32)).to(dev)
    yd, dxd = torch.empty_like(xd), torch.empty_like(xd)
    codec = jpeg.JPEG(50, 'soft')
    ql, qc = jpeg_qtable(50, 0), jpeg_qtable(50, 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def fwd():
        codec._model.forward_into(xd, yd)

    def fwd_bwd():
        ops.djpeg_fwd(xd, ql, qc, 'soft', out=yd)
        ops.djpeg_bwd(xd, dyd, ql, qc, 'soft', out=dxd)

    def e2e():
        xd.copy_(xh, non_blocking=True)
        codec._model.forward_into(xd, yd)
        yh.copy_(yd, non_blocking=True)
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        fwd(); fwd_bwd()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    L.ni_reset_launch_count()
    ms = timed(fwd, args.steps)
    launches = int(L.ni_launch_count())
    ms_fb = timed(fwd_bwd, args.steps)
    e2e(); ms_e2e = timed(e2e, args.steps)
    # single launches with L2 flushed in between (input + output = 503 MB per launch is already 4x the L2; the flush removes the tail)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record(); fwd(); b.record()
    torch.cuda.synchronize()
    ms_single = float(np.median([a.elapsed_time(b) for a, b in ev]))
    evb = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evb:
        flush.fill_(1)
        a.record(); ops.djpeg_bwd(xd, dyd, ql, qc, 'soft', out=dxd); b.record()
    torch.cuda.synchronize()
    ms_bwd = float(np.median([a.elapsed_time(b) for a, b in evb]))
    # the sampler covers every timed loop of this configuration (the K-step forward loop alone lasts ~1 ms: below nvidia-smi's period)
    clk = clocks.stop() if rank == 0 else None
    one = torch.from_numpy(np.random.RandomState(7).uniform(size=(1, 256, 256, 3)).astype(np.float32)).to(dev)
    one_y = torch.empty_like(one)
    for _ in range(3):
        codec._model.forward_into(one, one_y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        codec._model.forward_into(one, one_y)
    e1.record(); torch.cuda.synchronize()
    pk = peaks()
    px = n * RAW * RAW
    gbs = 24.0 * px / (ms_single * 1e-3) / 1e9
    gbs_b = 36.0 * px / (ms_bwd * 1e-3) / 1e9
    extra = {'config': {'workload': 'BASELINE config 1: dJPEG(50, soft) round trip (JPEG.process) of {} x 128x128x3 float32 images per step'.format(args.batch),
                        'name': 'c1', 'global_batch': args.batch, 'parallelism': 'dp%d' % world, 'l2': 'input + output of a launch = 503 MB > 126 MB L2; single-launch timings flush L2'},
             'roofline': {'kernel': 'djpeg_fwd kernel (fused colour + 8x8 DCT + quantisation + IDCT + colour)', 'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm_gbs'],
                          'unit': 'GB/s', 'frac': gbs / pk['hbm_gbs'], 'traffic': None, 'algorithmic_bytes_per_launch': 24.0 * px,
                          'ms_per_launch': ms_single, 'peak_source': pk['source']},
             'roofline_bwd': {'kernel': 'djpeg_bwd kernel (forward chain recomputed, nothing saved)', 'bound': 'hbm', 'achieved': gbs_b, 'peak': pk['hbm_gbs'],
                              'unit': 'GB/s', 'frac': gbs_b / pk['hbm_gbs'], 'algorithmic_bytes_per_launch': 36.0 * px, 'ms_per_launch': ms_bwd},
             'fwd_bwd_ms_per_step': ms_fb, 'single_256x256_image_us': e0.elapsed_time(e1) * 10.0}
    _emit(args, world, rank, ms, ms_e2e, args.batch, xh.numel() * 4, yh.numel() * 4, launches, clk, extra)


def run_c2(args, rank, world, local_rank):
    """config 2: UNet pretraining step (train_nip.py: L2 loss, Adam), global batch 32 raw patches of 128x128x4."""
    from neural_imaging_b200 import _lib
    from neural_imaging_b200.models import pipelines
    from neural_imaging_b200.parallel import GradSync, broadcast_parameters
    dev = torch.device('cuda', local_rank)
    L = _lib.lib()
    timed = _harness(world, dev)
    bl = args.batch // world
    model = pipelines.UNet(patch_size=RAW, seed=1234)
    sync = GradSync() if world > 1 else None
    if world > 1:
        broadcast_parameters([model._store])
    xh, yh = make_inputs(args.batch, 1234)
    xp = torch.from_numpy(xh[rank * bl:(rank + 1) * bl]).pin_memory()
    yp = torch.from_numpy(yh[rank * bl:(rank + 1) * bl]).pin_memory()
    xd, yd = xp.to(dev), yp.to(dev)

    def step():
        return model.training_step(xd, yd, LR, grad_sync=sync)

    def e2e():
        return float(model.training_step(xp.to(dev, non_blocking=True), yp.to(dev, non_blocking=True), LR, grad_sync=sync).numpy())
    for _ in range(max(args.warmup, 3)):
        step()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    L.ni_reset_launch_count()
    ms = timed(step, args.steps)
    launches = int(L.ni_launch_count())
    clk = clocks.stop() if rank == 0 else None
    e2e(); ms_e2e = timed(e2e, args.steps)
    agg, prof_ms = _profiled(step, args.steps)
    pk = peaks()
    extra = {'config': {'workload': 'BASELINE config 2: UNet (7.76 M parameters) pretraining step, L2 loss, Keras Adam, 128x128x4 raw -> 256x256x3', 'name': 'c2',
                        'global_batch': args.batch, 'per_gpu_batch': bl, 'parallelism': 'dp%d' % world, 'l2': 'activations of a step (GBs) >> 126 MB L2'},
             'roofline': _conv_roofline(agg, args.steps, prof_ms, pk), 'kernels': _kernel_list(agg, args.steps, prof_ms, pk),
             'param_checksum': param_checksum([model._store])}
    _emit(args, world, rank, ms, ms_e2e, args.batch, xp.numel() * 4 + yp.numel() * 4, 4, launches, clk, extra)


def run_c3(args, rank, world, local_rank):
    """config 3: TwitterDCN-32C pretraining step (train_dcn.py: soft-codebook, 5 bpf, entropy weight 250), global batch 64 images."""
    from neural_imaging_b200 import _lib
    from neural_imaging_b200.models import compression
    from neural_imaging_b200.parallel import GradSync, broadcast_parameters
    dev = torch.device('cuda', local_rank)
    L = _lib.lib()
    timed = _harness(world, dev)
    bl = args.batch // world
    model = compression.TwitterDCN(patch_size=RAW, seed=1234)
    sync = GradSync() if world > 1 else None
    if world > 1:
        model.set_data_parallel(world)
        broadcast_parameters([model._store])
    xh = np.random.RandomState(1234).uniform(size=(args.batch, RAW, RAW, 3)).astype(np.float32)
    xp = torch.from_numpy(xh[rank * bl:(rank + 1) * bl]).pin_memory()
    xd = xp.to(dev)

    def step():
        return model.training_step(xd, LR, grad_sync=sync)

    def e2e():
        return float(model.training_step(xp.to(dev, non_blocking=True), LR, grad_sync=sync)['loss'].numpy())
    for _ in range(max(args.warmup, 3)):
        step()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    L.ni_reset_launch_count()
    ms = timed(step, args.steps)
    launches = int(L.ni_launch_count())
    clk = clocks.stop() if rank == 0 else None
    e2e(); ms_e2e = timed(e2e, args.steps)
    agg, prof_ms = _profiled(step, args.steps)
    pk = peaks()
    lat = {k: agg[k] for k in agg if k.startswith('ni_latent_') or k == 'ni_entropy_from_hist'}
    nz = bl * (RAW // 8) * (RAW // 8) * 32
    lat_ms = {k: r['ms'] / max(r['calls'], 1) for k, r in lat.items()}
    fwd_ms = lat_ms.get('ni_latent_quantise_fwd', lat_ms.get('ni_latent_softcodebook_fwd'))
    extra = {'config': {'workload': 'BASELINE config 3: TwitterDCN-32C (2.53 M parameters) training step: l2_loss (sum) + 250 x entropy of the batch-global soft '
                                    'histogram (float64 latent path), Keras Adam, 128x128x3 images', 'name': 'c3', 'global_batch': args.batch, 'per_gpu_batch': bl,
                        'parallelism': 'dp%d (32-double histogram all-reduce before the entropy + gradient all-reduce)' % world,
                        'l2': 'activations of a step (GBs) >> 126 MB L2'},
             'roofline': _conv_roofline(agg, args.steps, prof_ms, pk),
             'roofline_latent': {'kernel': 'latent quantisation forward (ni_latent_quantise_fwd: scale, soft code book in float64, hard value, soft histogram)', 'bound': 'hbm', 'unit': 'GB/s', 'ms_per_launch': fwd_ms,
                                 'algorithmic_bytes_per_launch': 8.0 * nz, 'achieved': (8.0 * nz / (fwd_ms * 1e-3) / 1e9) if fwd_ms else None,
                                 'peak': pk['hbm_gbs'], 'frac': (8.0 * nz / (fwd_ms * 1e-3) / 1e9 / pk['hbm_gbs']) if fwd_ms else None,
                                 'limited_by': 'fp64 issue (ncu profiles/r2_latent_ncu.txt: issue slots 72-75 % busy, DRAM = the latents once)',
                                 'note': '4 B read + 4 B written per latent value; {} values per launch: a few MB, i.e. not bandwidth-sized at this batch '
                                         '(64 float64 code-book weights per value are the work); the HBM figure is reported for reference'.format(nz), 'all_latent_entries_ms': lat_ms},
             'kernels': _kernel_list(agg, args.steps, prof_ms, pk), 'param_checksum': param_checksum([model._store])}
    _emit(args, world, rank, ms, ms_e2e, args.batch, xp.numel() * 4, 4, launches, clk, extra)


def _host_state(cls, kwargs):
    """Initial (Keras-default, seeded) weights of a product model WITHOUT touching the GPU: parameter specs only."""
    from neural_imaging_b200 import nn
    nn.HOST_ONLY = True
    try:
        m = cls(**kwargs)
    finally:
        nn.HOST_ONLY = False
    return {p.name: p.init for p in m._store.params}


def cpu_step_factory(b, config='c4'):
    """The restated reference (oracle, pinned to the executed reference by tests/test_tf_graph_golden.py) on host cores: one step of
    `config` over b units (raw patches; images for c1 / c3). Returns (step callable, description)."""
    from neural_imaging_b200.models import compression, forensics, pipelines
    from oracle import ref_models as M
    from oracle import ref_ops as R
    if config == 'c1':
        x = torch.tensor(np.random.RandomState(1234).uniform(size=(b, RAW, RAW, 3)).astype(np.float32))
        ql, qc = R.jpeg_qtable(50, 0), R.jpeg_qtable(50, 1)
        return (lambda: R.djpeg(x, ql, qc, 'soft')[0]), 'dJPEG(50, soft) forward of {} 128x128x3 images'.format(b)
    if config == 'c2':
        Pn = M.to_params(_host_state(pipelines.UNet, dict(patch_size=RAW, seed=1234)))
        x, y = make_inputs(b, 1234)
        xt, yt = torch.tensor(x), torch.tensor(y)
        opt = {'t': 0, 'm': {}, 'v': {}}
        names = list(Pn.keys())

        def step():
            loss = R.mse(M.unet_forward(Pn, xt), yt)
            g = torch.autograd.grad(loss, [Pn[k] for k in names])
            opt['t'] += 1
            with torch.no_grad():
                ms = [opt['m'].setdefault(k, torch.zeros_like(Pn[k])) for k in names]
                vs = [opt['v'].setdefault(k, torch.zeros_like(Pn[k])) for k in names]
                R.adam_keras_step([Pn[k] for k in names], list(g), ms, vs, opt['t'], LR)
            return loss
        return step, 'UNet L2 training step (tape + Keras Adam) on {} raw patches'.format(b)
    if config == 'c3':
        Pd = M.to_params(_host_state(compression.TwitterDCN, dict(patch_size=RAW, seed=1234)))
        xt = torch.tensor(np.random.RandomState(1234).uniform(size=(b, RAW, RAW, 3)).astype(np.float32))
        opt = {'t': 0, 'm': {}, 'v': {}}
        return (lambda: M.dcn_training_step(Pd, opt, xt, LR)), 'TwitterDCN-32C training step on {} 128x128x3 images'.format(b)
    state_nip = _host_state(pipelines.UNet, dict(patch_size=RAW, seed=1234))
    state_fan = _host_state(forensics.FAN, dict(n_classes=5, patch_size=RAW, seed=1234))
    Pn, Pf = M.to_params(state_nip), M.to_params(state_fan)
    x, y = make_inputs(b, 1234)
    xt, yt = torch.tensor(x), torch.tensor(y)
    opt = {'t': 0, 'm': {}, 'v': {}}
    if config == 'c5':
        Pd = M.to_params(_host_state(compression.TwitterDCN, dict(patch_size=RAW, seed=1234)))
        return (lambda: M.training_step(Pn, Pf, opt, xt, yt, lambda_nip=LAMBDA_NIP, lr=LR, train_nip=True, P_dcn=Pd, lambda_dcn=LAMBDA_DCN,
                                        train_dcn=True)), 'joint UNet + TwitterDCN + FAN step on {} raw patches ({} codec/FAN images)'.format(b, 5 * b)
    return (lambda: M.training_step(Pn, Pf, opt, xt, yt, lambda_nip=LAMBDA_NIP, lr=LR, train_nip=True)), \
        'joint UNet + dJPEG(50) + FAN step on {} raw patches ({} codec/FAN images)'.format(b, 5 * b)


def time_cpu(step, max_steps, budget_s):
    """Warm-up step, then as many timed steps as fit the budget (at least 2 when max_steps allows)."""
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    k = int(max(min(max_steps, 2), min(max_steps, budget_s // max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    return (time.perf_counter() - t0) / k, k


def cpu_baseline(args, bounded_batch, steps=2):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, what = cpu_step_factory(bounded_batch, args.config)
    dt, k = time_cpu(step, steps, 30.0)
    return {'value': bounded_batch / dt, 'unit': UNITS[args.config], 'cores': cores, 'kind': 'port',
            'sample': '{}; {} timed steps of the restated reference (PyTorch-CPU float32 op-for-op oracle, {} threads; TensorFlow unavailable)'.format(what, k, cores),
            's_per_step': dt}


UNITS = {'c1': 'images/s', 'c2': 'patches/s', 'c3': 'images/s', 'c4': 'patches/s', 'c5': 'patches/s'}
METRICS = {'c1': 'images_per_sec_djpeg50_roundtrip_128x128', 'c2': 'patches_per_sec_unet_pretrain_step', 'c3': 'images_per_sec_twitterdcn32c_train_step',
           'c4': METRIC, 'c5': 'patches_per_sec_unet_twitterdcn_fan_train_step'}
DEFAULT_BATCH = {'c1': 1280, 'c2': 32, 'c3': 64, 'c4': GLOBAL_BATCH, 'c5': GLOBAL_BATCH}


def run_reference(args, rank, world):
    """The reference arm: the restated reference on ALL host cores, on the SAME configuration (global batch) as our arm; the number of
    timed steps is what fits ~150 s (at least 2), so that the run ends within a few minutes whatever --steps asks for."""
    if rank != 0:
        return
    b = args.cpu_batch or args.batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, what = cpu_step_factory(b, args.config)
    dt, k = time_cpu(step, max(2, args.steps), 150.0)
    v = b / dt
    sample = '{}; {} timed steps after 1 warm-up; restated reference (oracle/, PyTorch-CPU float32, {} threads) because TensorFlow 2.1 is not installable here'.format(what, k, cores)
    print(json.dumps({
        'impl': 'reference', 'metric': METRICS[args.config], 'value': v, 'unit': UNITS[args.config], 'n_gpus': world, 'steps': k, 'warmup': 1, 'ms_per_step': dt * 1e3,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'BASELINE config {} on the host cores: {}'.format(args.config[1], what), 'name': args.config, 'global_batch': b, 'parallelism': 'cpu'},
        'cpu_baseline': {'value': v, 'unit': UNITS[args.config], 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': UNITS[args.config], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c4', choices=['c1', 'c2', 'c3', 'c4', 'c5'], help='BASELINE.json configuration (c4 = headline)')
    ap.add_argument('--batch', type=int, default=None, help='global batch (default: the configuration\'s own: c1 1280, c2 32, c3 64, c4 / c5 256)')
    ap.add_argument('--sweep', action='store_true', help='c5: also time global batches 64 .. 1024 (reported under "sweep")')
    ap.add_argument('--cpu-batch', type=int, default=None, help='units per step of the CPU legs (default: the same global batch as the GPU arm)')
    ap.add_argument('--cpu-steps', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch the step kernel by kernel instead of replaying the captured CUDA graphs')
    ap.add_argument('--layer-report', default=None, help='write a per-layer (per conv shape) timing table to this JSON file')
    args = ap.parse_args()
    if args.batch is None:
        args.batch = DEFAULT_BATCH[args.config]
    rank, world, local_rank = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        return run_reference(args, rank, world)
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        {'c1': run_c1, 'c2': run_c2, 'c3': run_c3, 'c4': run_ours, 'c5': run_ours}[args.config](args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
