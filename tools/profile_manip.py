"""Run the fused manipulation-stack kernels and the constrained-filter kernels at the BASELINE sizes (256 RGB images of 256x256 -> 1280
pooled images of 128x128) for ncu captures / CUDA-event timing. usage: profile_manip.py [iters]"""
import json
import os
import sys
from collections import OrderedDict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_imaging_b200 import _lib, ops
from neural_imaging_b200.tensor import empty, ptr, stream

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
L = _lib.lib()
B, H = 256, 256
Y = torch.rand((B, H, H, 3), device='cuda')
opsd = OrderedDict([('sharpen', ops.SharpenOp()), ('resample', ops.ResampleOp()), ('gaussian', ops.GaussianOp(5))])
strengths = {'sharpen': 1, 'resample': 50, 'gaussian': 0.83}
stack = ops.PooledStack(opsd)
plan = stack.plan(strengths, H)
c = empty((4 * B, H // 2, H // 2, 3))
dc = torch.randn_like(c)
dY = torch.zeros_like(Y)
x = torch.rand((1280, 128, 128, 3), device='cuda')
nf = torch.randn((5, 5, 3, 3), device='cuda')
r, dr, dnf = torch.empty_like(x), torch.randn_like(x), torch.empty_like(nf)
flush = torch.empty(256 * 1024 * 1024 // 4, device='cuda')
cases = OrderedDict([
    ('manip_stack_pool2_fwd', (lambda: stack.forward(Y, c, plan, strengths, None, training=True), 25.0 * B * H * H, 0)),
    ('manip_stack_pool2_bwd', (lambda: stack.backward(Y, dc, dY, plan, strengths, None), 34.0 * B * H * H, 0)),
    ('cconv5_fwd', (lambda: L.ni_cconv5_fwd(ptr(x), ptr(nf), ptr(r), 1280, 128, 128, stream()), 24.0 * 1280 * 128 * 128, 450.0 * 1280 * 128 * 128)),
    ('cconv5_bwd_data', (lambda: L.ni_cconv5_bwd_data(ptr(dr), ptr(nf), ptr(r), 1280, 128, 128, 0, stream()), 24.0 * 1280 * 128 * 128, 450.0 * 1280 * 128 * 128)),
    ('cconv5_bwd_filter', (lambda: L.ni_cconv5_bwd_filter(ptr(x), ptr(dr), ptr(dnf), 1280, 128, 128, stream()), 24.0 * 1280 * 128 * 128, 450.0 * 1280 * 128 * 128)),
])
res = {}
for name, (fn, nbytes, flop) in cases.items():
    for _ in range(3 if iters > 0 else 0):
        fn()
    if iters == 0:          # ncu capture mode: every kernel exactly once, no timing
        fn()
        continue
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    res[name] = {'ms': ms, 'GBps': nbytes / (ms * 1e-3) / 1e9, 'algorithmic_bytes': nbytes}
    if flop:
        res[name]['TFLOPs'] = flop / (ms * 1e-3) / 1e12
print(json.dumps(res))
