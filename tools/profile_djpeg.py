"""Run the fused dJPEG kernels at the BASELINE size (1280 x 128x128x3) for ncu captures / CUDA-event timing.
usage: profile_djpeg.py [n_images] [iters]"""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_imaging_b200 import ops, _lib
from neural_imaging_b200.compression.jpeg_helpers import jpeg_qtable

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
# development library (NI_B200_LIB=.../libni_b200_dev.so): forward variants selected per call through the environment —
# NI_DJPEG_FWD = 3 (generation 3: one tile per CTA, cp.async) | 4 (generation 4: persistent CTAs, TMA-fed stage),
# NI_DJPEG_CTAS = CTAs per SM
variants = ['default']
if os.environ.get('NI_B200_LIB'):
    variants = ['3', '4:4', '3', '4:4', '4:3']
x = torch.rand((n, 128, 128, 3), device='cuda')
dy = torch.randn_like(x)
y, dx = torch.empty_like(x), torch.empty_like(x)
ql, qc = jpeg_qtable(50, 0), jpeg_qtable(50, 1)
flush = torch.empty(256 * 1024 * 1024 // 4, device='cuda')
res = {}
ref = {}
for vi, var in enumerate(variants):
    if var != 'default':
        os.environ['NI_DJPEG_FWD'], os.environ['NI_DJPEG_CTAS'] = var.split(':')[0], (var.split(':') + ['0'])[1]
    for name, fn, bpp, out in (('djpeg_fwd', lambda: ops.djpeg_fwd(x, ql, qc, 'soft', out=y), 24, y), ('djpeg_bwd', lambda: ops.djpeg_bwd(x, dy, ql, qc, 'soft', out=dx), 36, dx)):
        if var not in ('default', variants[0]) and name != 'djpeg_fwd':
            continue                            # the switches steer the forward kernel only
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()                       # L2 flush between timed launches (256 MB > 126 MB L2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        r = {'ms': ms, 'GBps': bpp * n * 128 * 128 / (ms * 1e-3) / 1e9}
        if name in ref:     # agreement between kernel generations on identical inputs
            d = (out - ref[name]).abs()
            r['max_abs_diff_vs_first'] = float(d.max()); r['frac_diff_gt_1e-6'] = float((d > 1e-6).float().mean())
        else:
            ref[name] = out.clone()
        res[name + ('' if var == 'default' else '@' + var + ('#%d' % vi if variants.count(var) > 1 else ''))] = r
print(json.dumps(res))
