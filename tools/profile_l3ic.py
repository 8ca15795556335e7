"""Time the l3ic codec kernels (csrc/l3ic.cu) at the BASELINE batch: 1280 images x (16,16,32) latents = 40,960 streams of 256 symbols,
and a 512 x 512 image batch (64 x (64,64,32) = 2,048 streams of 4,096 symbols). CUDA events around the C-ABI calls, device buffers
resident; the reference library (oracle/_ref, one host thread as in the reference's per-image loop) is timed beside it on a bounded
sample. usage: profile_l3ic.py [iters]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from neural_imaging_b200 import _lib
from neural_imaging_b200.tensor import ptr, stream

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
L = _lib.lib()
dev = torch.device('cuda')
book = torch.arange(-15, 17, dtype=torch.float32, device=dev)
res = {}
for tag, (n, h, w, c) in (('b1280_16x16x32', (1280, 16, 16, 32)), ('b64_64x64x32', (64, 64, 64, 32))):
    g = torch.Generator(device='cpu').manual_seed(3)
    z = torch.clamp(torch.round(torch.randn((n, h, w, c), generator=g) * 1.3), -15, 16).to(dev)
    hw = h * w
    slot = -(-hw // 16) * 16
    stride = -(-(5 + 2 * c + c * hw) // 16) * 16
    islot = -(-(hw + 4) // 16) * 16
    idx = torch.empty((n * c * hw,), dtype=torch.uint8, device=dev)
    lb = torch.empty((n * c, slot), dtype=torch.uint8, device=dev)
    ll = torch.empty((n * c,), dtype=torch.int32, device=dev)
    st = torch.empty((n, stride), dtype=torch.uint8, device=dev)
    sl = torch.empty((n,), dtype=torch.int32, device=dev)
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    lo = torch.empty((n * c,), dtype=torch.int32, device=dev)
    ll2 = torch.empty((n * c,), dtype=torch.int32, device=dev)
    idx2 = torch.zeros((n * c, islot), dtype=torch.uint8, device=dev)
    out = torch.empty_like(z)

    def enc():
        L.ni_l3ic_encode(ptr(z), n, h, w, c, ptr(book), 32, ptr(idx), ptr(lb), slot, ptr(ll), ptr(st), stride, ptr(sl), ptr(status), stream())

    def dec():
        L.ni_l3ic_decode(ptr(st), stride, ptr(sl), n, h, w, c, ptr(book), 32, ptr(lo), ptr(ll2), ptr(idx2), islot, ptr(out), ptr(status), stream())
    r = {}
    for name, fn in (('encode', enc), ('decode', dec)):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        r[name] = {'ms': ms, 'streams_per_s': n * c / (ms * 1e-3), 'symbols_per_s': n * c * hw / (ms * 1e-3), 'images_per_s': n / (ms * 1e-3)}
    assert int(status.abs().sum()) == 0 and torch.equal(out, z)
    coded = int(sl.sum())
    r['bits_per_symbol'] = 8.0 * coded / (n * c * hw)
    # the reference coder on the host (FSE_compress + FSE_decompress per layer), bounded sample
    try:
        from oracle import ref_l3ic as R
        lib = R.reference_library()
    except Exception:
        lib = None
    if lib is not None:
        sym = idx.cpu().numpy().reshape(n * c, hw)
        k = min(n * c, 4096)
        t0 = time.perf_counter()
        coded_layers = [R.ref_compress(lib, sym[i].tobytes()) for i in range(k)]
        t1 = time.perf_counter()
        for i, s in enumerate(coded_layers):
            if isinstance(s, bytes):
                R.ref_decompress(lib, s, 4 * hw)
        t2 = time.perf_counter()
        r['cpu_reference'] = {'kind': 'reference', 'cores': 1, 'sample_streams': k, 'encode_streams_per_s': k / (t1 - t0), 'decode_streams_per_s': k / (t2 - t1),
                              'note': 'oracle/_ref FSE library through ctypes, one thread (the reference codes layer by layer on one thread)'}
        r['speedup_encode'] = r['encode']['streams_per_s'] / r['cpu_reference']['encode_streams_per_s']
        r['speedup_decode'] = r['decode']['streams_per_s'] / r['cpu_reference']['decode_streams_per_s']
    res[tag] = r
print(json.dumps(res))
