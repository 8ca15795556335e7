"""Development probe (library built with NI_BUILD_TAG=dev): time fprop / dgrad of the N <= 64 layers with parts of the persistent gemm
switched off (NI_TC_EXP bits: 1 converters skip LDS + split, 2 skip tcgen05.st, 4 issuers skip the MMAs) to see which stage sets the
iteration period. Results of the switched-off variants are wrong by construction; only their times matter."""
import ctypes, os, subprocess, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import numpy as np, torch
    from neural_imaging_b200 import _lib, nn
    from neural_imaging_b200.tensor import empty, ptr, stream
    L = _lib.lib(); rs = np.random.RandomState(0)
    out = {}
    for (n, h, w, cin, cout, k) in [(256, 128, 128, 32, 32, 3), (1280, 64, 64, 32, 64, 5), (256, 64, 64, 64, 64, 3), (256, 32, 32, 128, 128, 3)]:
        st = nn.ParamStore(); conv = nn.Conv2D(st, 'c', k, cin, cout, activation='leaky_relu', rng=rs); st.finalize()
        d = conv.desc(n, h, w)
        x = torch.randn((n, h, w, cin), device='cuda'); dy = torch.randn((n, h, w, cout), device='cuda')
        y, dx = empty((n, h, w, cout)), empty((n, h, w, cin))
        for name, fn in (('fprop', lambda: L.ni_conv2d_fprop_tc(ctypes.byref(d), ptr(x), ptr(conv.w.value), ptr(conv.b.value), ptr(y), stream())),
                         ('dgrad', lambda: L.ni_conv2d_dgrad_tc(ctypes.byref(d), ptr(dy), ptr(conv.w.value), ptr(dx), stream()))):
            for _ in range(3): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): fn()
            e1.record(); e1.synchronize()
            out['%s %dx%d c%d->%d k%d' % (name, h, w, cin, cout, k)] = round(e0.elapsed_time(e1) / 5, 4)
    print(json.dumps(out))
else:
    res = {}
    for exp in (0, 1, 8, 16, 32, 33, 36, 39):
        env = dict(os.environ, NI_TC_EXP=str(exp))
        r = subprocess.run([sys.executable, __file__, 'child'], env=env, capture_output=True, text=True)
        try:
            res[exp] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            res[exp] = {'error': (r.stderr or r.stdout)[-400:]}
    keys = list(res[0].keys())
    print('%-32s' % 'layer' + ''.join('  exp=%2d' % e for e in res))
    for k in keys:
        print('%-32s' % k + ''.join('  %6.3f ' % res[e].get(k, float('nan')) if 'error' not in res[e] else '   error' for e in res))
    for e in res:
        if 'error' in res[e]: print(e, res[e]['error'])
