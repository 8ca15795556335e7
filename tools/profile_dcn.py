"""Timing of the learned-codec configurations with a per-layer table.
usage: python tools/profile_dcn.py [c3_batch] [c5_batch] [steps]  -> JSON on stdout, tables in gpurun_out/dcn_layers.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neural_imaging_b200 import _lib  # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def layer_table(fn, steps):
    prof = bench.EventProfiler()
    _lib.PROFILER = prof
    for _ in range(steps):
        fn()
    _lib.PROFILER = None
    return prof.layer_table()


def main():
    b3 = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    b5 = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    from neural_imaging_b200.models import compression
    from neural_imaging_b200.workflows.manipulation_classification import ManipulationClassification
    rs = np.random.RandomState(1234)
    out, tables = {}, {}
    # C3: TwitterDCN-32C training step on (b3,128,128,3)
    model = compression.TwitterDCN(patch_size=128, seed=1234)
    x = torch.from_numpy(rs.uniform(size=(b3, 128, 128, 3)).astype(np.float32)).cuda()
    f3 = lambda: model.training_step(x, learning_rate=1e-4)
    ms = timed(f3, steps)
    out['c3'] = {'batch': b3, 'ms_per_step': ms, 'images_per_s': b3 / ms * 1e3, 'tflops_model': 3 * 4.871e9 * b3 / ms / 1e9}
    tables['c3'] = layer_table(f3, 2)
    del model
    # C5: joint step with the learned codec, trainable {fan, nip, dcn}
    dist = {'downsampling': 'pool:2', 'compression': 'dcn', 'compression_params': {'patch_size': 128, 'seed': 1234}}
    flow = ManipulationClassification('UNet', distribution=dist, trainable={'nip', 'dcn'}, raw_patch_size=128, seed=1234)
    xb = torch.from_numpy(rs.uniform(size=(b5, 128, 128, 4)).astype(np.float32)).cuda()
    yb = torch.from_numpy(rs.uniform(size=(b5, 256, 256, 3)).astype(np.float32)).cuda()
    f5 = lambda: flow.training_step_device(xb, yb, 0.1, 0.1, False, 1e-4)
    ms = timed(f5, steps)
    out['c5'] = {'batch': b5, 'ms_per_step': ms, 'patches_per_s': b5 / ms * 1e3}
    tables['c5'] = layer_table(f5, 2)
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/dcn_layers.json', 'w') as f:
        json.dump(tables, f, indent=1)
    print(json.dumps(out))
    for k in ('c3', 'c5'):
        print(k, 'top layers:')
        for r in tables[k][:14]:
            print('  %-52s calls %3d  ms %8.3f  %s' % (r['key'], r['calls'], r['ms'] / 2, ('%.1f TF/s' % r['tflops']) if r['tflops'] else ''))


if __name__ == '__main__':
    main()
