"""Hardware probes behind the convolution kernels' design numbers: tcgen05.mma rate and 1-D bulk-copy ingest per SM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_imaging_b200 import _lib
from neural_imaging_b200.tensor import ptr, stream
L = _lib.lib()
cyc = torch.zeros(4096, dtype=torch.int64, device='cuda')
print('== tcgen05.mma kind::tf32 M=128 K=8: cycles per MMA (issue loop | until commit completes), median over CTAs')
for grid in (1, 148):
    for ts in (1, 0):
        for n in (32, 64, 128):
            for nacc in (1, 2):
                rounds = 200
                L.ni_mma_probe(n, ts, rounds, nacc, ptr(cyc), grid, stream())
                torch.cuda.synchronize()
                c = cyc[:2 * grid].view(grid, 2).float().median(dim=0).values / (12 * rounds)
                print('grid %3d  A from %s  N %3d  accumulators %d: issue %.1f  complete %.1f' % (grid, 'TMEM' if ts else 'smem', n, nacc, float(c[0]), float(c[1])))
print('== cp.async.bulk 1-D global->shared: bytes/clk per CTA (one CTA per SM), median over CTAs')
src = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
for same in (1, 0):
    for bytes_ in (8192, 16384, 32768):
        for depth in (1, 2, 4, (6 if bytes_ == 32768 else 8)):
            copies = 256
            for grid in (148,):
                L.ni_bulk_probe(ptr(src), (1 << 21) if same else src.numel(), bytes_, depth, copies, same, ptr(cyc), grid, stream())
                L.ni_bulk_probe(ptr(src), (1 << 21) if same else src.numel(), bytes_, depth, copies, same, ptr(cyc), grid, stream())
                torch.cuda.synchronize()
                c = float(cyc[:grid].float().median())
                print('%s source  copy %5d B  depth %d: %.0f clk/copy, %.1f B/clk/CTA' % ('same' if same else 'own ', bytes_, depth, c / copies, bytes_ * copies / c))
