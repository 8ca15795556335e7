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
print('== cp.async.bulk 1-D global->shared: cycles per copy and bytes/clk per SM (one CTA per SM), median over CTAs')
src = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
for same in (1, 0):
    for bytes_, depth, warps in ((8192, 4, 1), (16384, 4, 1), (32768, 4, 1), (4096, 4, 1), (4096, 4, 2), (4096, 4, 4), (8192, 4, 2), (8192, 4, 4), (16384, 2, 2), (16384, 2, 4),
                                 (32768, 2, 2), (32768, 1, 4), (65536, 2, 1)):
        copies, grid = 256, 148
        for _ in range(2):
            L.ni_bulk_probe(ptr(src), (1 << 21) if same else src.numel(), bytes_, depth, copies, same, warps, ptr(cyc), grid, stream())
        torch.cuda.synchronize()
        c = float(cyc[:grid].float().median())
        print('%s source  copy %5d B  depth %d  issuing warps %d: %.0f clk per copy (per warp), %.1f B/clk/SM' % (
            'same' if same else 'own ', bytes_, depth, warps, c / copies, bytes_ * copies * warps / c))
