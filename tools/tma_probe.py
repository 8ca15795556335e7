import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_imaging_b200 import _lib
from neural_imaging_b200.tensor import ptr, stream
L = _lib.lib()
for (n, h, w, c) in [(256, 128, 128, 32), (256, 32, 32, 128), (256, 16, 16, 256), (256, 64, 64, 64)]:
    x = torch.randn((n, h, w, c), device='cuda')
    for stages in (2, 4, 8):
        for grid_cap in (148, 296):
            cyc = torch.zeros(2048, dtype=torch.int64, device='cuda')
            boxes = 64
            g = L.ni_tma_probe(ptr(x), n, h, w, c, stages, boxes, ptr(cyc), grid_cap, stream())
            torch.cuda.synchronize()
            cy = cyc[:g].float()
            print('n%d %dx%d c%-3d stages %d grid %3d: %.0f cycles/box (median), %.1f B/clk/CTA' % (n, h, w, c, stages, g, float(cy.median()) / boxes, 16384 * boxes / float(cy.median())))
