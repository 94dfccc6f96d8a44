"""TMA latency / throughput per SM for 64 KB stages of 128-row x 128-byte boxes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nsgp_repre_b200 as pkg
from nsgp_repre_b200._lib import lib, check
s = torch.cuda.current_stream().cuda_stream
rows = 2304
for name, pitch, K in (("pitch 68000 floats (fpn_convs.0 copy: rows 272 KB apart)", 68000, 67200),
                       ("pitch 4200 floats (rows 16.8 KB apart)", 4200, 4192),
                       ("pitch 32 floats (a tile is 16 KB contiguous)", 32, 32)):
    buf = torch.randn(rows * pitch + 64, device="cuda")
    for n_ctas in (1, 148):
        for depth in (1, 2, 3):
            out = torch.zeros(n_ctas, dtype=torch.int64, device="cuda")
            iters = 400
            for _ in range(2):
                check(lib.nsgp_debug_tma_probe(buf.data_ptr(), pitch, K, rows, iters, depth,
                                               out.data_ptr(), n_ctas, s), "probe")
            torch.cuda.synchronize()
            cyc = out.double().mean().item() / iters
            print("%-58s ctas=%3d depth=%d  %.0f cycles per 64 KB stage  (%.1f B/cycle/SM)" %
                  (name, n_ctas, depth, cyc, 65536 / cyc))
    del buf
