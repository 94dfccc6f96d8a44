"""Raw tcgen05 MMA rate on smem-resident operands (csrc/mma_rate.cu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nsgp_repre_b200 as pkg
from nsgp_repre_b200._lib import lib, check
names = {0: "3xTF32 pattern 128x128x8 (12 MMAs/block)", 1: "hi*hi only (4 MMAs/block)",
         2: "3xTF32 pattern 128x256x8", 3: "bf16 128x128x16 (12 MMAs/block)"}
s = torch.cuda.current_stream().cuda_stream
for n_ctas in (1, 148):
    for mode in (0, 1, 2, 3):
        out = torch.zeros(n_ctas, dtype=torch.int64, device="cuda")
        iters = 2000
        for _ in range(2):
            check(lib.nsgp_debug_mma_rate(mode, iters, out.data_ptr(), n_ctas, s), "rate")
        torch.cuda.synchronize()
        cyc = out.double().mean().item() / iters
        nm = 4 if mode == 1 else 12
        print("ctas=%3d  %-45s %.1f cycles/K-block  %.1f cycles/MMA" % (n_ctas, names[mode], cyc, cyc / nm))
