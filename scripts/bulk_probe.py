"""How much HBM bandwidth can N SMs pull with cp.async.bulk (in-flight bytes in shared memory)?
Bring-up library only:  NSGP_BRINGUP_LIB=1 python scripts/bulk_probe.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nsgp_repre_b200 as pkg
from nsgp_repre_b200._lib import lib, check
s = torch.cuda.current_stream().cuda_stream
per_cta = 48 << 20
buf = torch.empty(148 * per_cta, dtype=torch.uint8, device="cuda")
buf.random_(0, 255)
for n_ctas in (24, 32, 40, 48, 64, 74, 148):
    for chunk, depth in ((8192, 4), (8192, 8), (16384, 8), (16384, 12), (32768, 6)):
        out = torch.zeros(n_ctas, dtype=torch.int64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        check(lib.nsgp_debug_bulk_probe(buf.data_ptr(), per_cta, chunk, depth, out.data_ptr(), n_ctas, s), "probe")
        e0.record()
        check(lib.nsgp_debug_bulk_probe(buf.data_ptr(), per_cta, chunk, depth, out.data_ptr(), n_ctas, s), "probe")
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        gb = n_ctas * per_cta / 1e9
        print("ctas=%3d chunk=%5d depth=%2d (%3d KB in flight/SM): %.2f ms  %.0f GB/s total  %.1f GB/s per SM  %.1f B/clk/SM" %
              (n_ctas, chunk, depth, chunk * depth // 1024, ms, gb / ms * 1e3, gb / ms * 1e3 / n_ctas,
               per_cta / (out.double().mean().item())))
