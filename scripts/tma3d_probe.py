"""HBM bandwidth N SMs pull with ONE 3-D tensor-map box per stage (256 floats x rows x B images):
the access pattern of a TMA-fed batch-mean staging kernel.  Bring-up library only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nsgp_repre_b200 as pkg
from nsgp_repre_b200._lib import lib, check
s = torch.cuda.current_stream().cuda_stream
B = 8
img = 64 * 200 * 336 * 16            # 68.8 M floats per image (x 8 images = 2.2 GB)
buf = torch.empty(B * img, dtype=torch.float32, device="cuda").normal_()
for n_ctas in (16, 24, 32):      # 32 CTAs x 2048 boxes = the whole buffer (more would run off it)
    for box_rows, depth in ((4, 6), (4, 4), (2, 12), (2, 8), (1, 12)):
        boxes = 2048
        out = torch.zeros(n_ctas, dtype=torch.int64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        args = (buf.data_ptr(), img, B, box_rows, boxes, depth, out.data_ptr(), n_ctas, s)
        check(lib.nsgp_debug_tma3d_probe(*args), "probe")
        e0.record()
        check(lib.nsgp_debug_tma3d_probe(*args), "probe")
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        gb = n_ctas * boxes * box_rows * 1024 * B / 1e9
        print("ctas=%3d box=(256,%d,%d)=%2d KB depth=%2d (%3d KB in flight/SM): %.2f ms %.0f GB/s total, %.1f GB/s per SM" %
              (n_ctas, box_rows, B, box_rows * B, depth, box_rows * B * depth, ms, gb / ms * 1e3, gb / ms * 1e3 / n_ctas))
