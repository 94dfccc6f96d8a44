"""Per-layer staging time (nsgp_cov_conv2d_stage) at R50-FPN 800x1344, B=8."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes
import bench
import nsgp_repre_b200 as pkg
from nsgp_repre_b200 import standin, _lib
from nsgp_repre_b200._lib import lib, CovLayout, ptr, check

B = 8
layers = bench.trace_layers(800, 1344, standin)
s = torch.cuda.current_stream().cuda_stream
rows = []
for r in layers:
    x = torch.relu(torch.randn(B, r["Cin"], r["H"], r["W"], device="cuda"))
    L = CovLayout()
    geom = (r["Cin"], r["H"], r["W"], r["k"], r["k"], r["s"], r["s"], r["p"], r["p"])
    check(lib.nsgp_cov_conv2d_layout(*geom, L), "layout")
    ws = torch.empty(L.workspace_bytes, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        check(lib.nsgp_cov_conv2d_stage(ptr(x), B, *geom, ptr(ws), ws.numel(), s), "stage")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        check(lib.nsgp_cov_conv2d_stage(ptr(x), B, *geom, ptr(ws), ws.numel(), s), "stage")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    rows.append((ms, r["name"], r["Cin"], r["H"], r["W"], r["k"], r["s"], x.numel() * 4 / 1e6, L.kind))
    del x, ws
tot = sum(r[0] for r in rows)
print("total %.3f ms" % tot)
for ms, name, C, H, W, k, st, mb, kind in sorted(rows, reverse=True)[:28]:
    print("%-36s C=%4d %3dx%-4d k%d s%d  in %6.1f MB  %.3f ms  %5.0f GB/s in  kind=%d" % (name, C, H, W, k, st, mb, ms, mb / ms, kind))
