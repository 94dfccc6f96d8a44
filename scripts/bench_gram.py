"""Micro-benchmark of the covariance call (stage + Gram) on one layer geometry.
usage: bench_gram.py Cin H W k s p [B]   -> per-kernel ms and TFLOP/s (algorithmic, issued)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nsgp_repre_b200 as pkg
from nsgp_repre_b200 import _lib

Cin, H, W, k, s, p = [int(v) for v in sys.argv[1:7]]
B = int(sys.argv[7]) if len(sys.argv) > 7 else 8
x = torch.relu(torch.randn(B, Cin, H, W, device="cuda"))
hooks = pkg.CovarianceHooks(torch.nn.Identity(), mode="immediate")
for _ in range(3):
    hooks._accumulate_conv(x, "k", (k, k), (s, s), (p, p))
torch.cuda.synchronize()
_lib.profile_read()
_lib.profile_enable(True)
reps = 10
for _ in range(reps):
    hooks._accumulate_conv(x, "k", (k, k), (s, s), (p, p))
torch.cuda.synchronize()
_lib.profile_enable(False)
prof = _lib.profile_read()
Hout = (H + 2 * p - k) // s + 1
Wout = (W + 2 * p - k) // s + 1
N, d = Hout * Wout, Cin * k * k
flops = 2.0 * N * d * d
g = prof["gram"][0] / reps
st = prof["stage"][0] / reps
tiles = -(-d // 128)
kpad = Hout * (-(-Wout // 32)) * 32 if k > 1 else -(-N // 32) * 32
issued = tiles * (tiles + 1) / 2 * 3 * 2.0 * 128 * 128 * kpad
print("Cin=%d %dx%d k%d s%d B=%d  N=%d d=%d | gram %.3f ms  %.0f TF alg  %.0f TF issued(128-tiles) | stage %.3f ms (%.0f GB/s in)" %
      (Cin, H, W, k, s, B, N, d, g, flops / g / 1e9, issued / g / 1e9, st, x.numel() * 4 / st / 1e6))
if os.environ.get("NSGP_DBG_COUNTERS"):
    import ctypes
    n = 148 * 8
    buf = (ctypes.c_ulonglong * n)()
    _lib.check(_lib.lib.nsgp_debug_read_counters(buf, n), "counters")
    import numpy as np
    c = np.array(list(buf), dtype=np.float64).reshape(148, 8)
    tot = c[:, 5].mean()
    print("per-CTA mean cycles: total %.0f | MMA warp: wait-operands %.0f (%.0f%%) issue %.0f (%.0f%%) wait-accum %.0f (%.0f%%) | producer: wait-stage %.0f (%.0f%%) issue %.0f (%.0f%%) | K blocks %.0f -> %.0f cycles/K-block" %
          (tot, c[:, 0].mean(), 100 * c[:, 0].mean() / tot, c[:, 1].mean(), 100 * c[:, 1].mean() / tot,
           c[:, 2].mean(), 100 * c[:, 2].mean() / tot, c[:, 3].mean(), 100 * c[:, 3].mean() / tot,
           c[:, 4].mean(), 100 * c[:, 4].mean() / tot, c[:, 6].mean(), tot / max(1, c[:, 6].mean())))
if os.environ.get("NSGP_DBG_COUNTERS"):
    buf = (ctypes.c_ulonglong * n)()
    _lib.check(_lib.lib.nsgp_debug_read_counters(buf, -n), "counters")
    c = np.array(list(buf), dtype=np.float64).reshape(148, 8)
    tot = c[:, 7].mean()
    if tot > 0:
        st = max(1.0, c[:, 4].mean())
        print("AC kernel per-CTA mean cycles: total %.0f, %.0f row steps -> %.0f cycles/step (MMA ideal 2304) | MMA warp: wait-A %.0f%% wait-B %.0f%% wait-epilogue %.0f%% issue %.0f%% | producer: wait-B-slot %.0f%% wait-A-slot %.0f%%" %
              (tot, st, tot / st, 100 * c[:, 0].mean() / tot, 100 * c[:, 1].mean() / tot, 100 * c[:, 2].mean() / tot,
               100 * c[:, 3].mean() / tot, 100 * c[:, 5].mean() / tot, 100 * c[:, 6].mean() / tot))
