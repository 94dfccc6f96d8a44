"""What does a co-resident persistent kernel cost the staging pass?  Stages the 61 layers
(B=8) alone, next to an idle occupier (threads, smem) on a second stream, and next to the
real grouped contraction."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import nsgp_repre_b200 as pkg
from nsgp_repre_b200 import standin, _lib

B = 8
layers = bench.trace_layers(800, 1344, standin)
g = torch.Generator(device="cuda").manual_seed(0)
xs = [torch.relu(torch.randn(B, r["Cin"], r["H"], r["W"], device="cuda", generator=g)) for r in layers]
hooks = pkg.CovarianceHooks(torch.nn.Identity(), mode="grouped")
side = torch.cuda.Stream()

def stage_only():
    for r, x in zip(layers, xs):
        hooks._accumulate_conv(x, r["name"], (r["k"],) * 2, (r["s"],) * 2, (r["p"],) * 2)
    js = hooks._sets[hooks._cur]
    js.pos = 0; js.seen.clear(); js.keep.clear()      # drop the staged set without contracting

def timed(fn, reps=5, occupy=None):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if occupy is not None:
        thr, smem = occupy
        _lib.check(_lib.lib.nsgp_debug_occupy(thr, smem, int(40e-3 * 1.9e9), 148, side.cuda_stream), "occupy")
        time.sleep(0.002)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

print("staging alone                         %.2f ms" % timed(stage_only))
for thr, smem in ((224, 0), (224, 64 << 10), (224, 128 << 10), (224, 193 << 10), (224, 225 << 10), (1024, 0)):
    print("staging + occupier(%4d thr, %3d KB)   %.2f ms" % (thr, smem >> 10, timed(stage_only, occupy=(thr, smem))))
