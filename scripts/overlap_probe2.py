"""Can a big-shared-memory persistent CTA join SMs on which the grouped staging kernel is
already resident?  Launch staging first, then an idle occupier (224 thr, smem KB, 1 ms) on a
second stream; serialised = staging + 1 ms, co-resident = max(staging, 1 ms)."""
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
hooks = pkg.CovarianceHooks(torch.nn.Identity(), mode="deferred")
side = torch.cuda.Stream(priority=-1)
main = torch.cuda.current_stream()

def record():
    for r, x in zip(layers, xs):
        hooks._accumulate_conv(x, r["name"], (r["k"],) * 2, (r["s"],) * 2, (r["p"],) * 2)

def stage_only(stream):
    record()
    js = hooks._sets[hooks._cur]
    del js.jobs[js.pos:]
    js.seen.clear()
    hooks._stage_deferred(js, stream)
    js.pos = 0

def run(smem_kb, occupy_first, occ_ms=3.0):
    stage_only(main); torch.cuda.synchronize()
    record()
    js = hooks._sets[hooks._cur]
    del js.jobs[js.pos:]
    js.seen.clear()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
    occ = lambda: _lib.check(_lib.lib.nsgp_debug_occupy(224, smem_kb << 10, int(occ_ms * 1e-3 * 1.965e9), 148, side.cuda_stream), "occupy")
    if occupy_first:
        occ()
        time.sleep(0.0003)
    s0.record(main)
    hooks._stage_deferred(js, main)
    s1.record(main)
    if not occupy_first:
        occ()
    js.pos = 0
    main.wait_stream(side)
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1), s0.elapsed_time(s1)

for smem in (0, 32, 100, 193, 225):
    a, b = run(smem, False), run(smem, True)
    print("occupier (3 ms) %3d KB: staging first: total %.2f ms, staging %.2f | occupier first: total %.2f ms, staging %.2f" % (smem, a[0], a[1], b[0], b[1]))
