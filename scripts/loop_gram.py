"""Run the grouped covariance contraction back to back for a few seconds (no staging in
between) and sample SM clock / power with nvidia-smi: is the kernel power-capped?"""
import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import nsgp_repre_b200 as pkg
from nsgp_repre_b200 import standin, _lib
import ctypes

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
layers = bench.trace_layers(800, 1344, standin)
g = torch.Generator(device="cuda").manual_seed(0)
hooks = pkg.CovarianceHooks(torch.nn.Identity(), mode="grouped")
for r in layers:
    x = torch.relu(torch.randn(1, r["Cin"], r["H"], r["W"], device="cuda", generator=g))
    hooks._accumulate_conv(x, r["name"], (r["k"],) * 2, (r["s"],) * 2, (r["p"],) * 2)
hooks.join()
js = hooks._sets[0]
s = torch.cuda.current_stream().cuda_stream
torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu",
                      "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0
t0 = time.time()
e0.record()
while time.time() - t0 < secs:
    for _ in range(10):
        _lib.check(_lib.lib.nsgp_group_launch(js.table.data_ptr(), ctypes.byref(js.group), s), "launch")
    n += 10
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
time.sleep(0.2)
p.terminate()
out = p.stdout.read().strip().splitlines()
print("launches %d, %.3f ms each" % (n, e0.elapsed_time(e1) / n))
print("clock/power samples (MHz, W, power_cap, hw_slowdown, sw_thermal, temp):")
for l in out[::3]:
    print("  ", l)
