"""Bring-up helper: one conv covariance on the tcgen05 engine, error vs fp64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nsgp_repre_b200 as pkg
from oracle import restated as O

Cin, H, W, k, s, p, B = [int(v) for v in sys.argv[1:8]]
g = torch.Generator().manual_seed(0)
x = torch.relu(torch.randn(B, Cin, H, W, generator=g))
conv = torch.nn.Conv2d(Cin, 4, k, stride=s, padding=p, bias=False)
hooks = pkg.CovarianceHooks(torch.nn.Sequential(conv))
hooks._accumulate_conv(x.cuda(), "k", (k, k), (s, s), (p, p))
torch.cuda.synchronize()
got = hooks.fea_in["k"].double().cpu()
want = O.cov_conv2d(x.double(), (k, k), (s, s), (p, p))
print("rel=%.3e" % float(torch.linalg.norm(got - want) / torch.linalg.norm(want)))
