"""Bring-up helper: one Gram on the tcgen05 engine, error vs fp64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nsgp_repre_b200 as pkg

N, d = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator().manual_seed(0)
rows = torch.randn(N, d, generator=g)
hooks = pkg.CovarianceHooks(torch.nn.Identity())
hooks.update_cov(rows.cuda(), "k")
torch.cuda.synchronize()
got = hooks.fea_in["k"].double().cpu()
want = rows.double().t() @ rows.double()
print("N=%d d=%d rel=%.3e" % (N, d, float(torch.linalg.norm(got - want) / torch.linalg.norm(want))))
