"""RePRE prototype statistics (a9-a12) on the BASELINE configs' shapes: phase time through
MultiPrototypeReplay.build + staged(), as GB/s of the algorithmic bytes (M*D*4: one read of
the RoI features).  usage: bench_repre.py [reps]   (run under ncu for per-kernel traffic)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import nsgp_repre_b200 as pkg

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
D = 256 * 7 * 7
for name, batch, classes, fg in (("cfg2 VOC 19+1 B=8", 8, 19, 0.25), ("cfg3 VOC 10+10 B=8", 8, 10, 0.25),
                                 ("cfg5 COCO 40+40 B=16", 16, 40, 0.25), ("all-foreground B=8", 8, 19, 1.0)):
    g = torch.Generator().manual_seed(7)
    M = batch * 512
    lab = torch.full((M,), classes, dtype=torch.int64)
    sel = torch.randperm(M, generator=g)[: int(M * fg)]
    lab[sel] = torch.randint(0, classes, (sel.numel(),), generator=g)
    cent = torch.randn(classes + 1, 3, D, generator=g)
    which = torch.randint(0, 3, (M,), generator=g)
    feats = (cent[lab, which] + 0.35 * torch.randn(M, D, generator=g)).cuda()
    lab = lab.cuda()
    proto = pkg.MultiPrototypeReplay(10)
    for _ in range(2):
        proto.build(feats, lab, range(classes)); proto.staged()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        proto.build(feats, lab, range(classes)); proto.staged()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nfg = int((lab < classes).sum())
    print("%-22s M=%d (%d foreground) classes=%d prototypes=%d: %.3f ms per build+gather, %.0f GB/s of M*D*4 (%.0f GB/s of the foreground rows)" %
          (name, M, nfg, classes, proto.bbox_featss.shape[0], ms, M * D * 4 / ms / 1e6, nfg * D * 4 / ms / 1e6))
