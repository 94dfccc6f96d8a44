"""SURVEY 8(f) rows 2-4 at the BASELINE configs[1] shapes: device time of the kernels next
to the reference-style torch implementation of the same step (CPU for the loops the
reference runs on the host, CUDA torch ops where the reference runs on the device)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nsgp_repre_b200 as pkg
from oracle import restated as O, synth

dev = torch.device("cuda")

def gpu_ms(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def wall_ms(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

# ---------------------------------------------------------------- 8f-2 RoIAlign
B, C, H, W, R = 8, 256, 800, 1344, 8 * 512
feats, rois, labels = synth.roi_case(0, batch=B, channels=C, img_h=H, img_w=W, n_rois=R, classes=19)
cf = [f.to(dev) for f in feats]
crois, clab = rois.to(dev), labels.to(dev)
ext = pkg.SingleRoIExtractor(dict(type="RoIAlign", output_size=7, sampling_ratio=0), C, [4, 8, 16, 32])
t_feat = gpu_ms(lambda: ext(cf, crois))
t_sum = gpu_ms(lambda: ext.class_sums(cf, crois, clab, 19))
from torchvision.ops import roi_align
def tv():
    lv = ext.map_roi_levels(crois, 4)
    out = cf[0].new_zeros(R, C, 7, 7)
    for i in range(4):
        inds = (lv == i).nonzero().squeeze(1)
        if inds.numel():
            out[inds] = roi_align(cf[i], crois[inds], (7, 7), 1.0 / ext.featmap_strides[i], 0, True)
    return out
t_tv = gpu_ms(tv)
def tv_then_sum():
    f = tv().flatten(1)
    return torch.stack([f[clab == c].sum(0) for c in range(19)])
t_tv_sum = gpu_ms(tv_then_sum)
gf = [f.clone().requires_grad_(True) for f in cf]
gout = torch.randn(R, C, 7, 7, device=dev)
def ours_fb():
    for f in gf:
        f.grad = None
    ext(gf, crois).backward(gout)
def tv_fb():
    for f in gf:
        f.grad = None
    lv = ext.map_roi_levels(crois, 4)
    out = gf[0].new_zeros(R, C, 7, 7)
    for i in range(4):
        inds = (lv == i).nonzero().squeeze(1)
        if inds.numel():
            out[inds] = roi_align(gf[i], crois[inds], (7, 7), 1.0 / ext.featmap_strides[i], 0, True)
    out.backward(gout)
t_fb = gpu_ms(ours_fb)
t_tv_fb = gpu_ms(tv_fb)
feat_bytes = sum(f.numel() * 4 for f in cf)
out_bytes = R * C * 49 * 4
print("8f-2 RoIAlign   R=%d C=%d 7x7, 4 levels of a %dx%d batch %d:" % (R, C, H, W, B))
print("   ours features           %.3f ms  (%.0f GB/s of features-read-once + output = %.0f MB)" %
      (t_feat, (feat_bytes + out_bytes) / t_feat / 1e6, (feat_bytes + out_bytes) / 1e6))
print("   ours fused class sums   %.3f ms  (no %.0f MB feature matrix)" % (t_sum, out_bytes / 1e6))
print("   torchvision per level (reference structure, CUDA)        %.3f ms" % t_tv)
print("   torchvision per level + per-class masked sums (CUDA)     %.3f ms" % t_tv_sum)
print("   forward + backward: ours %.3f ms | torchvision per level (CUDA) %.3f ms" % (t_fb, t_tv_fb))

# ---------------------------------------------------------------- 8f-3 pseudo-label merge
gt_b, gt_l, ps_b, ps_s, ps_l = synth.pseudo_label_case(21, images=8, max_gt=8, max_pseudo=100)
cu = lambda ts: [t.to(dev) for t in ts]
g_b, g_l, p_b, p_s, p_l = cu(gt_b), cu(gt_l), cu(ps_b), cu(ps_s), cu(ps_l)
t_merge = wall_ms(lambda: pkg.merge_pseudo_labels(g_b, g_l, p_b, p_s, p_l), reps=10)
t_ref_gpu = wall_ms(lambda: O.pseudo_label_merge(g_b, g_l, p_b, p_s, p_l), reps=2)
t_ref_cpu = wall_ms(lambda: O.pseudo_label_merge(gt_b, gt_l, ps_b, ps_s, ps_l), reps=2)
nps = sum(b.shape[0] for b in ps_b)
print("8f-3 pseudo-label merge, 8 images, %d teacher boxes:" % nps)
print("   ours (1 kernel + 1 D2H, wall)                            %.3f ms" % t_merge)
print("   reference loop on CUDA tensors (one .item() per box)     %.3f ms" % t_ref_gpu)
print("   reference loop on CPU tensors                            %.3f ms" % t_ref_cpu)

# ---------------------------------------------------------------- 8f-4 EWC
import torchvision
net = torchvision.models.resnet50().to(dev)
reg = {n: p for n, p in net.named_parameters() if "bn" in n}
terms = {"importance": {n: [torch.rand_like(p).unsqueeze(0) for _ in range(2)] for n, p in reg.items()},
         "task_param": {n: [(p.detach() + 0.01 * torch.randn_like(p)).unsqueeze(0) for _ in range(2)]
                        for n, p in reg.items()}}
net.loss = lambda: {}
hook = pkg.EWCHook(net, reg, terms, check_nonzero=False)
def ours():
    for p in reg.values():
        p.grad = None
    hook()["ewc_loss"].backward()
def ref():
    for p in reg.values():
        p.grad = None
    loss = 0
    for n, p in reg.items():
        imp = torch.cat(terms["importance"][n], dim=0)
        old = torch.cat(terms["task_param"][n], dim=0)
        new = p.unsqueeze(0).expand(old.shape)
        loss = loss + 1000 * (imp * (new - old) ** 2).sum()
    loss.backward()
t_ours = wall_ms(ours, reps=20)
t_ref = wall_ms(ref, reps=5)
acc = pkg.EWCImportance(reg)
for p in reg.values():
    p.grad = torch.randn_like(p)
t_acc = wall_ms(lambda: acc.accumulate(2, 100), reps=20)
def ref_acc():
    for n, p in acc.importance.items():
        p += (reg[n].grad ** 2) * 2 / 100
t_ref_acc = wall_ms(ref_acc, reps=5)
ne = sum(p.numel() for p in reg.values())
print("8f-4 EWC, ResNet-50 BatchNorm tensors (%d tensors, %d elements, 2 stored tasks):" % (len(reg), ne))
print("   penalty + gradient: ours %.3f ms | reference torch loop (CUDA) %.3f ms   (wall, per step)" % (t_ours, t_ref))
print("   importance update:  ours %.3f ms | reference torch loop (CUDA) %.3f ms" % (t_acc, t_ref_acc))
