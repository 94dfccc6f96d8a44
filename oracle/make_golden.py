"""TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/*.pt`` by running the
REFERENCE'S OWN functions (loaded read-only from /root/reference by
``oracle.ref_loader``) on seeded synthetic inputs.

Run in the build container only:   python -m oracle.make_golden

The reference ships no golden vectors for this path, so these
files are what pins ``oracle/restated.py`` (tests/test_oracle_golden.py) and,
through it, the CUDA path.  Inputs are stored in the fixture when small, or
regenerated from the recorded seed (with a checksum) when large.
"""
from __future__ import annotations

import os
import sys
import tempfile
import warnings

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as R  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden")


from oracle.synth import (ToyDetector, toy_batches, decaying_cov,  # noqa: E402
                          proto_features)


def main():
    assert R.available(), "reference tree not found"
    os.makedirs(OUT, exist_ok=True)
    warnings.filterwarnings("ignore")

    # 1. covariance hooks -----------------------------------------------------
    torch.manual_seed(0)
    net = ToyDetector()
    batches = toy_batches()
    fea = R.ref_covariances(net, batches)
    torch.save({"state_dict": net.state_dict(), "seed": 0,
                "batch_checksum": float(sum(b.double().sum() for b in batches)),
                "fea_in": {k: v for k, v in fea.items() if isinstance(v, torch.Tensor)}},
               os.path.join(OUT, "cov_toy.pt"))
    print("cov keys:", {k: tuple(v.shape) for k, v in fea.items() if isinstance(v, torch.Tensor)})

    # 2. projector build ------------------------------------------------------
    covs = {"backbone.a.weight": decaying_cov(72, 1, rank=6),      # d<128 branch
            "neck.b.weight": decaying_cov(256, 2),
            "backbone.c.weight": decaying_cov(288, 3, rank=20),
            "roi.d.weight": decaying_cov(130, 4, rank=4)}
    proj = {}
    for offset in (0.0, 0.5, -0.5, 3.0):
        params = [(n, nn.Parameter(torch.zeros(4, v.shape[0]))) for n, v in covs.items()]
        opt = R.ref_optimizer(params, lr=0.02, svd=True)
        opt.get_eigens(covs)
        opt.get_transforms(offset=offset)
        entry = {}
        for n in covs:
            sv = opt.eigens[n]["eigen_value"]
            mask = opt.adaptive_threshold(sv, offset=offset)
            entry[n] = {"svals": sv.clone(), "i_thres": int(mask.long().argmax()),
                        "transform": opt.transforms[n].clone()}
        proj[offset] = entry
    torch.save({"proj": proj, "seeds": {"backbone.a.weight": (72, 1, 6),
                                         "neck.b.weight": (256, 2, 12),
                                         "backbone.c.weight": (288, 3, 20),
                                         "roi.d.weight": (130, 4, 4)}},
               os.path.join(OUT, "projector.pt"))
    print("i_thres:", {o: {n: e["i_thres"] for n, e in ent.items()} for o, ent in proj.items()})

    # 3. optimizer steps --------------------------------------------------------
    g = torch.Generator().manual_seed(5)
    shapes = {"backbone.c.weight": (24, 32, 3, 3), "neck.b.weight": (16, 256, 1, 1),
              "roi.d.weight": (10, 130), "backbone.bn.weight": (24,), "head.bias": (10,)}
    init = {n: torch.randn(*s, generator=g) for n, s in shapes.items()}
    grads = [{n: torch.randn(*s, generator=g) for n, s in shapes.items()} for _ in range(2)]
    steps = {}
    for tag, kw in {"mom_wd": dict(lr=0.02, momentum=0.9, weight_decay=1e-4),
                    "plain": dict(lr=0.1),
                    "nesterov_damp": dict(lr=0.05, momentum=0.8, dampening=0.1,
                                          nesterov=True, weight_decay=1e-3)}.items():
        params = [(n, nn.Parameter(v.clone())) for n, v in init.items()]
        opt = R.ref_optimizer(params, svd=True, **kw)
        opt.get_eigens(covs)
        opt.get_transforms(offset=0.0)
        traj = []
        for gr in grads:
            for n, p in params:
                p.grad = gr[n].clone()
            opt.step()
            traj.append({"w": {n: p.detach().clone() for n, p in params},
                         "grad_after": {n: p.grad.clone() for n, p in params},
                         "buf": {n: opt.state[p]["previous_grad"].clone() for n, p in params}})
        steps[tag] = {"kw": kw, "traj": traj}
    torch.save({"init": init, "grads": grads, "steps": steps},
               os.path.join(OUT, "sgd_steps.pt"))

    # 4. prototypes ---------------------------------------------------------------
    feats, lab = proto_features()
    with tempfile.TemporaryDirectory() as td:
        protos, tmp_label, masks = R.ref_prototypes(feats, lab, [0, 3, 4], 2, 10, td)
    # second task: replay saved masks for classes 0..2, then add class 3
    feats2, lab2 = proto_features(seed=1, classes=4, per_class=60, bg=20)
    feats2 = torch.cat([feats, feats2[lab2 == 3]])
    lab2 = torch.cat([lab, lab2[lab2 == 3]])
    with tempfile.TemporaryDirectory() as td:
        protos2, tmp_label2, masks2 = R.ref_prototypes(
            feats2, lab2, [0, 3, 4, 5], 3, 10, td, saved_masks=[list(m) for m in masks])
    torch.save({"feat_checksum": float(feats.double().sum()),
                "labels": lab, "protos": protos, "tmp_label": tmp_label,
                "masks": masks, "labels2": lab2, "protos2": protos2,
                "tmp_label2": tmp_label2, "masks2": masks2,
                "feat2_checksum": float(feats2.double().sum())},
               os.path.join(OUT, "prototypes.pt"))
    print("protos", tuple(protos.shape), tmp_label.tolist(), [len(m) for m in masks])
    print("protos2", tuple(protos2.shape), tmp_label2.tolist(), [len(m) for m in masks2])
    # 5. teacher pseudo-label merge: the reference's FasterRCNNRoIReplay.loss
    from oracle.synth import pseudo_label_case, ToyBNNet, ewc_batches
    cases = []
    for seed in (0, 1, 2):
        gt_b, gt_l, ps_b, ps_s, ps_l = pseudo_label_case(seed)
        out = R.ref_pseudo_label_merge(gt_b, gt_l, ps_b, ps_s, ps_l)
        cases.append({"seed": seed, "out": out})
        print("pseudo merge", seed, [(tuple(o[0].shape), tuple(o[2].shape)) for o in out])
    torch.save(cases, os.path.join(OUT, "pseudo_merge.pt"))

    # 6. EWC: calculate_save_importance over two "tasks", then EWCHook
    torch.manual_seed(0)
    net = ToyBNNet()
    state0 = {k: v.clone() for k, v in net.state_dict().items()}
    loss_fn = lambda m, b: nn.functional.cross_entropy(m(b["inputs"]), b["data_samples"])
    terms, reg = R.ref_ewc_importance(net, ewc_batches(0), loss_fn)
    with torch.no_grad():
        for k, p in enumerate(reg.values()):
            p.add_(0.05 * torch.randn(p.shape, generator=torch.Generator().manual_seed(100 + k)))
    terms, reg = R.ref_ewc_importance(net, ewc_batches(1), loss_fn, previous_terms=terms)
    state2 = {k: v.clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        for k, p in enumerate(reg.values()):
            p.add_(0.02 * torch.randn(p.shape, generator=torch.Generator().manual_seed(200 + k)))
    state3 = {k: v.clone() for k, v in net.state_dict().items()}
    hook = R.ref_ewc_hook(net, reg, terms)
    b = ewc_batches(2)[0]
    net.train()
    res = hook(b["inputs"], b["data_samples"])
    for p in net.parameters():
        p.grad = None
    res["ewc_loss"].backward()
    torch.save({"state0": state0, "state2": state2, "state3": state3,
                "terms": {k: {n: [t.clone() for t in v] for n, v in d.items()}
                          for k, d in terms.items()},
                "reg_names": list(reg.keys()), "ewc_loss": res["ewc_loss"].detach(),
                "grads": {n: p.grad.clone() for n, p in reg.items() if p.grad is not None}},
               os.path.join(OUT, "ewc.pt"))
    print("ewc", list(reg.keys()), float(res["ewc_loss"]))
    # 7. RoI extraction: the reference's SingleRoIExtractor.forward / map_roi_levels
    from oracle.synth import roi_case
    feats, rois, labels = roi_case(0)
    out, lv = R.ref_roi_extract(feats, rois, out_channels=feats[0].shape[1])
    torch.save({"seed": 0, "out": out, "levels": lv}, os.path.join(OUT, "roi_extract.pt"))
    print("roi extract", tuple(out.shape), torch.bincount(lv, minlength=4).tolist())
    # 8. the 5-RoIs-per-batch selection: the reference's get_bbox_stuff
    from oracle.synth import roi_select_case
    sel = []
    for kind, seed in (("few", 11), ("many", 12), ("none", 13), ("tiny", 14)):
        out, counter = R.ref_get_bbox_stuff(*roi_select_case(kind, seed), 20, seed)
        sel.append({"kind": kind, "seed": seed, "out": [o.clone() for o in out],
                    "n_counted": sum(counter.values())})
        print("roi select", kind, out[1].tolist())
    torch.save(sel, os.path.join(OUT, "roi_select.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
