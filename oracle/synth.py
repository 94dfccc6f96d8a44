"""TEST INFRASTRUCTURE ONLY.  Seeded synthetic-input builders shared by
``oracle/make_golden.py`` (which feeds them to the reference's own functions) and
the tests (which feed the same inputs to the restatement and to the CUDA path).
Nothing here touches /root/reference."""
from __future__ import annotations

import torch
import torch.nn as nn


class ToyDetector(nn.Module):
    """Every conv geometry of Faster R-CNN R50-FPN at toy size, with mmdet-style
    names so that 'backbone' appears in some keys: 7x7 s2 p3 stem, 1x1, 3x3 s1,
    3x3 s2, 1x1 s2 downsample, FPN 3x3, and a Linear (rank-1 path)."""

    def __init__(self):
        super().__init__()
        self.backbone = nn.Module()
        self.backbone.conv1 = nn.Conv2d(3, 8, 7, stride=2, padding=3, bias=False)
        self.backbone.bn1 = nn.BatchNorm2d(8)
        self.backbone.c1x1 = nn.Conv2d(8, 16, 1, bias=False)
        self.backbone.c3x3 = nn.Conv2d(16, 16, 3, padding=1, bias=False)
        self.backbone.c3x3s2 = nn.Conv2d(16, 16, 3, stride=2, padding=1, bias=False)
        self.backbone.down = nn.Conv2d(16, 32, 1, stride=2, bias=False)
        self.neck = nn.Module()
        self.neck.fpn = nn.Conv2d(32, 8, 3, padding=1)
        self.fc = nn.Linear(8 * 5 * 7, 6)

    def forward(self, x):
        b = self.backbone
        x = torch.relu(b.bn1(b.conv1(x)))
        x = torch.relu(b.c1x1(x))
        y = torch.relu(b.c3x3(x))
        y = torch.relu(b.c3x3s2(y))
        z = torch.relu(b.down(y))
        z = self.neck.fpn(z)
        return self.fc(z.flatten(1))


def toy_batches(seed=0, n=3, B=2):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, 3, 37, 53, generator=g) for _ in range(n)]


def decaying_cov(d, seed, rank=12, top=5.0, ratio=0.93):
    """PSD matrix Q diag(sigma) Q^T with an explicit, well-separated spectrum:
    `rank` shelf values from `top` down to top/2, then a geometric tail starting
    at 1.0 with the given ratio.  Every adjacent gap is large compared with fp32
    SVD noise (1e-7 * top), so the reference's fp32 `torch.svd` projector is
    reproducible to ~1e-5 and a 1e-4 parity bound is meaningful.  (With a flat
    noise bulk at the cut the reference's own P moves by several percent when
    only the BLAS thread count changes - see DESIGN.md, 'projector noise'.)"""
    g = torch.Generator().manual_seed(seed)
    q, _ = torch.linalg.qr(torch.randn(d, d, generator=g, dtype=torch.float64))
    sig = torch.cat([torch.linspace(top, top / 2, rank, dtype=torch.float64),
                     ratio ** torch.arange(d - rank, dtype=torch.float64)])
    cov = (q * sig) @ q.t()
    return ((cov + cov.t()) / 2).float()


def proto_features(seed=0, classes=3, per_class=60, D=12544, sub=5, noise=0.35, bg=30):
    """SURVEY.md 8d RePRE inputs: class = mixture of `sub` sub-centres + noise;
    background rows carry label == classes (skipped by the prototype build)."""
    g = torch.Generator().manual_seed(seed)
    cent = torch.randn(classes, sub, D, generator=g)
    M = classes * per_class + bg
    lab = torch.cat([torch.arange(classes).repeat_interleave(per_class),
                     torch.full((bg,), classes)])
    perm = torch.randperm(M, generator=g)
    lab = lab[perm]
    which = torch.randint(0, sub, (M,), generator=g)
    feats = torch.empty(M, D)
    for i in range(M):
        c = int(lab[i])
        base = cent[c, which[i]] if c < classes else torch.zeros(D)
        feats[i] = base + noise * torch.randn(D, generator=g)
    return feats, lab




def pseudo_label_case(seed=0, images=4, max_gt=6, max_pseudo=40):
    """SURVEY 8(f)-3 inputs: per image a few ground-truth boxes (one image has none) and
    teacher boxes in descending score order like mmdet's predict - some jittered copies of
    the ground truth (IoU above 0.7), some near-duplicates of each other (so the growing
    RoI set matters), scores spread over the 0.5 / 0.7 thresholds."""
    g = torch.Generator().manual_seed(seed)

    def boxes(n):
        xy = torch.rand(n, 2, generator=g) * 300
        wh = torch.rand(n, 2, generator=g) * 120 + 8
        return torch.cat([xy, xy + wh], 1)

    gt_b, gt_l, ps_b, ps_s, ps_l = [], [], [], [], []
    for i in range(images):
        ng = 0 if i == 1 else int(torch.randint(1, max_gt + 1, (1,), generator=g))
        gb = boxes(ng)
        npz = int(torch.randint(max_pseudo // 2, max_pseudo + 1, (1,), generator=g))
        fresh = boxes(npz)
        parts = [fresh]
        if ng:
            parts.append(gb + torch.randn(ng, 4, generator=g) * 2.0)     # overlaps the GT
        parts.append(fresh[: npz // 3] + torch.randn(npz // 3, 4, generator=g) * 3.0)  # near-duplicates
        pb = torch.cat(parts)
        pb = pb[torch.randperm(pb.shape[0], generator=g)]
        sc = torch.rand(pb.shape[0], generator=g).sort(descending=True).values
        sc[::7] = 0.7            # exactly on the RoI threshold (fp32 0.7 > 0.7 is False)
        sc[3::11] = 0.5
        gt_b.append(gb); gt_l.append(torch.randint(0, 20, (ng,), generator=g))
        ps_b.append(pb); ps_s.append(sc)
        ps_l.append(torch.randint(0, 20, (pb.shape[0],), generator=g))
    return gt_b, gt_l, ps_b, ps_s, ps_l


class ToyBNNet(torch.nn.Module):
    """SURVEY 8(f)-4 model: conv/BN stacks whose BatchNorm parameter names contain "bn"
    (the reference registers exactly those, nsrunner_roi_replay.py:1013-1031), one
    "teacher_model" copy that must be ignored, one frozen BN parameter."""

    def __init__(self):
        super().__init__()
        nn = torch.nn
        self.conv1 = nn.Conv2d(3, 8, 3, padding=1)
        self.bn1 = nn.BatchNorm2d(8)
        self.conv2 = nn.Conv2d(8, 16, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(16)
        self.layer = nn.Sequential(nn.Conv2d(16, 1100, 1), nn.BatchNorm2d(1100))   # > 1 chunk
        self.layer_bn_named = nn.BatchNorm2d(1100)
        self.fc = nn.Linear(1100, 4)
        self.teacher_model = nn.Sequential(nn.Conv2d(3, 4, 1), nn.BatchNorm2d(4))
        self.bn2.bias.requires_grad_(False)

    def forward(self, x):
        x = torch.relu(self.bn1(self.conv1(x)))
        x = torch.relu(self.bn2(self.conv2(x)))
        x = self.layer_bn_named(self.layer(x))
        return self.fc(torch.relu(x).mean((2, 3)))

    def loss(self, x, y):
        return {"loss_cls": torch.nn.functional.cross_entropy(self(x), y)}


def ewc_batches(seed=0, n=3):
    g = torch.Generator().manual_seed(seed)
    return [{"inputs": torch.randn(2, 3, 8, 8, generator=g),
             "data_samples": torch.randint(0, 4, (2,), generator=g)} for _ in range(n)]


def roi_case(seed=0, batch=2, channels=16, img_h=256, img_w=320, n_rois=64, classes=5,
             strides=(4, 8, 16, 32)):
    """SURVEY 8(f)-2 inputs: FPN-shaped feature maps and proposals whose sizes span all four
    levels (including boxes larger than the image, a degenerate zero-area box and boxes
    sticking out of the image), labels with background = `classes`."""
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(batch, channels, -(-img_h // s), -(-img_w // s), generator=g) for s in strides]
    side = torch.exp(torch.rand(n_rois, generator=g) * 5.2 + 2.3)        # ~10 .. 1800 px
    ar = torch.exp((torch.rand(n_rois, generator=g) - 0.5) * 1.4)
    w, h = side * ar.sqrt(), side / ar.sqrt()
    cx = torch.rand(n_rois, generator=g) * img_w
    cy = torch.rand(n_rois, generator=g) * img_h
    rois = torch.stack([torch.randint(0, batch, (n_rois,), generator=g).float(),
                        cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    rois[0, 1:] = torch.tensor([10.0, 20.0, 10.0, 20.0])                 # zero area
    rois[1, 1:] = torch.tensor([-50.0, -40.0, 30.0, 60.0])               # sticks out
    rois[2, 1:] = torch.tensor([0.0, 0.0, 112.0, 112.0])                 # scale == 2 * finest
    rois[3, 1:] = torch.tensor([0.0, 0.0, 56.0, 56.0])
    labels = torch.randint(0, classes + 1, (n_rois,), generator=g)
    return feats, rois, labels


def roi_select_case(kind, seed=0, M=64, D=24, num_classes=20):
    """SURVEY 8(f)-1 inputs of the 5-per-batch selection: sampled-RoI tensors of one batch
    with few (``"few"``), many (``"many"``) or no (``"none"``) foreground RoIs, or fewer than
    five RoIs in total (``"tiny"``)."""
    g = torch.Generator().manual_seed(seed)
    if kind == "tiny":
        M = 3
    cls = torch.full((M,), num_classes, dtype=torch.int64)
    nfg = {"few": 2, "many": 17, "none": 0, "tiny": 1}[kind]
    pos = torch.randperm(M, generator=g)[:nfg]
    cls[pos] = torch.randint(0, num_classes, (nfg,), generator=g)
    return (torch.randn(M, D, generator=g), cls, torch.ones(M), torch.randn(M, 4, generator=g),
            torch.ones(M, 4), torch.cat([torch.randint(0, 2, (M, 1), generator=g).float(),
                                         torch.rand(M, 4, generator=g) * 100], 1))
