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


