"""Plain-torch Faster R-CNN R50-FPN stand-in whose module paths equal mmdet's
(backbone.layer2.0.conv1, neck.fpn_convs.0.conv, rpn_head.rpn_conv,
roi_head.bbox_head.shared_fcs.0 ...), so synthetic runs produce the reference's
covariance keys and parameter names (mmdet/models/backbones/
resnet.py:154-209, necks/fpn.py:116-204, dense_heads/rpn_head.py:73-99).  It
exists to produce the activations the hooks see - it is NOT part of the hot
path; convolutions run through torch/cuDNN like the reference's."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class Bottleneck(nn.Module):
    def __init__(self, inplanes, planes, stride=1, downsample=False):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = None
        if downsample:
            self.downsample = nn.Sequential(
                nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * 4))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return F.relu(out + idt)


class ResNet50(nn.Module):
    def __init__(self, width=64, blocks=(3, 4, 6, 3)):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        inplanes = width
        for i, n in enumerate(blocks):
            planes = width * 2 ** i
            layers = [Bottleneck(inplanes, planes, 1 if i == 0 else 2, True)]
            inplanes = planes * 4
            layers += [Bottleneck(inplanes, planes) for _ in range(n - 1)]
            setattr(self, "layer%d" % (i + 1), nn.Sequential(*layers))
        self.out_channels = [width * 4 * 2 ** i for i in range(4)]

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        outs = []
        for i in range(4):
            x = getattr(self, "layer%d" % (i + 1))(x)
            outs.append(x)
        return outs


class ConvModule(nn.Module):
    def __init__(self, cin, cout, k, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding)

    def forward(self, x):
        return self.conv(x)


class FPN(nn.Module):
    def __init__(self, in_channels, out_channels=256, num_outs=5):
        super().__init__()
        self.lateral_convs = nn.ModuleList(ConvModule(c, out_channels, 1) for c in in_channels)
        self.fpn_convs = nn.ModuleList(ConvModule(out_channels, out_channels, 3, 1)
                                       for _ in in_channels)
        self.num_outs = num_outs

    def forward(self, feats):
        lat = [l(f) for l, f in zip(self.lateral_convs, feats)]
        for i in range(len(lat) - 1, 0, -1):
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:],
                                                    mode="nearest")
        outs = [c(l) for c, l in zip(self.fpn_convs, lat)]
        while len(outs) < self.num_outs:
            outs.append(F.max_pool2d(outs[-1], 1, stride=2))
        return outs


class RPNHead(nn.Module):
    def __init__(self, ch=256, anchors=3):
        super().__init__()
        self.rpn_conv = nn.Conv2d(ch, ch, 3, padding=1)
        self.rpn_cls = nn.Conv2d(ch, anchors, 1)
        self.rpn_reg = nn.Conv2d(ch, anchors * 4, 1)

    def forward(self, feats):
        outs = []
        for f in feats:                  # shared across the 5 levels -> 5 hook calls
            t = F.relu(self.rpn_conv(f))
            outs.append((self.rpn_cls(t), self.rpn_reg(t)))
        return outs


class Shared2FCBBoxHeadTask(nn.Module):
    """Shape-compatible stand-in of convfc_bbox_head_task.py:516 (one fc_cls over
    all classes + background; the per-task split is out of scope)."""

    def __init__(self, in_dim=256 * 7 * 7, fc=1024, num_classes=20):
        super().__init__()
        self.shared_fcs = nn.ModuleList([nn.Linear(in_dim, fc), nn.Linear(fc, fc)])
        self.fc_cls = nn.Linear(fc, num_classes + 1)
        self.fc_reg = nn.Linear(fc, 4 * num_classes)
        self.num_classes = num_classes

    def forward(self, x):
        x = x.flatten(1)
        for fc in self.shared_fcs:
            x = F.relu(fc(x))
        return self.fc_cls(x), self.fc_reg(x)


class RoIHead(nn.Module):
    def __init__(self, num_classes=20):
        super().__init__()
        self.bbox_head = Shared2FCBBoxHeadTask(num_classes=num_classes)


class FasterRCNNStandIn(nn.Module):
    """backbone + neck (+ rpn_head, + roi_head on synthetic RoI features)."""

    def __init__(self, num_classes=20, with_rpn=True, with_roi=False, width=64,
                 blocks=(3, 4, 6, 3), frozen_stages=1):
        super().__init__()
        self.backbone = ResNet50(width, blocks)
        self.neck = FPN(self.backbone.out_channels, 4 * width)
        self.with_rpn, self.with_roi = with_rpn, with_roi
        if with_rpn:
            self.rpn_head = RPNHead(4 * width)
        if with_roi:
            self.roi_head = RoIHead(num_classes)
        # frozen_stages=1 (resnet.py): stem + layer1 do not train
        if frozen_stages >= 1:
            for m in (self.backbone.conv1, self.backbone.bn1, self.backbone.layer1):
                for p in m.parameters():
                    p.requires_grad = False

    def forward(self, x, roi_feats=None):
        feats = self.neck(self.backbone(x))
        out = {"feats": feats}
        if self.with_rpn:
            out["rpn"] = self.rpn_head(feats)
        if self.with_roi and roi_feats is not None:
            out["roi"] = self.roi_head.bbox_head(roi_feats)
        return out
