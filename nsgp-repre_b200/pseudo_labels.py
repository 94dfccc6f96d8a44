"""Teacher pseudo-label merge - drop-in for the per-box loop of
``FasterRCNNRoIReplay.loss`` (mmdet/models/detectors/faster_rcnn_roi_replay.py:67-108).

The reference walks every teacher box in Python with a ``box_iou(...).max().item()`` host
sync and up to two ``InstanceData.cat`` per box (hundreds of syncs per training step).
Here ONE kernel (one CTA per image) makes the keep decisions for the whole batch with the
reference's exact arithmetic - torchvision's fp32 ``box_iou``, the sequentially growing RoI
ground-truth set, ``> 0.7`` as a python-float compare, the score thresholds as fp32 compares
- and the host reads the masks back once.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import lib, check, ptr


@torch.no_grad()
def pseudo_label_keep(gt_boxes, ps_boxes, ps_scores, rpn_thresh=0.5, roi_thresh=0.7,
                      iou_thresh=0.7):
    """Per image boolean masks (keep_rpn, keep_roi) over its teacher boxes.
    ``gt_boxes`` / ``ps_boxes``: lists of (G_i,4) / (P_i,4) fp32 xyxy CUDA tensors."""
    n = len(gt_boxes)
    if n == 0:
        return [], []
    dev = ps_boxes[0].device
    for t in list(gt_boxes) + list(ps_boxes):
        _lib.require_cuda(t, "boxes")
    g_len = [int(b.shape[0]) for b in gt_boxes]
    p_len = [int(b.shape[0]) for b in ps_boxes]
    off = lambda lens: torch.tensor([0] + list(torch.tensor(lens).cumsum(0).tolist()),
                                    dtype=torch.int32).to(dev, non_blocking=True)
    g_off, p_off = off(g_len), off(p_len)
    cat = lambda ts, w: (torch.cat([t.reshape(-1, w) for t in ts]).float().contiguous()
                         if sum(t.shape[0] for t in ts) else
                         torch.zeros(1, w, dtype=torch.float32, device=dev))
    gt = cat(gt_boxes, 4)
    ps = cat(ps_boxes, 4)
    sc = torch.cat([s.reshape(-1) for s in ps_scores]).float().contiguous() if sum(p_len) else \
        torch.zeros(1, dtype=torch.float32, device=dev)
    total = max(sum(p_len), 1)
    out = torch.zeros(2 * total, dtype=torch.uint8, device=dev)
    keep_rpn, keep_roi = out[:total], out[total:]
    counts = torch.empty(2 * n, dtype=torch.int32, device=dev)
    check(lib.nsgp_pseudo_label_merge(ptr(gt), ptr(g_off), ptr(ps), ptr(sc), ptr(p_off), n,
                                      max(p_len + [0]), float(rpn_thresh), float(roi_thresh),
                                      float(iou_thresh), ptr(keep_rpn), ptr(keep_roi),
                                      ptr(counts), _lib.current_stream(dev)),
          "nsgp_pseudo_label_merge")
    host = out[:2 * total].cpu().bool()                      # the one sync of the merge
    kr, ko, o = [], [], 0
    for L in p_len:
        kr.append(host[o:o + L])
        ko.append(host[total + o:total + o + L])
        o += L
    return kr, ko


@torch.no_grad()
def merge_pseudo_labels(gt_boxes, gt_labels, ps_boxes, ps_scores, ps_labels, rpn_thresh=0.5,
                        roi_thresh=0.7, iou_thresh=0.7):
    """Tensor-level form of :78-108.  Returns per image
    (rpn_boxes, rpn_labels, roi_boxes, roi_labels): the ground truth followed by the kept
    teacher boxes in their original order."""
    kr, ko = pseudo_label_keep(gt_boxes, ps_boxes, ps_scores, rpn_thresh, roi_thresh, iou_thresh)
    out = []
    for gb, gl, pb, pl, r, o in zip(gt_boxes, gt_labels, ps_boxes, ps_labels, kr, ko):
        ri = r.nonzero().flatten().to(pb.device)
        oi = o.nonzero().flatten().to(pb.device)
        out.append((torch.cat([gb, pb[ri]]), torch.cat([gl, pl[ri]]),
                    torch.cat([gb, pb[oi]]), torch.cat([gl, pl[oi]])))
    return out


@torch.no_grad()
def merge_into_samples(teacher_predictions, batch_data_samples, rpn_data_samples,
                       rpn_thresh=0.5, roi_thresh=0.7, iou_thresh=0.7):
    """Object-level form for ``FasterRCNNRoIReplay.loss``: duck-typed on mmengine's
    ``InstanceData`` (``.bboxes/.scores/.labels``, index-tensor ``__getitem__``, ``del``,
    ``cat``).  Updates ``gt_instances`` of both sample lists in place like :101-106."""
    gt_b = [s.gt_instances.bboxes for s in batch_data_samples]
    ps_b = [t.pred_instances.bboxes for t in teacher_predictions]
    ps_s = [t.pred_instances.scores for t in teacher_predictions]
    kr, ko = pseudo_label_keep(gt_b, ps_b, ps_s, rpn_thresh, roi_thresh, iou_thresh)
    for t, gs, rs, r, o in zip(teacher_predictions, batch_data_samples, rpn_data_samples, kr, ko):
        pred = t.pred_instances
        dev = pred.bboxes.device
        for sample, keep in ((rs, r), (gs, o)):
            idx = keep.nonzero().flatten().to(dev)
            if idx.numel() == 0:
                continue
            sub = pred[idx]
            del sub.scores                                     # :98-99
            sample.gt_instances = sample.gt_instances.cat([sample.gt_instances, sub])
    return batch_data_samples, rpn_data_samples
