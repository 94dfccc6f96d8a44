"""RoI harvest tail - the producer side of ``rois_etc.pth``.

Mirrors ``all_gather_different_shape`` (mmdet/engine/runner/nsrunner_roi_replay.py:73-105)
and the gather / reserve / merge / save tail of ``BRNullSpaceRunner.cal_rois``
(:815-865).  The reference emulates a variable-length all-gather with 2*W zero-padded
all-reduces per tensor (one pair per rank); here it is one all-gather of the row counts
and one all-gather of the padded tensor - same result list, rank order preserved.
``select_rois`` is the 5-RoIs-per-batch selection of ``get_bbox_stuff``
(mmdet/models/roi_heads/standard_roi_replay_head.py:165-201) with the reference's exact
consumption of torch's global CPU generator, so a seeded run keeps the same RoIs; the RoI
features it selects from come from ``roi_extract.SingleRoIExtractor``.
"""
from __future__ import annotations

import os.path as osp

import torch
import torch.distributed as dist


def all_gather_different_shape(t: torch.Tensor, group=None):
    """list[Tensor]: the tensors of all ranks, each (N_i, ...), N_i may differ (:73-105)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [t]
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts)
    padded = t.new_zeros((n_max,) + tuple(t.shape[1:]))
    padded[:t.shape[0]] = t
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return [p[:c] for p, c in zip(parts, counts)]


@torch.no_grad()
def select_rois(cls_target: torch.Tensor, bg_class_id: int, target_count: int = 5) -> torch.Tensor:
    """Boolean mask over the sampled RoIs of one batch (:165-196): all foreground RoIs
    (label != ``bg_class_id`` = num_classes), topped up with random background RoIs or thinned
    by random removal to exactly ``target_count`` (all RoIs when there are fewer).  The random
    picks are ``torch.randperm(n)[:k]`` on the default CPU generator, as in the reference."""
    keep = cls_target != bg_class_id
    missing = target_count - int(keep.sum().item())
    if missing > 0:
        # too few foreground RoIs: top up with random background ones (:171-183)
        background = torch.where(~keep)[0]
        if len(background) < missing:
            keep[:] = True
        else:
            chosen = torch.randperm(len(background))[:missing]
            keep[background[chosen.to(background.device)]] = True
    elif missing < 0:
        # too many: drop random foreground RoIs (:184-191)
        foreground = torch.where(keep)[0]
        dropped = torch.randperm(len(foreground))[:-missing]
        keep[foreground[dropped.to(foreground.device)]] = False
    return keep


class RoIHarvest:
    """Accumulates the six per-batch tensors of ``mode='roi_replay'`` (:805-813) and
    produces ``rois_etc.pth`` = [feats (M,12544) f32, cls (M,) i64, cls_w (M,) f32,
    bbox_t (M,4), bbox_w (M,4), rois (M,5)] in the reference's list-of-6 format."""

    FIELDS = 6

    def __init__(self):
        self._parts = [[] for _ in range(self.FIELDS)]

    def add(self, bbox_feats, cls_target, cls_weight, bbox_target, bbox_weight, rois):
        for lst, t in zip(self._parts, (bbox_feats, cls_target, cls_weight, bbox_target,
                                        bbox_weight, rois)):
            lst.append(t)

    @torch.no_grad()
    def add_selected(self, bbox_feats, cls_target, cls_weight, bbox_target, bbox_weight, rois,
                     bg_class_id, target_count=5, counter=None):
        """The tail of ``get_bbox_stuff`` (:165-201): select ``target_count`` RoIs of the batch,
        count them per class (``self.counter[c] += 1``, :198-200) and keep the six selected
        tensors.  Returns them like the reference method does."""
        mask = select_rois(cls_target, bg_class_id, target_count)
        idx = mask.nonzero().flatten()                       # one sync instead of six
        picked = tuple(t.index_select(0, idx) for t in (bbox_feats, cls_target, cls_weight,
                                                        bbox_target, bbox_weight, rois))
        if counter is not None:
            for c in picked[1].tolist():
                counter[c] += 1
        self.add(*picked)
        return picked

    @torch.no_grad()
    def finish(self, work_dir=None, previous_dir=None, task_id=1, reserve_per_class=0,
               num_classes=20, generator=None, group=None):
        """Gather over ranks (:815-820), optionally keep ``reserve_per_class`` random rows of
        every class (:825-842; the reference hard-codes ``range(20)`` - ``num_classes``
        here), prepend the previous task's file (:844-856), save (:860-865)."""
        gathered = [torch.cat(all_gather_different_shape(torch.cat(p, dim=0), group))
                    for p in self._parts]
        if reserve_per_class != 0:
            cls_targets = gathered[1]
            picks = {}
            out = []
            for tns in gathered:
                rows = []
                for c in range(num_classes):
                    sel = cls_targets == c
                    if c not in picks:
                        picks[c] = torch.randperm(int(sel.sum()), generator=generator)[
                            :reserve_per_class]
                    rows.append(tns[sel][picks[c].to(tns.device)])
                out.append(torch.cat(rows, dim=0))
            gathered = out
        if task_id != 1:
            old = torch.load(osp.join(previous_dir, "rois_etc.pth"),
                             map_location=gathered[0].device)
            gathered = [torch.cat([o, g], dim=0) for o, g in zip(old, gathered)]
        if work_dir is not None:
            torch.save(gathered, osp.join(work_dir, "rois_etc.pth"))
        return gathered
