"""Registers the drop-in classes into the mmengine / mmdet registries when those
packages are importable (they are not vendored by the reference:
requirements/mminstall.txt:1-2).  Names equal the reference's registry entries
(mmdet/registry.py:35,62,75) so ``cl_faster_rcnn_cfgs`` configs resolve unchanged:
``optimizer=dict(type='SGDNSCL', ...)`` (_base_/schedules/schedule_1x_sgdnscl.py:21)
and ``roi_head=dict(type='StandardMultiPrototypeReplayHead', ...)``
(incremental_task/cl_faster_rcnn_nsgp_repre_19_1_2.py).  Without mmengine the
local ``REGISTRY`` dict is the lookup table."""
from __future__ import annotations

REGISTRY = {}
MMENGINE_AVAILABLE = False


def _register_local(name, obj):
    REGISTRY[name] = obj
    return obj


def register_all(force=True):
    """Idempotent; returns the dict of registered names."""
    global MMENGINE_AVAILABLE
    from .optim import SGDNSCL
    from .prototypes import StandardMultiPrototypeReplayHead
    from .covariance import CovarianceHooks
    _register_local("SGDNSCL", SGDNSCL)
    _register_local("StandardMultiPrototypeReplayHead", StandardMultiPrototypeReplayHead)
    _register_local("BRNullSpaceCovariance", CovarianceHooks)
    from .roi_extract import SingleRoIExtractor
    from .ewc import EWCHook
    # local names only: mmdet's own 'SingleRoIExtractor' is replaced explicitly by the
    # integrator, not behind the user's back
    _register_local("SingleRoIExtractor", SingleRoIExtractor)
    _register_local("EWCHook", EWCHook)
    try:
        from mmengine.registry import OPTIMIZERS
        OPTIMIZERS.register_module(name="SGDNSCL", module=SGDNSCL, force=force)
        MMENGINE_AVAILABLE = True
    except Exception:        # mmengine absent: local registry only
        MMENGINE_AVAILABLE = False
    try:
        from mmdet.registry import MODELS
        from mmdet.models.roi_heads import StandardRoIHead
        from .prototypes import MultiPrototypeReplay, get_work_dir
        import os.path as osp
        import torch

        class _MMDetMultiPrototypeReplayHead(StandardRoIHead):
            """The same build/replay logic bound onto mmdet's StandardRoIHead."""

            def __init__(self, *args, previous_path=None, task_id=1, task_split=(0, 10, 20),
                         max_prototype=10, work_dir=None, **kwargs):
                super().__init__(*args, **kwargs)
                self.replay = False
                self.task_split, self.task_id, self.max_proto = list(task_split), task_id, max_prototype
                self._proto = MultiPrototypeReplay(max_prototype)
                if previous_path is not None and osp.exists(previous_path):
                    assert task_id != 1
                    self.replay = True
                    data = torch.load(osp.join(previous_path, "rois_etc.pth"), map_location="cuda")
                    (feats, self.cls_targets, self.cls_weights, self.bbox_targets,
                     self.bbox_weights, self.roiss) = data
                    saved = None
                    if osp.exists(osp.join(previous_path, "mask.pth")):
                        saved = torch.load(osp.join(previous_path, "mask.pth"), map_location="cpu")
                    self._proto.build(feats, self.cls_targets,
                                      range(self.task_split[0], self.task_split[task_id - 1]), saved)
                    self.bbox_featss, self.tmp_label = self._proto.bbox_featss, self._proto.tmp_label
                    torch.save(self._proto.save_idx,
                               osp.join(work_dir or get_work_dir(previous_path), "mask.pth"))

            replay_loss = StandardMultiPrototypeReplayHead.replay_loss

            def loss(self, x, rpn_results_list, batch_data_samples):
                losses = super().loss(x, rpn_results_list, batch_data_samples)
                if self.replay:
                    losses.update(self.replay_loss(self._proto.staged())["replay_loss"])
                return losses

        MODELS.register_module(name="StandardMultiPrototypeReplayHead",
                               module=_MMDetMultiPrototypeReplayHead, force=force)
    except Exception:
        pass
    return dict(REGISTRY)


def build(cfg: dict, **extra):
    """Minimal ``Registry.build`` for the local table: ``dict(type=..., **kw)``."""
    cfg = dict(cfg)
    cls = REGISTRY[cfg.pop("type")]
    cfg.update(extra)
    return cls(**cfg)


register_all()
