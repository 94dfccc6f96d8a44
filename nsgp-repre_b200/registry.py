"""Registers the drop-in classes into the mmengine / mmdet registries (the reference's
plug-in surface, mmdet/registry.py:35,62,75), under the reference's own names, so that
``cl_faster_rcnn_cfgs`` configs resolve unchanged:

* ``optimizer=dict(type='SGDNSCL', ...)``       _base_/schedules/schedule_1x_sgdnscl.py:21
* ``runner_type = "BRNullSpaceRunner"``         _base_/brnsrunetime.py:26
* ``roi_head=dict(type='StandardMultiPrototypeReplayHead', ...)``
                                                incremental_task/cl_faster_rcnn_nsgp_repre_19_1_2.py:65
* ``StandardRoIReplayHead``                     mmdet/models/roi_heads/standard_roi_replay_head.py:31

Importing the package fills the local ``REGISTRY`` table only.  Replacing the entries of
the mm registries is an explicit step: ``register_all(force=True)``, or - from a config,
without touching the fork - ``custom_imports = dict(imports=['nsgp_repre_b200.mm'])``.
A registration that fails is logged and re-raised with ``strict=True``; it is never
swallowed."""
from __future__ import annotations

import logging

REGISTRY = {}
MMENGINE_AVAILABLE = False
MM_REGISTERED = {}          # registry name -> [entry names] actually replaced
_log = logging.getLogger("nsgp_repre_b200.registry")


def _register_local(name, obj):
    REGISTRY[name] = obj
    return obj


def register_local():
    from .optim import SGDNSCL
    from .prototypes import StandardMultiPrototypeReplayHead, StandardRoIReplayHead
    from .covariance import CovarianceHooks
    from .runner import NullSpaceRunnerMixin
    from .roi_extract import SingleRoIExtractor
    from .ewc import EWCHook
    _register_local("SGDNSCL", SGDNSCL)
    _register_local("StandardMultiPrototypeReplayHead", StandardMultiPrototypeReplayHead)
    _register_local("StandardRoIReplayHead", StandardRoIReplayHead)
    _register_local("BRNullSpaceCovariance", CovarianceHooks)
    _register_local("BRNullSpaceRunner", NullSpaceRunnerMixin)
    # local names only: mmdet's own 'SingleRoIExtractor' is replaced explicitly by the
    # integrator, not behind the user's back
    _register_local("SingleRoIExtractor", SingleRoIExtractor)
    _register_local("EWCHook", EWCHook)
    return dict(REGISTRY)


def _mm_heads(ref_sampled_head):
    """The two replay heads bound onto the reference's ``StandardRoIReplayHead`` (so that
    ``get_bbox_stuff`` / ``counter`` of ``mode='roi_replay'`` and the whole StandardRoIHead
    machinery stay the fork's own code); only the replay paths change."""
    from .prototypes import ReplayHeadMixin

    class StandardRoIReplayHead(ReplayHeadMixin, ref_sampled_head):
        def __init__(self, *args, previous_path=None, **kwargs):
            # the fork's constructor would torch.load the six tensors onto the CPU and index
            # them there every step; it is given no path and the device store is built here
            super().__init__(*args, **kwargs)
            dev = next((p.device for p in self.parameters()), None)
            self.init_sampled_replay(previous_path, dev)

        replay_loss = ReplayHeadMixin.teacher_replay_loss

        def loss(self, x, rpn_results_list, batch_data_samples, replay=True):
            # the grandparent's plain RoI-head loss; the fork's own replay branch is replaced
            losses = super().loss(x, rpn_results_list, batch_data_samples, replay=False)
            if self.replay and replay:
                losses.update(self.sampled_replay_losses())
            return losses

    class StandardMultiPrototypeReplayHead(ReplayHeadMixin, ref_sampled_head):
        def __init__(self, *args, previous_path=None, task_id=1, task_split=(0, 10, 20),
                     max_prototype=10, work_dir=None, **kwargs):
            super().__init__(*args, **kwargs)
            dev = next((p.device for p in self.parameters()), None)
            self.init_prototype_replay(previous_path, task_id, task_split, max_prototype, dev)

        def loss(self, x, rpn_results_list, batch_data_samples):
            losses = super().loss(x, rpn_results_list, batch_data_samples, replay=False)
            if self.replay:
                losses.update(self.prototype_replay_losses())
            return losses

    return StandardRoIReplayHead, StandardMultiPrototypeReplayHead


def register_all(force=True, strict=False):
    """Local table + (when importable) the mm registries.  Returns the local table."""
    global MMENGINE_AVAILABLE
    register_local()
    MM_REGISTERED.clear()

    def attempt(what, fn):
        try:
            fn()
            return True
        except ImportError as e:
            _log.info("nsgp_repre_b200: %s not registered (%s)", what, e)
        except Exception:
            _log.exception("nsgp_repre_b200: registering %s failed", what)
            if strict:
                raise
        return False

    def reg_optimizer():
        from mmengine.registry import OPTIMIZERS
        OPTIMIZERS.register_module(name="SGDNSCL", module=REGISTRY["SGDNSCL"], force=force)
        MM_REGISTERED.setdefault("OPTIMIZERS", []).append("SGDNSCL")

    def reg_heads():
        from mmdet.registry import MODELS
        from mmdet.models.roi_heads.standard_roi_replay_head import \
            StandardRoIReplayHead as ref_head
        sampled, multi = _mm_heads(ref_head)
        MODELS.register_module(name="StandardRoIReplayHead", module=sampled, force=force)
        MODELS.register_module(name="StandardMultiPrototypeReplayHead", module=multi,
                               force=force)
        MM_REGISTERED.setdefault("MODELS", []).extend(
            ["StandardRoIReplayHead", "StandardMultiPrototypeReplayHead"])

    def reg_runner():
        from mmdet.registry import RUNNERS
        from mmdet.engine.runner.nsrunner_roi_replay import BRNullSpaceRunner as ref_runner
        from .runner import NullSpaceRunnerMixin

        class BRNullSpaceRunner(NullSpaceRunnerMixin, ref_runner):
            pass

        RUNNERS.register_module(name="BRNullSpaceRunner", module=BRNullSpaceRunner, force=force)
        REGISTRY["BRNullSpaceRunner"] = BRNullSpaceRunner
        MM_REGISTERED.setdefault("RUNNERS", []).append("BRNullSpaceRunner")

    MMENGINE_AVAILABLE = attempt("OPTIMIZERS['SGDNSCL']", reg_optimizer)
    attempt("MODELS replay heads", reg_heads)
    attempt("RUNNERS['BRNullSpaceRunner']", reg_runner)
    return dict(REGISTRY)


def build(cfg: dict, **extra):
    """Minimal ``Registry.build`` for the local table: ``dict(type=..., **kw)``."""
    cfg = dict(cfg)
    cls = REGISTRY[cfg.pop("type")]
    cfg.update(extra)
    return cls(**cfg)


register_local()
