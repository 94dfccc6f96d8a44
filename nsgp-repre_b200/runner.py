"""``BRNullSpaceRunner`` drop-in (mmdet/engine/runner/nsrunner_roi_replay.py:111).

The NSGP methods of the reference runner, re-implemented over ``CovarianceHooks`` and
``SGDNSCL`` and packaged as a mixin, so that they bind onto the reference's runner class
when the fork is importable (``registry.register_all`` then registers the subclass under
``RUNNERS['BRNullSpaceRunner']``, the name ``cl_faster_rcnn_cfgs/_base_/brnsrunetime.py:26``
selects) and onto a plain object in the tests:

* ``compute_cov(module, fea_in, fea_out)``      :876-916  forward hook
* ``update_cov(fea_in, k)``                      :923-934
* ``cal_fea_in(train_loader)``                   :704-763  checkpoint -> hooks -> loop ->
  all-reduce -> merge with the previous task -> ``covariance.pth``
* ``update_optim_transforms(train_loader)``      :634-662  ``covariance.pth`` -> eigens ->
  projectors (layer-sharded over the ranks, see ``SGDNSCL.get_eigens``)
* ``update_model_transforms(train_loader)``      :664-692  the reference repeats the same
  build a second time; here it is skipped when the projectors were
  already built from the same file

Everything else of the runner (loops, checkpoints, ``cal_rois``, EWC) stays the class it
is mixed into.  Attributes used: ``model, work_dir, logger, ignore_keys, task_id, offset,
fea_in_load_path, fea_in_save_path, ckpt_keywords, optim_wrapper, load_or_resume()``.
"""
from __future__ import annotations

import os
import os.path as osp
import re

import torch

from .covariance import CovarianceHooks


def _unwrap(model):
    """``model.module`` of a DDP-style wrapper (``is_model_wrapper``, :639-642)."""
    try:
        from mmengine.model import is_model_wrapper
        return model.module if is_model_wrapper(model) else model
    except Exception:                       # mmengine absent: duck-type the wrapper
        from torch.nn.parallel import DataParallel, DistributedDataParallel
        return model.module if isinstance(model, (DataParallel, DistributedDataParallel)) \
            else model


def _rank() -> int:
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class NullSpaceRunnerMixin:
    cov_mode = "deferred"           # CovarianceHooks mode used by cal_fea_in

    # ------------------------------------------------------------------ helpers
    def _nsgp_log(self, msg):
        logger = getattr(self, "logger", None)
        if logger is not None:
            logger.info(msg)

    def _nsgp_check_if_ignore(self, n: str) -> bool:
        """:643-650 - ``re.match`` = prefix match; kept (not ignored) names are logged."""
        ignore = False
        for ignore_key in self.ignore_keys:
            ignore = ignore or bool(re.match(ignore_key, n))
        if not ignore:
            self._nsgp_log("** %s" % n)
        return ignore

    def _nsgp_hooks(self) -> CovarianceHooks:
        hooks = getattr(self, "_nsgp_cov", None)
        model = _unwrap(self.model)
        if hooks is None or hooks.model is not model:
            # self.ignore_keys already carries the reference's defaults (:354)
            hooks = CovarianceHooks(model, ignore_keys=self.ignore_keys,
                                    add_default_ignores=False, mode=self.cov_mode)
            self._nsgp_cov = hooks
        return hooks

    # --------------------------------------------------------- compute_cov / update_cov
    @torch.no_grad()
    def compute_cov(self, module, fea_in, fea_out):
        return self._nsgp_hooks().compute_cov(module, fea_in, fea_out)

    @torch.no_grad()
    def update_cov(self, fea_in, k):
        return self._nsgp_hooks().update_cov(fea_in, k)

    # ------------------------------------------------------------------ cal_fea_in
    @torch.no_grad()
    def cal_fea_in(self, train_loader):
        self._nsgp_log("Doing cal_fea_in......")
        # the checkpoint selected by ``ckpt_keywords`` (:709-714; like the reference, the
        # last directory entry is taken when nothing matches)
        i = None
        for i in os.listdir(self.work_dir):
            if self.ckpt_keywords in i:
                break
        if i is not None:
            self._load_from = osp.join(self.work_dir, i)
        self.load_or_resume()
        model = _unwrap(self.model)
        self._nsgp_cov = None
        hooks = self._nsgp_hooks()
        hooks.reset()
        for n, _ in hooks.hooked_modules():          # the "** name" log of :727-729
            self._nsgp_log("** %s" % n)
        hooks.register()
        model.eval()
        try:
            for data_batch in train_loader:
                data = model.data_preprocessor(data_batch, True)
                model(data["inputs"], data["data_samples"], mode="nullspace")
        finally:
            hooks.remove()
        hooks.all_reduce()                           # barrier / all_reduce_dict / barrier
        if self.task_id != 1:
            self._nsgp_log("During cal_fea_in, trying to load Covariance from %s"
                           % self.fea_in_load_path)
            old_fea_in = torch.load(self.fea_in_load_path,
                                    map_location=next(model.parameters()).device)
            hooks.merge_previous(old_fea_in)
        self._nsgp_log("Trying to save Covariance to %s" % self.fea_in_save_path)
        # every rank holds the same sums after the reduce; one writer is enough (the reference
        # lets every rank write the same file)
        if _rank() == 0:
            hooks.save(self.fea_in_save_path)
        self._nsgp_log("Covariance saved to %s" % self.fea_in_save_path)
        self._nsgp_cov = None                        # ``del self.fea_in`` (:759)

    # ------------------------------------------------------- update_*_transforms
    @torch.no_grad()
    def update_optim_transforms(self, train_loader):
        model = _unwrap(self.model)
        self._nsgp_log("Load Covariance from %s" % self.fea_in_load_path)
        dev = next(model.parameters()).device
        fea_in = torch.load(self.fea_in_load_path, map_location=dev)
        fea_in = {k: v for k, v in fea_in.items() if not self._nsgp_check_if_ignore(k)}
        opt = self.optim_wrapper.optimizer
        opt.get_eigens(fea_in)
        opt.get_transforms(offset=self.offset)
        self._nsgp_transforms_from = (self.fea_in_load_path, float(self.offset), id(opt))
        del fea_in

    @torch.no_grad()
    def update_model_transforms(self, train_loader):
        opt = self.optim_wrapper.optimizer
        if getattr(self, "_nsgp_transforms_from", None) == \
                (self.fea_in_load_path, float(self.offset), id(opt)):
            return                                   # same file, same offset: same projectors
        self.update_optim_transforms(train_loader)
