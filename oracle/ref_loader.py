"""TEST INFRASTRUCTURE ONLY.  Loads the reference's OWN hot-path functions,
read-only, from ``/root/reference`` under a stub importer.

The reference package cannot be imported as a whole (``mmengine`` / ``mmcv`` /
``matplotlib`` are not installed and there is no network), but the three files
that carry the hot-path arithmetic only need those packages for decorators,
base classes and logging.  We fabricate empty stand-ins for exactly those roots
and then load the three files by path.  Nothing is copied into this repo.

Available only in the build container: ``/root/reference`` does not exist on the
GPU box, so nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` may call this
module - they use ``oracle.restated`` (pinned against fixtures produced here by
``oracle/make_golden.py``).
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types
from collections import defaultdict
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("NSGP_REFERENCE_ROOT", "/root/reference")

_STUB_ROOTS = ("mmengine", "mmcv", "matplotlib", "mmdet")


def available() -> bool:
    return os.path.isfile(os.path.join(
        REFERENCE_ROOT, "mmdet/engine/optimizers/SGD_NSCL.py"))


class _Registry:
    """Stand-in for an mmengine Registry: ``register_module()`` is identity."""

    def register_module(self, *a, **k):
        def deco(obj):
            return obj
        return deco


class _StubModule(types.ModuleType):
    """Empty package whose attributes are fabricated on demand."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []          # mark as package so sub-imports resolve
        self.__file__ = "<stub %s>" % name

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        if item.isupper():                       # registries: RUNNERS, MODELS ...
            val = _Registry()
        elif item == "is_model_wrapper":
            val = lambda m: False
        elif item == "get_rank":
            val = lambda *a, **k: 0
        elif item == "get_world_size":
            val = lambda *a, **k: 1
        elif item in ("print_log",):
            val = lambda *a, **k: None
        elif item in ("barrier", "all_reduce", "all_reduce_dict", "broadcast"):
            val = lambda *a, **k: None
        else:                                    # base classes / type hints
            val = type(item, (), {})
        setattr(self, item, val)
        return val


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        return None


_finder = None


def _install_stubs():
    global _finder
    if _finder is None:
        for root in _STUB_ROOTS:
            if root in sys.modules and not isinstance(sys.modules[root], _StubModule):
                raise RuntimeError(
                    "real %s is importable; use it instead of the stub loader" % root)
        _finder = _StubFinder()
        sys.meta_path.insert(0, _finder)


def _load_by_path(modname: str, relpath: str, package: str | None = None):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    if package is not None:
        mod.__package__ = package
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_cache: dict = {}


def load_sgd_nscl():
    """The reference ``SGDNSCL`` class (mmdet/engine/optimizers/SGD_NSCL.py:15)."""
    if "sgd" not in _cache:
        _install_stubs()
        mod = _load_by_path("_ref_sgd_nscl", "mmdet/engine/optimizers/SGD_NSCL.py")
        _cache["sgd"] = mod.SGDNSCL
    return _cache["sgd"]


def load_runner():
    """The reference ``BRNullSpaceRunner`` class
    (mmdet/engine/runner/nsrunner_roi_replay.py:111); only its unbound
    ``compute_cov`` / ``update_cov`` are ever called."""
    if "runner" not in _cache:
        _install_stubs()
        mod = _load_by_path("_ref_nsrunner",
                            "mmdet/engine/runner/nsrunner_roi_replay.py")
        _cache["runner"] = mod.BRNullSpaceRunner
    return _cache["runner"]


def load_multi_prototype_head():
    """The reference ``StandardMultiPrototypeReplayHead`` class
    (mmdet/models/roi_heads/standard_roi_replay_head.py:375) with
    ``StandardRoIHead`` replaced by a trivial ``nn.Module``."""
    if "head" not in _cache:
        _install_stubs()
        import torch
        import torch.nn as nn

        pkg = "_ref_models"
        for name in (pkg, pkg + ".roi_heads", pkg + ".task_modules",
                     pkg + ".task_modules.samplers", pkg + ".utils",
                     pkg + ".roi_heads.base_roi_head",
                     pkg + ".roi_heads.standard_roi_head"):
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m

        class StandardRoIHead(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()
                self._dummy = nn.Parameter(torch.zeros(1))
                self.with_shared_head = False

        sys.modules[pkg + ".roi_heads.standard_roi_head"].StandardRoIHead = StandardRoIHead
        sys.modules[pkg + ".roi_heads.base_roi_head"].BaseRoIHead = type("BaseRoIHead", (), {})
        sys.modules[pkg + ".task_modules.samplers"].SamplingResult = type("SamplingResult", (), {})
        sys.modules[pkg + ".utils"].empty_instances = None
        sys.modules[pkg + ".utils"].unpack_gt_instances = None
        mod = _load_by_path(pkg + ".roi_heads.standard_roi_replay_head",
                            "mmdet/models/roi_heads/standard_roi_replay_head.py",
                            package=pkg + ".roi_heads")
        _cache["head"] = mod.StandardMultiPrototypeReplayHead
        _cache["head_mod"] = mod
    return _cache["head"]


# --------------------------------------------------------------------------- #
# Thin drivers around the reference's own functions
# --------------------------------------------------------------------------- #

def ref_covariances(model, batches, mode_call=None):
    """Run the reference ``compute_cov`` / ``update_cov`` as real forward hooks.

    ``model`` is any ``nn.Module``; ``batches`` an iterable of input tensors.
    Returns the reference's ``fea_in`` dict  {"<module path>.weight": (d,d)}.
    Follows the hook registration of nsrunner_roi_replay.py:731-744.
    """
    import torch
    Runner = load_runner()
    fake = SimpleNamespace(model=model, fea_in=defaultdict(dict))
    fake.update_cov = lambda fea, k: Runner.update_cov(fake, fea, k)
    hook = lambda m, i, o: Runner.compute_cov(fake, m, i, o)
    handles = [m.register_forward_hook(hook)
               for n, m in model.named_modules() if hasattr(m, "weight")]
    was_training = model.training
    model.eval()
    with torch.no_grad():
        for x in batches:
            (mode_call or model)(x)
    for h in handles:
        h.remove()
    model.train(was_training)
    return dict(fake.fea_in)


def ref_optimizer(named_params, **kw):
    """Build the reference SGDNSCL with the runner's ``names`` convention
    (nsrunner_roi_replay.py:473-484)."""
    SGDNSCL = load_sgd_nscl()
    names = [n for n, _ in named_params]
    params = [p for _, p in named_params]
    opt = SGDNSCL(params, **kw)
    opt.param_groups[0]["names"] = names
    return opt


class stable_sort_ties:
    """The reference orders neighbour counts with ``sim_mask.sum(-1).sort(dim=-1,
    descending=True)`` (standard_roi_replay_head.py:421) and the counts are small
    integers, so ties are the rule.  Under the reference's pinned torch 1.12
    (README.md:20-22) the CPU sort is a stable sort: equal counts keep ascending row
    order.  torch >= 2.x in this container dispatches the non-stable call to a
    vectorised sort whose tie order is unspecified (and differs between CPUs), which
    would make the greedy picks irreproducible.  This context manager runs the
    reference with its original tie order by making ``Tensor.sort`` stable - nothing
    else about the call changes."""

    def __enter__(self):
        import torch
        self._orig = torch.Tensor.sort
        orig = self._orig

        def sort(t, *args, **kwargs):
            if "stable" in kwargs or (args and isinstance(args[0], bool)):
                return orig(t, *args, **kwargs)
            return orig(t, *args, stable=True, **kwargs)

        torch.Tensor.sort = sort
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.sort = self._orig
        return False


def ref_prototypes(feats, cls_targets, task_split, task_id, max_prototype, tmpdir,
                   saved_masks=None, stable_ties=True):
    """Run the reference StandardMultiPrototypeReplayHead.__init__ prototype build
    (standard_roi_replay_head.py:397-452) on a synthetic rois_etc.pth.

    Returns (bbox_featss (P,D), tmp_label (P,), mask list as saved to mask.pth).
    ``stable_ties``: see ``stable_sort_ties``.
    """
    import torch
    if stable_ties:
        with stable_sort_ties():
            return ref_prototypes(feats, cls_targets, task_split, task_id, max_prototype,
                                  tmpdir, saved_masks, stable_ties=False)
    Head = load_multi_prototype_head()
    prev = os.path.join(tmpdir, "x_%d" % (task_id - 1))
    cur = os.path.join(tmpdir, "x_%d" % task_id)
    os.makedirs(prev, exist_ok=True)
    os.makedirs(cur, exist_ok=True)
    M = feats.shape[0]
    torch.save([feats, cls_targets, torch.ones(M), torch.zeros(M, 4),
                torch.zeros(M, 4), torch.zeros(M, 5)],
               os.path.join(prev, "rois_etc.pth"))
    if saved_masks is not None:
        torch.save(saved_masks, os.path.join(prev, "mask.pth"))
    head = Head(previous_path=prev, task_id=task_id, task_split=list(task_split),
                max_prototype=max_prototype)
    masks = torch.load(os.path.join(cur, "mask.pth"), map_location="cpu")
    return head.bbox_featss, head.tmp_label, masks


# --------------------------------------------------------------------------- #
# EWC importance + penalty (nsrunner_roi_replay.py:946-990,1038-1073)
# --------------------------------------------------------------------------- #

def ref_ewc_importance(model, batches, loss_fn, previous_terms=None):
    """Run the reference ``BRNullSpaceRunner.calculate_save_importance`` (:946-990) and
    ``register_params`` (:1006-1031) as unbound functions on a stand-in runner: the
    model is any ``nn.Module``, ``batches`` a list of "data_batch" dicts (the reference
    scales by ``len(data_batch)`` = number of dict keys, :980), ``loss_fn(model, batch)``
    the scalar loss.  Returns the ``ewc_reg_terms`` dict that the reference pickles."""
    import tempfile
    import torch
    Runner = load_runner()

    class _Logger:
        def info(self, *a, **k):
            pass

    class _OptimWrapper:
        def scale_loss(self, loss):
            return loss

        def backward(self, loss):
            loss.backward()

        def zero_grad(self):
            for p in model.parameters():
                p.grad = None

    model.data_preprocessor = lambda batch, training: batch
    model._run_forward = lambda data, mode: {"loss": loss_fn(model, data)}
    model.parse_losses = lambda losses: (losses["loss"], {})
    fake = SimpleNamespace(model=model, logger=_Logger(), train_dataloader=list(batches),
                           optim_wrapper=_OptimWrapper(),
                           ewc_reg_terms=previous_terms if previous_terms is not None else {},
                           work_dir=tempfile.mkdtemp())
    fake.register_params = lambda: Runner.register_params(fake)
    try:
        Runner.calculate_save_importance(fake, list(batches))
    finally:
        for attr in ("data_preprocessor", "_run_forward", "parse_losses"):
            try:
                delattr(model, attr)
            except AttributeError:
                pass
    return fake.ewc_reg_terms, fake.reg_params


def ref_ewc_hook(module, reg_params, ewc_reg_terms):
    """The reference ``EWCHook`` (:1038-1073) wrapped around ``module.loss``."""
    load_runner()
    mod = sys.modules["_ref_nsrunner"]
    return mod.EWCHook(module=module, reg_params=reg_params, ewc_reg_terms=ewc_reg_terms)


# --------------------------------------------------------------------------- #
# teacher pseudo-label merge (faster_rcnn_roi_replay.py:67-108)
# --------------------------------------------------------------------------- #

class Instances:
    """Minimal stand-in for ``mmengine.structures.InstanceData`` (absent here): tensor
    fields of equal length, ``len``, int / index-tensor ``__getitem__``, ``['field']``,
    ``del``, ``cat``.  Only what the reference's merge loop touches."""

    def __init__(self, **fields):
        object.__setattr__(self, "_f", dict(fields))

    def __getattr__(self, k):
        f = object.__getattribute__(self, "_f")
        if k in f:
            return f[k]
        raise AttributeError(k)

    def __setattr__(self, k, v):
        self._f[k] = v

    def __delattr__(self, k):
        del self._f[k]

    def __len__(self):
        return 0 if not self._f else len(next(iter(self._f.values())))

    def __getitem__(self, item):
        if isinstance(item, str):
            return self._f[item]
        if isinstance(item, int):
            if item >= len(self) or item < -len(self):
                raise IndexError(item)
            item = slice(item, None, len(self)) if item >= 0 else slice(item, None, len(self))
        return Instances(**{k: v[item] for k, v in self._f.items()})

    def keys(self):
        return list(self._f.keys())

    @staticmethod
    def cat(items):
        import torch
        keys = items[0].keys()
        assert all(set(i.keys()) == set(keys) for i in items), "fields differ"
        return Instances(**{k: torch.cat([i._f[k] for i in items]) for k in keys})

    def __deepcopy__(self, memo):
        return Instances(**{k: v.clone() for k, v in self._f.items()})


def load_detector():
    """The reference ``FasterRCNNRoIReplay`` class
    (mmdet/models/detectors/faster_rcnn_roi_replay.py:14) on a trivial base class."""
    if "det" not in _cache:
        _install_stubs()
        import torch.nn as nn
        pkg = "_ref_detectors"
        m = types.ModuleType(pkg)
        m.__path__ = []
        m.TwoStageDetector = type("TwoStageDetector", (nn.Module,), {})
        sys.modules[pkg] = m
        mod = _load_by_path(pkg + ".faster_rcnn_roi_replay",
                            "mmdet/models/detectors/faster_rcnn_roi_replay.py", package=pkg)
        _cache["det"] = mod.FasterRCNNRoIReplay
    return _cache["det"]


def ref_pseudo_label_merge(gt_boxes, gt_labels, ps_boxes, ps_scores, ps_labels,
                           rpn_thresh=0.5, roi_thresh=0.7):
    """Run the reference ``FasterRCNNRoIReplay.loss`` (:40-145) on a stand-in ``self`` whose
    teacher returns the given predictions; the RPN head and RoI head record the data
    samples they are handed.  Returns per image (rpn_boxes, rpn_labels, roi_boxes,
    roi_labels); the RPN labels are recorded before :118-120 zero them."""
    Det = load_detector()
    n = len(gt_boxes)
    samples = [SimpleNamespace(gt_instances=Instances(bboxes=gt_boxes[i].clone(),
                                                      labels=gt_labels[i].clone()))
               for i in range(n)]
    preds = [SimpleNamespace(pred_instances=Instances(bboxes=ps_boxes[i].clone(),
                                                      scores=ps_scores[i].clone(),
                                                      labels=ps_labels[i].clone()))
             for i in range(n)]
    rec = {}

    class _Teacher:
        def eval(self):
            return self

        def predict(self, inputs, data_samples, rescale=False):
            return preds

    class _Rpn:
        def loss_and_predict(self, x, rpn_data_samples, proposal_cfg=None):
            rec["rpn"] = [(s.gt_instances.bboxes.clone(), None) for s in rpn_data_samples]
            return {}, None

    class _Roi:
        def loss(self, x, rpn_results_list, batch_data_samples):
            rec["roi"] = [(s.gt_instances.bboxes.clone(), s.gt_instances.labels.clone())
                          for s in batch_data_samples]
            return {}

    class _Cfg(dict):
        rpn = None

    # :118-120 zero the RPN labels in place: wrap Instances.cat to remember them
    rpn_labels = {}
    fake = SimpleNamespace(extract_feat=lambda x: None, teacher_model=_Teacher(),
                           rpn_thresh=rpn_thresh, roi_thresh=roi_thresh, with_rpn=True,
                           train_cfg=_Cfg(), test_cfg=_Cfg(), rpn_head=_Rpn(), roi_head=_Roi())
    import torch
    orig_zeros_like = torch.zeros_like

    def spy_zeros_like(t, *a, **k):
        rpn_labels.setdefault("seq", []).append(t.clone())
        return orig_zeros_like(t, *a, **k)

    torch.zeros_like = spy_zeros_like
    try:
        Det.loss(fake, None, samples)
    finally:
        torch.zeros_like = orig_zeros_like
    out = []
    for i in range(n):
        out.append((rec["rpn"][i][0], rpn_labels["seq"][i], rec["roi"][i][0], rec["roi"][i][1]))
    return out


# --------------------------------------------------------------------------- #
# SingleRoIExtractor (single_level_roi_extractor.py:45-118)
# --------------------------------------------------------------------------- #

def load_roi_extractor():
    """The reference ``SingleRoIExtractor`` class on a trivial ``BaseRoIExtractor``."""
    if "roi_ext" not in _cache:
        _install_stubs()
        import torch.nn as nn
        pkg = "_ref_roi_extractors"
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
        base = types.ModuleType(pkg + ".base_roi_extractor")
        base.BaseRoIExtractor = type("BaseRoIExtractor", (nn.Module,), {})
        sys.modules[pkg + ".base_roi_extractor"] = base
        mod = _load_by_path(pkg + ".single_level_roi_extractor",
                            "mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py",
                            package=pkg)
        _cache["roi_ext"] = mod.SingleRoIExtractor
    return _cache["roi_ext"]


def ref_roi_extract(feats, rois, featmap_strides=(4, 8, 16, 32), out_channels=256,
                    output_size=7, sampling_ratio=0, finest_scale=56):
    """Run the reference ``SingleRoIExtractor.forward`` and ``map_roi_levels`` (its own level
    mapping, per-level gather and scatter).  ``mmcv.ops.RoIAlign`` is an un-vendored
    dependency that is absent here; its place is taken by ``torchvision.ops.roi_align``
    (same algorithm: aligned=True, average pooling, adaptive sampling grid)."""
    import torch
    import torch.nn as nn
    from torchvision.ops import roi_align
    Ext = load_roi_extractor()

    class _Layer(nn.Module):
        def __init__(self, scale):
            super().__init__()
            self.output_size = (output_size, output_size)
            self.scale = scale

        def forward(self, f, r):
            return roi_align(f, r, self.output_size, self.scale, sampling_ratio, True)

    fake = nn.Module()
    fake.roi_layers = nn.ModuleList([_Layer(1.0 / s) for s in featmap_strides])
    fake.out_channels = out_channels
    fake.finest_scale = finest_scale
    fake.map_roi_levels = lambda r, n: Ext.map_roi_levels(fake, r, n)
    with torch.no_grad():
        out = Ext.forward(fake, tuple(feats), rois)
        lv = Ext.map_roi_levels(fake, rois, len(feats))
    return out, lv


# --------------------------------------------------------------------------- #
# the 5-RoIs-per-batch selection of get_bbox_stuff
# (standard_roi_replay_head.py:106-202)
# --------------------------------------------------------------------------- #

def ref_get_bbox_stuff(bbox_feats, cls_target, cls_weight, bbox_target, bbox_weight, rois,
                       num_classes, seed):
    """Run the reference ``StandardRoIReplayHead.get_bbox_stuff`` on a stand-in ``self`` whose
    assigner / sampler / extractor / bbox head hand back the given tensors; what is
    exercised is the reference's own selection (:165-196), counter (:198-200) and
    gather (:202) under ``torch.manual_seed(seed)``."""
    import torch
    load_multi_prototype_head()
    mod = _cache["head_mod"]
    n_img = 2
    mod.unpack_gt_instances = lambda samples: ([None] * n_img, [None] * n_img, None)
    mod.bbox2roi = lambda priors: rois

    class _RpnResults(dict):
        def pop(self, k):
            return dict.pop(self, k)

    class _Ext:
        num_inputs = 4

        def __call__(self, x, r):
            return bbox_feats

    class _BBoxHead:
        def __init__(self):
            self.num_classes = num_classes

        def get_mid_features(self, f):
            return f

        def get_roi_targets(self, sampling_results, rcnn_train_cfg):
            return cls_target.clone(), cls_weight, bbox_target, bbox_weight

    fake = SimpleNamespace(
        bbox_assigner=SimpleNamespace(assign=lambda *a: None),
        bbox_sampler=SimpleNamespace(sample=lambda *a, **k: SimpleNamespace(priors=None)),
        bbox_roi_extractor=_Ext(), with_shared_head=False, bbox_head=_BBoxHead(),
        train_cfg=None, counter=defaultdict(int))
    x = [torch.zeros(n_img, 1, 1, 1)] * 4
    rpn = [_RpnResults(bboxes=None) for _ in range(n_img)]
    for r in rpn:
        r.priors = None
    torch.manual_seed(seed)
    out = mod.StandardRoIReplayHead.get_bbox_stuff(fake, x, rpn, [None] * n_img)
    return out, dict(fake.counter)
