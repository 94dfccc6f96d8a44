"""B200-native NSGP-RePRE hot path behind the reference's plug-in surface.

Host layer (Python/PyTorch) mirroring the reference classes, over hand-written
sm_100a kernels reached through the C ABI of ``include/nsgp_repre_b200.h``:

* ``covariance.CovarianceHooks``   - BRNullSpaceRunner.compute_cov / update_cov /
  fea_in and the reduce / merge / save tail of cal_fea_in
  (mmdet/engine/runner/nsrunner_roi_replay.py:704-763, 876-934)
* ``optim.SGDNSCL``                - mmdet/engine/optimizers/SGD_NSCL.py:15
* ``prototypes.MultiPrototypeReplay`` / ``StandardMultiPrototypeReplayHead``
  - mmdet/models/roi_heads/standard_roi_replay_head.py:375
* ``rois``                         - all_gather_different_shape + the cal_rois tail
  (mmdet/engine/runner/nsrunner_roi_replay.py:73-105, 815-865)
* ``roi_extract.SingleRoIExtractor`` - multi-level RoIAlign, optionally fused with the
  per-class sums (mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:45-118)
* ``ewc``                          - EWC importance accumulation and penalty
  (mmdet/engine/runner/nsrunner_roi_replay.py:946-1073)
* ``pseudo_labels``                - teacher pseudo-label merge
  (mmdet/models/detectors/faster_rcnn_roi_replay.py:67-108)
* ``runner.NullSpaceRunnerMixin``  - cal_fea_in / update_optim_transforms of BRNullSpaceRunner
  (mmdet/engine/runner/nsrunner_roi_replay.py:634-763)
* ``registry`` / ``mm``            - registers the above into mmengine / mmdet under the
  reference's names (``import nsgp_repre_b200.mm`` or ``registry.register_all()``).

There is no CPU fallback: importing ``_lib`` raises if the CUDA library has not
been built (``python -c "import __graft_entry__ as g; g.build()"``).
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (fails loudly when the .so is missing)
from .covariance import CovarianceHooks, BRNullSpaceCovariance  # noqa: F401
from .optim import SGDNSCL  # noqa: F401
from .prototypes import (MultiPrototypeReplay, StandardMultiPrototypeReplayHead,  # noqa: F401
                         StandardRoIReplayHead, SampledRoIReplay, kmeans_prototypes)
from .runner import NullSpaceRunnerMixin  # noqa: F401
from .rois import all_gather_different_shape, RoIHarvest  # noqa: F401
from .roi_extract import SingleRoIExtractor, reduce_class_sums  # noqa: F401
from .ewc import EWCHook, EWCImportance, register_params  # noqa: F401
from .pseudo_labels import merge_pseudo_labels, merge_into_samples  # noqa: F401
from . import registry  # noqa: F401
