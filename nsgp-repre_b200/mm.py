"""``custom_imports = dict(imports=['nsgp_repre_b200.mm'], allow_failed_imports=False)`` in a
``cl_faster_rcnn_cfgs`` config (or ``import nsgp_repre_b200.mm`` in ``tools/train.py``) swaps
the reference's registry entries for the B200 drop-ins:
``OPTIMIZERS['SGDNSCL']``, ``RUNNERS['BRNullSpaceRunner']``,
``MODELS['StandardMultiPrototypeReplayHead' | 'StandardRoIReplayHead']``."""
from .registry import register_all, MM_REGISTERED  # noqa: F401

register_all(force=True, strict=True)
