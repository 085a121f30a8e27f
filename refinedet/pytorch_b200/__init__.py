"""refinedet.pytorch_b200 — B200-native detect / anchor-matching hot path of RefineDet.

Host-side mirror of the reference's operator interface for this path (same module layout:
``layers.box_utils``, ``layers.functions.detection_refinedet``,
``layers.modules.refinedet_multibox_loss``, ``utils.nms_wrapper``), calling hand-written
sm_100a CUDA kernels through the C ABI in ``include/refinedet_b200.h``.
Importing the package does not load the native library; the first call does, and fails
loudly if it is missing (``python -m refinedet.pytorch_b200.build`` builds it).
"""
from . import _ffi  # noqa: F401
from .layers import box_utils  # noqa: F401
from .layers.functions.detection_refinedet import (Detect_RefineDet, Detections, DetectPlan,  # noqa: F401
                                                         DetectHostPipeline)  # noqa: F401
from .layers.functions.prior_box import PriorBox, REFINEDET_ANCHORS  # noqa: F401
from .layers.modules.refinedet_multibox_loss import RefineDetMultiBoxLoss, RefineDetCriterionPair  # noqa: F401
from .utils import nms_wrapper  # noqa: F401

__all__ = ['Detect_RefineDet', 'Detections', 'DetectPlan', 'DetectHostPipeline', 'RefineDetMultiBoxLoss', 'RefineDetCriterionPair', 'PriorBox', 'REFINEDET_ANCHORS',
           'box_utils', 'nms_wrapper']
