from .detection_refinedet import Detect_RefineDet, Detections
from .prior_box import PriorBox, REFINEDET_ANCHORS

__all__ = ['PriorBox', 'Detect_RefineDet', 'Detections', 'REFINEDET_ANCHORS']
