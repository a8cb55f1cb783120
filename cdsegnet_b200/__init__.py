"""cdsegnet_b200: B200-native (sm_100a) CDSegNet single-step forward hot path.

Importing the package registers the drop-in classes under the reference's keys.
"""
from .registry import MODELS, build_model, register
from .ptv3 import PointTransformerV3, Point
from .segmentor import DefaultSegmentorV2

register("PT-v3m1", PointTransformerV3)
register("DefaultSegmentorV2", DefaultSegmentorV2)

__all__ = ["MODELS", "build_model", "PointTransformerV3", "DefaultSegmentorV2", "Point"]
