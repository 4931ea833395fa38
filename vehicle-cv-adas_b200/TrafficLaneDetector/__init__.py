from .ufldDetector.ultrafastLaneDetectorV2 import UltrafastLaneDetectorV2
from .ufldDetector.utils import LaneModelType, OffsetType, CurvatureType
