from .ufldDetector.ultrafastLaneDetectorV2 import UltrafastLaneDetectorV2
from .ufldDetector.ultrafastLaneDetector import UltrafastLaneDetector
from .ufldDetector.utils import LaneModelType, OffsetType, CurvatureType
