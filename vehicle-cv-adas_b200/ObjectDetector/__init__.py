from .yoloDetector import YoloDetector
from .core import RectInfo
from .utils import ObjectModelType, CollisionType, Scaler
