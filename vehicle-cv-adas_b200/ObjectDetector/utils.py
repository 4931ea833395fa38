"""ObjectDetector/utils.py -- enums and letterbox bookkeeping of the reference's ObjectDetector/utils.py.

`Scaler` (reference utils.py:30-99) only records geometry here: the resize itself (cv2-exact fixed point), the
blob conversion and the inverse box mapping run on the device (csrc/preprocess.cu, csrc/yolo_post.cu).
`NMS.fast_soft_nms` (utils.py:161-256) is the device kernel `yolo_compact_nms_kernel`; there is no host NMS.
"""
from dataclasses import dataclass
from enum import Enum
from typing import Optional, Tuple


class CollisionType(Enum):
    UNKNOWN = "Determined ..."
    NORMAL = "Normal Risk"
    PROMPT = "Prompt Risk"
    WARNING = "Warning Risk"


class ObjectModelType(Enum):
    YOLOV5 = 0
    YOLOV5_LITE = 1
    YOLOV6 = 2
    YOLOV7 = 3
    YOLOV8 = 4
    YOLOV9 = 5
    YOLOV10 = 6
    EfficientDet = 7


def hex_to_rgb(value):
    value = value.lstrip("#")
    n = len(value) // 3
    return tuple(int(value[i:i + n], 16) for i in range(0, len(value), n))


@dataclass
class Scaler:
    """Letterbox geometry, all shapes (H, W).  Same arithmetic as Scaler.process_image (reference utils.py:45-62)."""
    target_size: Tuple[int, int]
    keep_ratio: bool = True
    _new_shape: Optional[Tuple[int, int]] = None
    _old_shape: Optional[Tuple[int, int]] = None
    _pad_shape: Optional[Tuple[int, int]] = None

    def set_source(self, src_h: int, src_w: int) -> "Scaler":
        th, tw = self.target_size
        padh = padw = 0
        newh, neww = th, tw
        if self.keep_ratio and src_h != src_w:
            r = src_h / src_w
            if r > 1:
                neww = int(tw / r)
                padw = int((tw - neww) * 0.5)
            else:
                newh = int(th * r) + 1
                padh = int((th - newh) * 0.5)
        self._old_shape, self._new_shape, self._pad_shape = (src_h, src_w), (newh, neww), (padh, padw)
        return self

    def get_scale_ratio(self):
        if self._old_shape is None or self._new_shape is None:
            raise Exception("Please operate 'process_image' before conversion")
        return self._old_shape[0] / self._new_shape[0], self._old_shape[1] / self._new_shape[1]
