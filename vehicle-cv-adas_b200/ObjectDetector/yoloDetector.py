"""ObjectDetector/yoloDetector.py -- YoloDetector with the reference's API, fused on the device.

Reference: ObjectDetector/yoloDetector.py (YoloDetector 52-191).  `DetectFrame` keeps its contract (fills
`object_info` with RectInfo in NMS emission order, `box_score` strict threshold, class-agnostic NMS with
`box_nms_iou`), but __prepare_input (96-102), engine_inference (162), __process_output (104-133),
Scaler.convert_boxes_coordinate (utils.py:70-87) and NMS.fast_soft_nms (utils.py:161-256) execute as one device
pipeline behind `adas_yolo_detect`; only RectInfo construction (135-157) stays in Python.
`DetectFrames(frames)` is the batched extension (per-frame results equal the batch-1 results).
"""
import os
import random

import numpy as np

from ..coreEngine import B200Engine
from .core import ObjectDetectBase, RectInfo
from .utils import ObjectModelType, Scaler, hex_to_rgb


class YoloDetector(ObjectDetectBase):
    _defaults = {
        "model_path": "./models/yolov8l-coco.b200w",
        "model_type": ObjectModelType.YOLOV8,
        "classes_path": "./models/coco_label.txt",
        "box_score": 0.4,
        "box_nms_iou": 0.45,
    }
    MAX_DET = 1024         # output arrays per frame; the library fails loudly if more detections survive the NMS

    def __init__(self, logger=None, **kwargs):
        ObjectDetectBase.__init__(self, logger)
        self.__dict__.update(kwargs)
        self.device = kwargs.get("device", None)
        self.max_batch = int(kwargs.get("max_batch", 1))
        self._initialize_class(self.classes_path)
        self._initialize_model(self.model_path)

    def _initialize_model(self, model_path: str) -> None:
        model_path = os.path.expanduser(model_path)
        if self.logger:
            self.logger.debug("model path: %s." % model_path)
        self.engine = B200Engine(model_path, device=self.device, max_batch=self.max_batch)
        if self.logger:
            self.logger.info(f"YoloDetector Type : [{self.engine.framework_type}] || Version : [{self.engine.providers}]")
        self.set_input_details(self.engine)
        self.set_output_details(self.engine)
        # model_type <-> plan consistency (the reference trusts the user; a wrong pairing silently decodes boxes twice or never)
        kind, meta = self.engine.handle.model_kind, self.engine.handle.meta
        # yoloDetector.py:114-124: YOLOv8 / v9 / v10 share the transposed [4 + nc, anchors] head layout; v5 / v6 / v7 the [anchors, 5 + nc] one
        want_v8 = self.model_type in (ObjectModelType.YOLOV8, ObjectModelType.YOLOV9, ObjectModelType.YOLOV10)
        if want_v8 != (kind == 0):
            raise Exception(f"model_type {self.model_type} does not match the plan (kind {kind})")
        is_lite = kind == 1 and meta[2] != 0
        if (self.model_type == ObjectModelType.YOLOV5_LITE) != is_lite:
            raise Exception(f"model_type {self.model_type} needs a {'lite ' if not is_lite else 'non-lite '}YOLOv5 plan (plan.build_yolov5(lite=...))")
        if self.model_type == ObjectModelType.EfficientDet:
            raise Exception("EfficientDet is a different detector class in the reference (efficientdetDetector.py); not provided here")

    def _initialize_class(self, classes_path) -> None:
        if classes_path is None:      # synthetic runs: COCO-sized anonymous label list
            self.class_names = [f"class{i}" for i in range(80)]
        else:
            classes_path = os.path.expanduser(classes_path)
            assert os.path.isfile(classes_path), Exception("%s is not exist." % classes_path)
            with open(classes_path) as f:
                self.class_names = [c.strip() for c in f.readlines()]
        colors = [hex_to_rgb("#%06x" % random.randint(0, 0xFFFFFF)) for _ in self.class_names]
        self.colors_dict = dict(zip(self.class_names, colors))

    def _label(self, cid: int) -> str:
        try:
            return self.class_names[cid]
        except Exception:
            return "unknown"

    def DetectFrames(self, frames):
        """frames: sequence of HxWx3 uint8 BGR images of one size -> list (per frame) of list[RectInfo]."""
        batch = np.ascontiguousarray(np.stack(frames) if not isinstance(frames, np.ndarray) else frames, dtype=np.uint8)
        out = []
        for s in range(0, batch.shape[0], self.max_batch):
            boxes, scores, cls, _, counts, _ = self.engine.handle.yolo_detect(batch[s:s + self.max_batch], float(self.box_score),
                                                                              float(self.box_nms_iou), self.MAX_DET)
            for b in range(boxes.shape[0]):
                n = int(counts[b])
                out.append([RectInfo(*boxes[b, i], conf=float(scores[b, i]), label=self._label(int(cls[b, i])), kpss=[]) for i in range(n)])
        return out

    def DetectFrame(self, srcimg) -> None:
        self.scaler = Scaler(tuple(self.input_shapes[-2:]), True).set_source(srcimg.shape[0], srcimg.shape[1])
        self._object_info = self.DetectFrames(srcimg[None])[0]

    def DrawDetectedOnFrame(self, frame_show) -> None:
        import cv2
        for info in getattr(self, "_object_info", []):
            x0, y0, x1, y1 = info.tolist()
            color = self.colors_dict.get(info.label, (0, 0, 0))
            self.cornerRect(frame_show, [x0, y0, x1, y1], colorR=color, colorC=color)
            cv2.putText(frame_show, info.label, (x0 + 2, y0 - 7), cv2.FONT_HERSHEY_TRIPLEX, 0.75, (255, 255, 255), 2)
