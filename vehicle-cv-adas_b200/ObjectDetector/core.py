"""ObjectDetector/core.py -- RectInfo + ObjectDetectBase with the reference's surface (ObjectDetector/core.py:8-121)."""
from __future__ import annotations

import abc
from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass
class RectInfo:
    x: float
    y: float
    width: float
    height: float
    conf: float
    label: str
    kpss: List[Tuple[int, int]] = field(default_factory=list)

    def tolist(self, dtype=int, format_type: str = "xyxy"):
        if format_type == "xyxy":
            vals = (self.x, self.y, self.x + self.width, self.y + self.height)
        else:
            vals = (self.x, self.y, self.width, self.height)
        return [dtype(v) for v in vals]

    def pad(self, padding: int) -> "RectInfo":
        return RectInfo(self.x - padding, self.y - padding, self.width + 2 * padding, self.height + 2 * padding, self.conf, self.label,
                        self.kpss)


class ObjectDetectBase(abc.ABC):
    _defaults = {"model_path": None, "model_type": None, "classes_path": None, "box_score": None}

    @classmethod
    def set_defaults(cls, config):
        cls._defaults = config

    @classmethod
    def check_defaults(cls):
        return cls._defaults

    @classmethod
    def get_defaults(cls, n):
        return cls._defaults[n] if n in cls._defaults else "Unrecognized attribute name '" + n + "'"

    def __init__(self, logger):
        self.__dict__.update(self._defaults)
        self.logger = logger

    def _warn(self, msg):
        if self.logger is not None:
            (getattr(self.logger, "war", None) or getattr(self.logger, "warning"))(msg)

    @property
    def object_info(self):
        if not hasattr(self, "_object_info"):
            self._object_info = []
            self._warn("Can't get object information, maybe you forget to use detect api.")
        return self._object_info

    def set_input_details(self, engine) -> None:
        if hasattr(engine, "get_engine_input_shape"):
            self.input_shapes = engine.get_engine_input_shape()
            self.input_types = engine.engine_dtype
            self.channes, self.input_height, self.input_width = self.input_shapes[1:]
            if self.logger:
                self.logger.info(f"-> Input Shape : {self.input_shapes}")
                self.logger.info(f"-> Input Type  : {self.input_types}")
        elif self.logger:
            self.logger.error("engine does not adhere to the naming convention of the 'EngineBase' class")

    def set_output_details(self, engine) -> None:
        if hasattr(engine, "get_engine_output_shape"):
            self.output_shapes, self.output_names = engine.get_engine_output_shape()
            if self.logger:
                self.logger.info(f"-> Output Shape : {self.output_shapes}")
        elif self.logger:
            self.logger.error("engine does not adhere to the naming convention of the 'EngineBase' class")

    @staticmethod
    def cornerRect(img, bbox, t=5, rt=1, colorR=(255, 0, 255), colorC=(0, 255, 0)):
        import cv2
        x0, y0, x1, y1 = bbox
        ln = max(1, int(min(y1 - y0, x1 - x0) * 0.2))
        if rt != 0:
            cv2.rectangle(img, (x0, y0), (x1, y1), colorR, rt)
        for cx, cy, sx, sy in ((x0, y0, 1, 1), (x1, y0, -1, 1), (x0, y1, 1, -1), (x1, y1, -1, -1)):
            cv2.line(img, (cx, cy), (cx + sx * ln, cy), colorC, t)
            cv2.line(img, (cx, cy), (cx, cy + sy * ln), colorC, t)
        return img

    @abc.abstractmethod
    def DetectFrame(self):
        return NotImplemented

    @abc.abstractmethod
    def DrawDetectedOnFrame(self):
        return NotImplemented
