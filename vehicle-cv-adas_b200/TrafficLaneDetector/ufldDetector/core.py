"""LaneInfo + LaneDetectBase with the reference's surface (TrafficLaneDetector/ufldDetector/core.py:7-180).

The ego-lane polygon (`__update_lanes_area`, core.py:150-158) and the optional degree-2 polyfit resampling
(`__adjust_lanes_points`, core.py:102-141) stay on the host (SURVEY 8f "next" row): <= 144 points per frame.
"""
import abc
from dataclasses import dataclass

import numpy as np


@dataclass
class LaneInfo:
    _lanes_points: np.ndarray
    _lanes_status: np.ndarray
    _area_points: np.ndarray
    _area_status: bool

    @property
    def lanes_points(self):
        return self._lanes_points

    @lanes_points.setter
    def lanes_points(self, arr) -> None:
        if not isinstance(arr, np.ndarray):
            raise Exception("The 'lanes_points' must be np.array[List[Tuple[x, y], ...], ...].")
        self._lanes_points = arr

    @property
    def lanes_status(self):
        return self._lanes_status

    @lanes_status.setter
    def lanes_status(self, value) -> None:
        if any(type(v) != bool for v in value):
            raise Exception("The elements of 'lanes_status' must be of type bool List[bool, ...].")
        self._lanes_status = value

    @property
    def area_status(self):
        return self._area_status

    @area_status.setter
    def area_status(self, value) -> None:
        raise Exception("You need to use the '__update_lanes_status' API to modify it.")

    @property
    def area_points(self):
        return self._area_points

    @area_points.setter
    def area_points(self, value) -> None:
        raise Exception("You need to use the '__update_lanes_area' API to modify it.")


class LaneDetectBase(abc.ABC):
    _defaults = {"model_path": None, "model_type": None}

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
        self.adjust_lanes = False
        empty = lambda: np.array([], dtype=object)
        self.lane_info = LaneInfo(empty(), empty(), empty(), False)

    def set_input_details(self, engine) -> None:
        self.input_shape = engine.get_engine_input_shape()
        self.input_types = engine.engine_dtype
        self.channes, self.input_height, self.input_width = self.input_shape[1:]
        if self.logger:
            self.logger.info(f"-> Input Shape : {self.input_shape}")
            self.logger.info(f"-> Input Type  : {self.input_types}")

    def set_output_details(self, engine) -> None:
        self.output_shape, self.output_names = engine.get_engine_output_shape()
        if self.logger:
            self.logger.info(f"-> Output Shape : {self.output_shape}")

    @staticmethod
    def _adjust_lanes_points(left_pts, right_pts, image_height):
        """Degree-2 x(y) fit of both ego lanes, resampled on a shared y grid (core.py:102-141 semantics)."""
        if len(left_pts) == 0 or len(left_pts[1]) == 0:      # reference indexes [1]: needs >= 2 points
            return left_pts, right_pts
        lx, ly = zip(*left_pts)
        if len(ly) <= 10 or len(right_pts) == 0:
            return left_pts, right_pts
        rx, ry = zip(*right_pts)
        if len(ry) <= 10:
            return left_pts, right_pts
        lfit = np.polyfit(ly, lx, 2)
        rfit = np.polyfit(ry, rx, 2)
        maxy = max(image_height - 1, np.max(ly), np.max(ry))
        miny = min(image_height // 3, np.min(ly), np.min(ry))
        ys = np.linspace(miny, maxy, image_height)
        lxs = lfit[0] * ys ** 2 + lfit[1] * ys + lfit[2]
        rxs = rfit[0] * ys ** 2 + rfit[1] * ys + rfit[2]
        lmin, rmin = min(ly), min(ry)
        new_l = [(int(x), int(y)) for x, y in zip(lxs, ys) if y >= lmin and x >= 0]
        new_r = [(int(x), int(y)) for x, y in zip(rxs, ys) if y >= rmin and x >= 0]
        return new_l, new_r

    def _update_lanes_status(self, lanes_status) -> None:
        self.lane_info._area_status = False
        if lanes_status != [] and len(lanes_status) % 2 == 0:
            mid = len(lanes_status) // 2
            if lanes_status[mid - 1] and lanes_status[mid]:
                self.lane_info._area_status = True

    def _update_lanes_area(self, lanes_points, img_height) -> None:
        self.lane_info._area_points = np.array([], dtype=object)
        if self.lane_info._area_status:
            mid = len(lanes_points) // 2
            left, right = lanes_points[mid - 1], lanes_points[mid]
            if self.adjust_lanes:
                left, right = self._adjust_lanes_points(left, right, img_height)
            self.lane_info._area_points = np.vstack((left, np.flipud(right)))

    # name-mangled aliases used by the reference's call sites (ultrafastLaneDetectorV2.py:192-194)
    _LaneDetectBase__update_lanes_status = _update_lanes_status
    _LaneDetectBase__update_lanes_area = _update_lanes_area

    @abc.abstractmethod
    def DetectFrame(self):
        return NotImplemented

    @abc.abstractmethod
    def DrawDetectedOnFrame(self):
        return NotImplemented

    @abc.abstractmethod
    def DrawAreaOnFrame(self):
        return NotImplemented

    def AutoDrawLanes(self, image, draw_points=True, draw_area=True):
        self.DetectFrame(image, adjust_lanes=True)
        if draw_points:
            self.DrawDetectedOnFrame(image)
        if draw_area:
            self.DrawAreaOnFrame(image)
        return image
