"""UltrafastLaneDetector (UFLD v1) with the reference's API, fused on the device.

Reference: TrafficLaneDetector/ufldDetector/ultrafastLaneDetector.py.  `DetectFrame(image, adjust_lanes)` keeps its contract (fills
lane_info.lanes_points / lanes_status / area_points / area_status); __prepare_input (80-95: whole frame resized to 800x288, mean / std),
engine_inference (142) and __process_output (97-136: grid softmax expectation per row anchor) run as one device pipeline behind
`adas_ufld_detect` on a UFLD v1 plan (`plan.build_ufldv1`).  `ModelConfig` (15-37) keeps the two dataset geometries; the plan header
names its dataset and the model type must agree.

Note on the reference: under numpy 2 its `DetectFrame` raises inside `LaneDetectBase.__update_lanes_status` (the object array
`lanes_detected` no longer compares with `[]`, core.py:145); here `lanes_status` is a list of bools, as the v2 detector produces.
"""
import numpy as np

from ...coreEngine import B200Engine
from .core import LaneDetectBase
from .utils import LaneModelType, OffsetType, lane_colors


class ModelConfig:
    def __init__(self, model_type):
        if model_type == LaneModelType.UFLD_TUSIMPLE:
            self.img_w, self.img_h, self.griding_num, self.cls_num_per_lane = 1280, 720, 100, 56
            self.row_anchor = np.linspace(64, 284, self.cls_num_per_lane)
        else:
            self.img_w, self.img_h, self.griding_num, self.cls_num_per_lane = 1640, 590, 200, 18
            self.row_anchor = [round(value) for value in np.linspace(121, 287, self.cls_num_per_lane)]
        self.num_lanes = 4


class UltrafastLaneDetector(LaneDetectBase):
    _defaults = {"model_path": "models/tusimple_18.b200w", "model_type": LaneModelType.UFLD_TUSIMPLE}

    def __init__(self, model_path=None, model_type=None, logger=None, device=None, max_batch=1):
        LaneDetectBase.__init__(self, logger)
        if None not in [model_path, model_type]:
            self.model_path, self.model_type = model_path, model_type
        if self.model_type not in [LaneModelType.UFLD_TUSIMPLE, LaneModelType.UFLD_CULANE]:
            if self.logger:
                self.logger.error("UltrafastLaneDetector can't use %s type." % self.model_type.name)
            raise Exception("UltrafastLaneDetector can't use %s type." % self.model_type.name)
        self.cfg = ModelConfig(self.model_type)
        self.device, self.max_batch = device, int(max_batch)
        self._initialize_model(self.model_path)

    def _initialize_model(self, model_path: str) -> None:
        if self.logger:
            self.logger.debug("model path: %s." % model_path)
        self.engine = B200Engine(model_path, device=self.device, max_batch=self.max_batch)
        if self.logger:
            self.logger.info(f"UfldDetector Type : [{self.engine.framework_type}] || Version : {self.engine.providers}")
        self.set_input_details(self.engine)
        self.set_output_details(self.engine)
        if len(self.output_names) != 1:
            raise Exception("Output dims is error, please check model. load %d channels not match 1." % len(self.output_names))
        want = 1 if self.model_type == LaneModelType.UFLD_TUSIMPLE else 0
        if self.engine.handle.model_kind != 4 or self.engine.handle.meta[6] != want:
            raise Exception("UltrafastLaneDetector: plan %s is not a UFLD v1 plan of dataset id %d (model_type %s)." % (model_path, want, self.model_type.name))

    def DetectFrames(self, frames):
        """Batched extension: list of (lanes_points object-array, lanes_status list[bool]) per frame."""
        batch = np.ascontiguousarray(np.stack(frames) if not isinstance(frames, np.ndarray) else frames, dtype=np.uint8)
        res = []
        for s in range(0, batch.shape[0], self.max_batch):
            pts, npts, status, _ = self.engine.handle.ufld_detect(batch[s:s + self.max_batch])
            for b in range(pts.shape[0]):
                arr = np.empty(4, dtype=object)
                for l in range(4):
                    arr[l] = [[int(x), int(y)] for x, y in pts[b, l, :npts[b, l]]]
                res.append((arr, [bool(v) for v in status[b]]))
        return res

    def DetectFrame(self, image, adjust_lanes: bool = True) -> None:
        self.img_height, self.img_width, self.img_channels = image.shape
        self.h_ratio, self.w_ratio = image.shape[0] / self.cfg.img_h, image.shape[1] / self.cfg.img_w
        pts, status = self.DetectFrames(image[None])[0]
        self.lane_info.lanes_points, self.lane_info.lanes_status = pts, status
        self.adjust_lanes = adjust_lanes
        self._update_lanes_status(self.lane_info.lanes_status)
        self._update_lanes_area(self.lane_info.lanes_points, self.img_height)

    def DrawDetectedOnFrame(self, image, type: OffsetType = OffsetType.UNKNOWN, alpha: float = 0.3) -> None:
        import cv2
        for lane_num, pts in enumerate(self.lane_info.lanes_points):
            for p in pts:
                cv2.circle(image, (int(p[0]), int(p[1])), 3, lane_colors[lane_num], -1)

    def DrawAreaOnFrame(self, image, color=(255, 191, 0), alpha: float = 0.85) -> None:
        import cv2
        if self.lane_info.area_status and len(self.lane_info.area_points):
            overlay = image.copy()
            cv2.fillPoly(overlay, pts=[np.asarray(self.lane_info.area_points, dtype=np.int32)], color=color)
            image[:] = cv2.addWeighted(image, alpha, overlay, 1 - alpha, 0)
