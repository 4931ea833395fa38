"""UltrafastLaneDetectorV2 with the reference's API, fused on the device.

Reference: TrafficLaneDetector/ufldDetector/ultrafastLaneDetectorV2.py.  `DetectFrame(image, adjust_lanes)`
keeps its contract (fills lane_info.lanes_points / lanes_status / area_points / area_status); __prepare_input
(96-112), engine_inference (187) and __process_output (114-181) run as one device pipeline behind
`adas_ufld_detect`.  `ModelConfig` (21-55) keeps the dataset geometries; the plan header names its dataset and the model type must agree.
"""
import numpy as np

from ...coreEngine import B200Engine
from .core import LaneDetectBase
from .utils import LaneModelType, OffsetType, lane_colors


class ModelConfig:
    def __init__(self, model_type):
        if model_type == LaneModelType.UFLDV2_TUSIMPLE:
            self.img_w, self.img_h, self.griding_num, self.crop_ratio = 800, 320, 100, 0.8
            self.row_anchor = np.linspace(160, 710, 56) / 720
            self.col_anchor = np.linspace(0, 1, 41)
        elif model_type == LaneModelType.UFLDV2_CURVELANES:
            self.img_w, self.img_h, self.griding_num, self.crop_ratio = 1600, 800, 200, 0.8
            self.row_anchor = np.linspace(0.4, 1, 72)
            self.col_anchor = np.linspace(0, 1, 81)
        else:
            self.img_w, self.img_h, self.griding_num, self.crop_ratio = 1600, 320, 200, 0.6
            self.row_anchor = np.linspace(0.42, 1, 72)
            self.col_anchor = np.linspace(0, 1, 81)
        self.num_lanes = 4


class UltrafastLaneDetectorV2(LaneDetectBase):
    _defaults = {"model_path": "models/culane_res34.b200w", "model_type": LaneModelType.UFLDV2_CULANE}
    LANE_NAMES = ("left-side", "left-ego", "right-ego", "right-side")

    def __init__(self, model_path=None, model_type=None, logger=None, device=None, max_batch=1):
        LaneDetectBase.__init__(self, logger)
        if None not in [model_path, model_type]:
            self.model_path, self.model_type = model_path, model_type
        if self.model_type not in [LaneModelType.UFLDV2_TUSIMPLE, LaneModelType.UFLDV2_CULANE]:
            # same rejection as the reference (ultrafastLaneDetectorV2.py:69-72): CurveLanes has a config but no detector
            if self.logger:
                self.logger.error("UltrafastLaneDetectorV2 can't use %s type." % self.model_type.name)
            raise Exception("UltrafastLaneDetectorV2 can't use %s type." % self.model_type.name)
        self.cfg = ModelConfig(self.model_type)
        self.device, self.max_batch = device, int(max_batch)
        self._initialize_model(self.model_path)

    def _initialize_model(self, model_path: str) -> None:
        if self.logger:
            self.logger.debug("model path: %s." % model_path)
        self.engine = B200Engine(model_path, device=self.device, max_batch=self.max_batch)
        if self.logger:
            self.logger.info(f"UfldDetectorV2 Type : [{self.engine.framework_type}] || Version : [{self.engine.providers}]")
        self.set_input_details(self.engine)
        self.set_output_details(self.engine)
        if len(self.output_names) != 4:
            raise Exception("Output dims is error, please check model. load %d channels not match 4." % len(self.output_names))
        # the plan carries its dataset (header meta[6]: 0 CULane, 1 TuSimple); crop ratio and anchors follow from it inside the library,
        # so a model_type that names the other dataset would silently use the wrong geometry in the reference -- here it is an error
        want = 1 if self.model_type == LaneModelType.UFLDV2_TUSIMPLE else 0
        got = self.engine.handle.meta[6]
        if got != want:
            raise Exception("UltrafastLaneDetectorV2: plan %s was packed for dataset id %d, model_type %s needs %d." % (model_path, got, self.model_type.name, want))

    def DetectFrames(self, frames):
        """Batched extension: list of (lanes_points object-array, lanes_status list[bool]) per frame."""
        batch = np.ascontiguousarray(np.stack(frames) if not isinstance(frames, np.ndarray) else frames, dtype=np.uint8)
        res = []
        for s in range(0, batch.shape[0], self.max_batch):
            pts, npts, status, _ = self.engine.handle.ufld_detect(batch[s:s + self.max_batch])
            for b in range(pts.shape[0]):
                lanes = [[(int(x), int(y)) for x, y in pts[b, l, :npts[b, l]]] for l in range(4)]
                arr = np.empty(4, dtype=object)
                for l in range(4):
                    arr[l] = lanes[l]
                res.append((arr, [bool(v) for v in status[b]]))
        return res

    def DetectFrame(self, image, adjust_lanes: bool = True) -> None:
        self.img_height, self.img_width, self.img_channels = image.shape
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
