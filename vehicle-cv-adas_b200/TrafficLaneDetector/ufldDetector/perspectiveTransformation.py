"""Bird-view geometry of the lane pipeline (host side): same class, methods and results as
TrafficLaneDetector/ufldDetector/perspectiveTransformation.py of the reference.

What the reference computes (file:line of the reference):
  * `__init__` (20-37): source trapezoid at (0.3W,0.7H) (0.2W,H) (0.95W,H) (0.8W,0.7H), destination rectangle inset W/4 on both
    sides, `M` / `M_inv` from `cv2.getPerspectiveTransform` (float32 corner arrays);
  * `updateTransformParams` (39-87): re-derives the trapezoid from the ego-lane points ("Top" / "Bottom" / "Default") and
    refreshes both matrices;
  * `transformToBirdView` / `transformToFrontalView` (90-117): `cv2.warpPerspective` with `M` / `M_inv`;
  * `transformToBirdViewPoints` (120-142): homogeneous transform in float64, perspective divide, truncation to int;
  * `calcCurveAndOffset` (145-214): degree-2 `np.polyfit` x(y) per ego lane, direction from the dominant quadratic term
    (+-1.5e-4) and the lane's first vs middle point, refit in metres (30/720 m per px in y, 3.7/700 in x), mean radius of
    curvature at the image bottom, lane-centre offset at row 719 scaled by 3.7 m / lane width; draws two arrows and two text lines.

This file is the per-frame host tail of the lane path (a few hundred flops per frame: 4-point homography, two 3x3 least-squares
fits); it stays on the host like SURVEY row K, the heavy parts (`warpPerspective` of a 2.76 MB frame, drawing) are OpenCV calls as
in the reference.  The numeric core is separated from the drawing so that it can be tested against golden vectors produced by the
reference class (tests/golden/make_golden.py -> birdview.npz).
"""
from typing import Optional, Tuple, Union

import cv2
import numpy as np

from .utils import OffsetType, lane_colors

YM_PER_PIX = 30 / 720      # metres per pixel along y (reference :187)
XM_PER_PIX = 3.7 / 700     # metres per pixel along x (reference :188)


def curve_and_offset(left_lanes, right_lanes, img_h: int, img_w: int):
    """Numeric core of `calcCurveAndOffset` (reference :157-205): returns (direction, curvature_m, offset_m, veh_pos, cen_pos, y_eval)."""
    left = np.squeeze(np.asarray(left_lanes))
    right = np.squeeze(np.asarray(right_lanes))
    lfit = np.polyfit(left[:, 1], left[:, 0], 2)
    rfit = np.polyfit(right[:, 1], right[:, 0], 2)
    lead = lfit[0] if abs(lfit[0]) > abs(rfit[0]) else rfit[0]
    if lead < -0.00015 and left[0, 0] <= left[int(len(left) / 2), 0]:
        direction = "L"
    elif lead > 0.00015 and right[0, 0] >= right[int(len(right) / 2), 0]:
        direction = "R"
    else:
        direction = "F"
    ys = np.linspace(0, img_h - 1, img_h)
    lx = lfit[0] * ys ** 2 + lfit[1] * ys + lfit[2]
    rx = rfit[0] * ys ** 2 + rfit[1] * ys + rfit[2]
    y_eval = np.max(ys)
    radii = []
    for xs in (lx, rx):
        a, b, _ = np.polyfit(ys * YM_PER_PIX, xs * XM_PER_PIX, 2)
        radii.append(((1 + (2 * a * y_eval * YM_PER_PIX + b) ** 2) ** 1.5) / np.absolute(2 * a))
    curvature = (radii[0] + radii[1]) / 2
    # the reference evaluates the lane width at row 719 whatever the image height (:199-202)
    lane_width = np.absolute(lx[719] - rx[719])
    veh_pos = (lx[719] + rx[719]) / 2.0
    cen_pos = img_w / 2.0
    offset = (veh_pos - cen_pos) * (3.7 / lane_width)
    return direction, curvature, offset, veh_pos, cen_pos, y_eval


class PerspectiveTransformation(object):
    """Transforms images and points between the frontal view and the bird view (reference class of the same name)."""

    def __init__(self, img_size=(1280, 720), logger=None):
        self.img_size = img_size
        self.logger = logger
        w, h = self.img_size
        self.src = np.float32([(w * 0.3, h * 0.7), (w * 0.2, h), (w * 0.95, h), (w * 0.8, h * 0.7)])      # tl, bl, br, tr
        ox, oy = w / 4, 0
        self.dst = np.float32([(ox, oy), (ox, h - oy), (w - ox, h - oy), (w - ox, oy)])
        self._refresh()

    def _refresh(self) -> None:
        self.M = cv2.getPerspectiveTransform(self.src, self.dst)
        self.M_inv = cv2.getPerspectiveTransform(self.dst, self.src)

    def updateTransformParams(self, left_lanes: Union[list, np.ndarray], right_lanes: Union[list, np.ndarray], type: str = "Default") -> None:
        left = left_lanes if isinstance(left_lanes, list) else left_lanes.tolist()
        right = right_lanes if isinstance(right_lanes, list) else right_lanes.tolist()
        if not (len(left) and len(right)) or type not in ("Top", "Bottom", "Default"):
            return
        left, right = np.squeeze(left), np.squeeze(right)
        (tl, bl, br, tr) = [tuple(p) for p in self.src]
        top_y = min(min(left[:, 1]), min(right[:, 1]))
        if type == "Top":
            corners = [(max(left[:, 0]) - 20, top_y), (bl[0] - 10, bl[1]), (br[0] + 10, br[1]), (min(right[:, 0]) + 20, top_y)]
        elif type == "Bottom":
            corners = [tl, (min(left[:, 0]) - 20, bl[1]), (max(right[:, 0]) + 20, br[1]), tr]
        else:
            corners = [(max(left[:, 0]) - 20, top_y), (min(left[:, 0]) - 5, bl[1]), (max(right[:, 0]) + 5, br[1]), (min(right[:, 0]) + 20, top_y)]
        if self.logger is not None:
            self.logger.debug("Transform Type : " + type)
            for name, c in zip(("top-left", "bottom-left", "bottom-right", "top-right"), corners):
                self.logger.debug("\t%s :%s" % (name, str(c)))
        self.src = np.float32(corners)
        self._refresh()

    def transformToBirdView(self, img, flags=cv2.INTER_LINEAR):
        return cv2.warpPerspective(img, self.M, self.img_size, flags=flags)

    def transformToFrontalView(self, img, flags=cv2.INTER_LINEAR):
        return cv2.warpPerspective(img, self.M_inv, self.img_size, flags=flags)

    def transformToBirdViewPoints(self, points: list) -> Union[list, np.ndarray]:
        """[[x, y], ...] frontal-view points -> int bird-view points (float64 homography, truncation); [] for no points."""
        if not len(points):
            return []
        p = np.array([[x, y] for x, y in points])
        ones = np.ones((*p.shape[:-1], 1), dtype=p.dtype)
        # The truncation below exposes the last ulp whenever a coordinate lands on an integer, and the rounding of the 3-term dot
        # product depends on how it is accumulated: numpy's einsum kernel adds (M0*x + M2*1) + M1*y (measured; a BLAS matmul or a
        # left-to-right sum differs on ~1 point in 300).  The reference calls einsum (:139), so the same routine is used here.
        h = np.einsum("kl,...l->...k", self.M, np.concatenate([p, ones], axis=-1))
        return np.asarray(h[..., :2] / h[..., 2][..., None], dtype="int")

    def calcCurveAndOffset(self, img, left_lanes: np.ndarray, right_lanes: np.ndarray) -> Tuple[Tuple, Optional[float]]:
        """((direction "L"/"R"/"F", curvature in m), offset in m from the lane centre); ((None, None), None) without both lanes."""
        if not (len(left_lanes) and len(right_lanes)):
            return (None, None), None
        direction, curvature, offset, veh_pos, cen_pos, y_eval = curve_and_offset(left_lanes, right_lanes, img.shape[0], img.shape[1])
        cv2.arrowedLine(img, (int(veh_pos), int(y_eval)), (int(veh_pos), int(img.shape[1] / 3)), (255, 255, 255), 5, 0, 0, 0.2)
        cv2.arrowedLine(img, (int(cen_pos), int(y_eval)), (int(cen_pos), int(img.shape[0] / 1.3)), (150, 150, 150), 10, 0, 0, 0.5)
        cv2.putText(img, "Offset: %.1f m" % offset, (20, 80), cv2.FONT_HERSHEY_SIMPLEX, 3, (0, 0, 255), 5)
        cv2.putText(img, "R : %.1f m" % curvature, (20, 180), cv2.FONT_HERSHEY_SIMPLEX, 3, (0, 0, 255), 5)
        return (direction, curvature), offset

    def DrawDetectedOnBirdView(self, image, lanes_points: list, type: OffsetType = OffsetType.UNKNOWN) -> None:
        warn = {OffsetType.RIGHT: 1, OffsetType.LEFT: 2}.get(type)       # the ego lane being drifted over is drawn red
        for lane_num, lane_points in enumerate(lanes_points):
            color = (0, 0, 255) if lane_num == warn else lane_colors[lane_num]
            for x, y in lane_points:
                cv2.circle(image, (int(x), int(y)), 10, color, -1)

    def DrawTransformFrontalViewArea(self, image) -> None:
        pts = [tuple(int(v) for v in p) for p in self.src]
        for a, b in zip(pts, pts[1:] + pts[:1]):
            cv2.line(image, a, b, (0, 0, 255), 5)
