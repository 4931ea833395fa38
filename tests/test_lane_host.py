"""CPU: SURVEY 8a row K -- the host tail of the lane detector (`LaneDetectBase.__update_lanes_status`, `__update_lanes_area`,
`__adjust_lanes_points`, reference TrafficLaneDetector/ufldDetector/core.py:102-158) against golden vectors produced by the
UNMODIFIED reference class (tests/golden/make_golden.py: `_area`, `_area_adj`, `_area_status` of ufld_post.npz).
The product class is fed the reference's own lane points (`_lane{l}` / `_status`), so only row K is under test."""
import os

import numpy as np
import pytest

import adas_b200  # noqa: F401
from adas_b200.TrafficLaneDetector.ufldDetector.core import LaneDetectBase


class _Host(LaneDetectBase):
    """concrete shell: the abstract drawing / inference entry points are not under test"""
    _defaults = {}

    def DetectFrame(self):
        return None

    def DrawDetectedOnFrame(self):
        return None

    def DrawAreaOnFrame(self):
        return None


KEYS = ["s0_720x1280", "s0_480x640", "s1_720x1280", "s2_720x1280", "s3_720x1280"]


@pytest.mark.parametrize("key", KEYS)
@pytest.mark.parametrize("adjust", [False, True])
def test_lane_area_matches_reference(golden_dir, key, adjust):
    g = np.load(os.path.join(golden_dir, "ufld_post.npz"))
    h = int(key.split("_")[1].split("x")[0])
    lanes = [[(int(x), int(y)) for x, y in g[f"{key}_lane{l}"]] for l in range(4)]
    status = [bool(v) for v in g[key + "_status"]]
    det = _Host(None)
    det.adjust_lanes = adjust
    det._update_lanes_status(status)
    assert det.lane_info.area_status == bool(g[key + "_area_status"][0])
    pts = np.empty(4, dtype=object)
    for l in range(4):
        pts[l] = lanes[l]
    det._update_lanes_area(pts, h)
    want = g[key + ("_area_adj" if adjust else "_area")]
    got = np.array(det.lane_info.area_points, np.int32).reshape(-1, 2) if det.lane_info.area_status else np.zeros((0, 2), np.int32)
    assert got.shape == want.shape, (key, adjust, got.shape, want.shape)
    assert np.array_equal(got, want), (key, adjust)


def test_lane_status_edge_cases():
    det = _Host(None)
    for status, want in (([], False), ([True, True, True], False), ([False, True, True, False], True), ([True, False, True, True], False),
                         ([True, True], True)):
        det._update_lanes_status(status)
        assert det.lane_info.area_status == want, status
    # fewer than 11 points on a side: the fit is skipped and the raw points are kept (core.py:108-117)
    left = [(100 + i, 400 + 10 * i) for i in range(8)]
    right = [(600 - i, 400 + 10 * i) for i in range(20)]
    l2, r2 = LaneDetectBase._adjust_lanes_points(left, right, 720)
    assert l2 == left and r2 == right
