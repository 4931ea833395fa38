"""Bird-view geometry (host side): adas_b200's PerspectiveTransformation against golden vectors produced by the
reference class on the same seeded lanes (tests/golden/make_golden.py birdview)."""
import hashlib
import os

import numpy as np

import adas_b200  # noqa: F401
from adas_b200.TrafficLaneDetector.ufldDetector.perspectiveTransformation import PerspectiveTransformation, curve_and_offset
import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "birdview.npz")


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def test_birdview_matches_reference_golden():
    g = np.load(GOLD)
    dirs = set()
    for case, (seed, kind) in enumerate([(s, k) for s in range(8) for k in ("Default", "Top", "Bottom", None)]):
        left, right = synth.ego_lanes(100 + seed)
        t = PerspectiveTransformation((1280, 720))
        if kind is not None:
            t.updateTransformParams(left, right, kind)
        assert np.array_equal(t.src, g[f"c{case}_src"])
        assert np.array_equal(t.M, g[f"c{case}_M"]) and np.array_equal(t.M_inv, g[f"c{case}_Minv"])
        bl, br = t.transformToBirdViewPoints(left), t.transformToBirdViewPoints(right)
        assert bl.dtype.kind == "i" and np.array_equal(bl, g[f"c{case}_bl"]) and np.array_equal(br, g[f"c{case}_br"])      # integer points: bit-exact
        img = np.zeros((720, 1280, 3), np.uint8)
        (direction, curv), off = t.calcCurveAndOffset(img, bl, br)
        want = g[f"c{case}_curve"]
        assert {"L": -1.0, "F": 0.0, "R": 1.0}[direction] == want[0]
        # float64 least squares through the same LAPACK: 1e-9 relative is the tolerance of this test
        assert abs(curv - want[1]) <= 1e-9 * abs(want[1]) and abs(off - want[2]) <= 1e-9 * max(1.0, abs(want[2]))
        assert np.array_equal(_sha(img), g[f"c{case}_draw_sha"])                   # arrows + text drawn at the same pixels
        assert np.array_equal(_sha(t.transformToBirdView(synth.frame(seed))), g[f"c{case}_warp_sha"])
        dirs.add(direction)
    assert len(dirs) >= 2, "the seeded lanes should exercise more than one curvature direction"


def test_birdview_edge_cases():
    t = PerspectiveTransformation((1280, 720))
    assert t.transformToBirdViewPoints([]) == []
    assert t.calcCurveAndOffset(np.zeros((720, 1280, 3), np.uint8), [], []) == ((None, None), None)
    m0 = t.M.copy()
    t.updateTransformParams([], [(1, 2)], "Default")           # one lane missing: unchanged
    t.updateTransformParams([(1, 2)], [(3, 4)], "Sideways")    # unknown type: unchanged
    assert np.array_equal(t.M, m0)
    left, right = synth.ego_lanes(3)
    t.updateTransformParams(np.array(left), np.array(right), "Top")      # ndarray inputs are accepted like lists
    assert not np.array_equal(t.M, m0)
    # frontal -> bird -> frontal is the identity up to interpolation on the inverse matrix
    assert np.allclose(t.M @ t.M_inv / (t.M @ t.M_inv)[2, 2], np.eye(3), atol=1e-6)
    d, c, o, *_ = curve_and_offset(t.transformToBirdViewPoints(left), t.transformToBirdViewPoints(right), 720, 1280)
    assert d in ("L", "R", "F") and c > 0 and np.isfinite(o)
