"""Association primitives, executed by the device kernels in csrc/track.cu.

Mirror of ObjectTracker/byteTrack/matching.py: `iou_distance` (55-80), `fuse_score` (108-116) and
`linear_assignment` (20-31, lap.lapjv with extend_cost / cost_limit).  `associate` fuses the three into one
library call per association stage.
"""
import numpy as np

from ... import _capi

DEVICE = 0


def _tlbrs(tracks):
    if len(tracks) > 0 and isinstance(tracks[0], np.ndarray):
        return np.ascontiguousarray(tracks, dtype=float).reshape(-1, 4)
    return np.ascontiguousarray([np.asarray(t.tlbr, dtype=float) for t in tracks], dtype=float).reshape(-1, 4)


def iou_distance(atracks, btracks):
    a, b = _tlbrs(atracks), _tlbrs(btracks)
    if a.shape[0] == 0 or b.shape[0] == 0:
        return np.zeros((a.shape[0], b.shape[0]), dtype=float)
    # the single-problem entry keeps its device scratch alive between calls (no cudaMalloc on the per-frame path)
    return _capi.associate(a, b, None, 2.0, device=DEVICE, want_cost=True)[2]


def fuse_score(cost_matrix, detections):
    if cost_matrix.size == 0:
        return cost_matrix
    scores = np.array([d.score for d in detections], dtype=float)
    return 1 - (1 - cost_matrix) * scores[None, :]


def linear_assignment(cost_matrix, thresh):
    if cost_matrix.size == 0:
        return np.empty((0, 2), dtype=int), tuple(range(cost_matrix.shape[0])), tuple(range(cost_matrix.shape[1]))
    x, y = _capi.lap([np.asarray(cost_matrix, float)], [thresh], device=DEVICE)[0]
    matches = np.asarray([[i, m] for i, m in enumerate(x) if m >= 0])
    return matches, np.where(x < 0)[0], np.where(y < 0)[0]


def associate(tracks, detections, thresh, fuse):
    """iou_distance -> [fuse_score] -> linear_assignment in one device call."""
    if len(tracks) == 0 or len(detections) == 0:
        return np.empty((0, 2), dtype=int), tuple(range(len(tracks))), tuple(range(len(detections)))
    scores = np.array([d.score for d in detections], dtype=float) if fuse else None
    x, y, _ = _capi.associate(_tlbrs(tracks), _tlbrs(detections), scores, thresh, device=DEVICE)
    matches = np.asarray([[i, m] for i, m in enumerate(x) if m >= 0])
    return matches, np.where(x < 0)[0], np.where(y < 0)[0]
