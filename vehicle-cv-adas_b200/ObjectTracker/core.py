"""ObjectTrackBase -- drawing helpers of the reference's ObjectTracker/core.py:68-245 (host-side, out of the hot path)."""
import numpy as np


class ObjectTrackBase:
    def __init__(self, names=None, **kwargs):
        self.names = list(names) if names is not None else []
        rng = np.random.default_rng(7)
        self._palette = rng.integers(0, 255, size=(256, 3)).tolist()

    def _color(self, tid):
        return tuple(int(c) for c in self._palette[int(tid) % 256])

    def plot_bbox(self, frame, tlwh, class_id, track_id):
        import cv2
        x, y, w, h = [int(v) for v in tlwh]
        cv2.rectangle(frame, (x, y), (x + w, y + h), self._color(track_id), 2)
        cv2.putText(frame, f"{class_id}-{track_id}", (x, max(0, y - 4)), cv2.FONT_HERSHEY_SIMPLEX, 0.5, self._color(track_id), 1)

    def plot_trajectories(self, frame, trajectories, class_id, track_id):
        import cv2
        pts = [(int((b[0] + b[2]) / 2), int(b[3])) for b in trajectories]
        for a, b in zip(pts[:-1], pts[1:]):
            cv2.line(frame, a, b, self._color(track_id), 2)

    def plot_directions(self, frame, xyah, trajectories, class_id):
        import cv2
        if len(trajectories) >= 2:
            a, b = trajectories[0], trajectories[-1]
            p0 = (int((a[0] + a[2]) / 2), int((a[1] + a[3]) / 2))
            p1 = (int((b[0] + b[2]) / 2), int((b[1] + b[3]) / 2))
            cv2.arrowedLine(frame, p0, p1, (255, 255, 255), 1)
