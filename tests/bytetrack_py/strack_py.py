"""TEST INFRASTRUCTURE: Python STrack (reference ObjectTracker/byteTrack/dtypes/strack.py:33-215) used by the Python state machine
in this directory.  `is_activated` only on frame 1 at birth (82-83), velocity-height zeroed for non-tracked states before prediction
(66-68), the class-id vote with its "new class starts at 2" quirk (122-129)."""
import numpy as np

import adas_b200  # noqa: F401
from adas_b200.ObjectTracker.byteTrack.strack import BaseTrack, LimitedList, TrackState

from . import kalman


class STrack(BaseTrack):
    def __init__(self, tlwh, score, class_id):
        self._tlwh = np.asarray(tlwh, dtype=float)
        self.mean = None
        self.covariance = None
        self.is_activated = False
        self.state = TrackState.New
        self.track_id = 0
        self.frame_id = 0
        self.start_frame = 0
        self.time_since_update = 0
        self.location = (np.inf, np.inf)
        self.crops = []
        self.score = score
        self.tracklet_len = 0
        self.class_id = class_id
        self.class_id_history = {class_id: 1}
        self.trajectories = LimitedList(30)

    # ---- geometry ----
    @property
    def tlwh(self):
        if self.mean is None:
            return self._tlwh.copy()
        r = self.mean[:4].copy()
        r[2] *= r[3]
        r[:2] -= r[2:] / 2
        return r

    @property
    def tlbr(self):
        r = self.tlwh
        r[2:] += r[:2]
        return r

    @property
    def xyah(self):
        return self.tlwh_to_xyah(self.tlwh)

    @property
    def end_frame(self):
        return self.frame_id

    @staticmethod
    def tlwh_to_xyah(tlwh):
        r = np.asarray(tlwh).copy()
        r[:2] += r[2:] / 2
        r[2] /= r[3]
        return r

    @staticmethod
    def tlbr_to_tlwh(tlbr):
        r = np.asarray(tlbr).copy()
        r[2:] -= r[:2]
        return r

    @staticmethod
    def tlwh_to_tlbr(tlwh):
        r = np.asarray(tlwh).copy()
        r[2:] += r[:2]
        return r

    # ---- life cycle ----
    @staticmethod
    def multi_predict(stracks):
        if not stracks:
            return
        means = np.asarray([t.mean.copy() for t in stracks])
        covs = np.asarray([t.covariance for t in stracks])
        for i, t in enumerate(stracks):
            if t.state != TrackState.Tracked:
                means[i][7] = 0
        means, covs = kalman.multi_predict(means, covs)
        for t, m, c in zip(stracks, means, covs):
            t.mean, t.covariance = m, c

    @staticmethod
    def multi_tlbr(stracks):
        """[N,4] tlbr of many tracks at once (same float operations as the `tlbr` property, vectorised)."""
        if not stracks:
            return np.zeros((0, 4), dtype=float)
        out = np.empty((len(stracks), 4), dtype=float)
        has = np.array([t.mean is not None for t in stracks])
        if has.any():
            m = np.asarray([t.mean[:4] for t, h in zip(stracks, has) if h])
            w = m[:, 2] * m[:, 3]
            x = m[:, 0] - w / 2
            y = m[:, 1] - m[:, 3] / 2
            out[has] = np.stack([x, y, w + x, m[:, 3] + y], 1)
        if (~has).any():
            r = np.asarray([t._tlwh for t, h in zip(stracks, has) if not h])
            out[~has] = np.stack([r[:, 0], r[:, 1], r[:, 2] + r[:, 0], r[:, 3] + r[:, 1]], 1)
        return out

    @staticmethod
    def multi_update(pairs, frame_id):
        """Kalman-correct every matched (track, detection) pair of one association stage in one batched solve, then apply
        the per-track bookkeeping of `update` (Tracked) / `re_activate` (Lost) -- strack.py:88-120 of the reference."""
        if not pairs:
            return
        means = np.asarray([t.mean for t, _ in pairs])
        covs = np.asarray([t.covariance for t, _ in pairs])
        zs = np.asarray([STrack.tlwh_to_xyah(d.tlwh) for _, d in pairs])
        nm, nc = kalman.multi_update(means, covs, zs)
        for (t, d), m, c in zip(pairs, nm, nc):
            t.mean, t.covariance = m, c
            if t.state == TrackState.Tracked:
                t.tracklet_len += 1
                t.trajectories.append(d.tlbr)
            else:
                t.tracklet_len = 0
            t.frame_id = frame_id
            t.state = TrackState.Tracked
            t.is_activated = True
            t.score = d.score
            t.update_class_id(d.class_id)

    def activate(self, frame_id):
        self.track_id = self.next_id()
        self.mean, self.covariance = kalman.initiate(self.tlwh_to_xyah(self._tlwh))
        self.tracklet_len = 0
        self.state = TrackState.Tracked
        if frame_id == 1:
            self.is_activated = True
        self.frame_id = frame_id
        self.start_frame = frame_id

    def re_activate(self, new_track, frame_id, new_id=False):
        self.mean, self.covariance = kalman.update(self.mean, self.covariance, self.tlwh_to_xyah(new_track.tlwh))
        self.tracklet_len = 0
        self.state = TrackState.Tracked
        self.is_activated = True
        self.frame_id = frame_id
        if new_id:
            self.track_id = self.next_id()
        self.score = new_track.score
        self.update_class_id(new_track.class_id)

    def update(self, new_track, frame_id):
        self.frame_id = frame_id
        self.tracklet_len += 1
        self.mean, self.covariance = kalman.update(self.mean, self.covariance, self.tlwh_to_xyah(new_track.tlwh))
        self.trajectories.append(new_track.tlbr)
        self.state = TrackState.Tracked
        self.is_activated = True
        self.score = new_track.score
        self.update_class_id(new_track.class_id)

    def update_class_id(self, class_id) -> None:
        self.class_id_history[class_id] = self.class_id_history.get(class_id, 1) + 1
        self.class_id = max(self.class_id_history, key=self.class_id_history.get)

    def update_crops(self, frame) -> None:
        if frame is None:
            return
        tx1, ty1, tw, th = self._tlwh.astype(int)
        x1, y1 = max(0, tx1), max(0, ty1)
        x2, y2 = min(frame.shape[1], tx1 + tw), min(frame.shape[0], ty1 + th)
        self.crops.append(frame[y1:y2, x1:x2, :].copy())

    def filter_trajectories(self, frame, pad=(0, 0)):
        ph, pw = pad
        return [b for b in list(self.trajectories)
                if b[0] >= pw and b[1] >= ph and b[2] <= frame.shape[1] - pw and b[3] <= frame.shape[0] - ph]

    def mark_lost(self):
        self.state = TrackState.Lost

    def mark_removed(self):
        self.state = TrackState.Removed

    def __repr__(self):
        return f"OT_{self.track_id}_({self.start_frame}-{self.end_frame})"

    def get_track_message(self):
        return {
            "track_id": self.track_id, "count": BaseTrack._count, "is_activated": self.is_activated, "state": self.state,
            "score": self.score, "start_frame_number": self.start_frame, "curr_frame_number": self.frame_id,
            "time_since_update": self.time_since_update, "location": str(self.location), "crops": self.crops, "class_id": self.class_id,
        }
