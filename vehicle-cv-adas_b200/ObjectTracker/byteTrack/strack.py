"""Track-record vocabulary shared by the tracker views: `TrackState` (base_track.py:5-9), the 30-entry trajectory ring
`LimitedList` (dtypes/strack.py:8-31), the `BaseTrack` id counter facade (base_track.py:12,33-36) and the tlwh -> xyah conversion
(strack.py:175-184).  The track state itself (Kalman mean / covariance, votes, life cycle) lives in the native tracker
(csrc/tracker.cu); `TrackView` in byteTracker.py exposes it with the reference's STrack attribute names."""
import numpy as np


class TrackState:
    New = 0
    Tracked = 1
    Lost = 2
    Removed = 3


class LimitedList(list):
    def __init__(self, maxlen):
        super().__init__()
        self._maxlen = maxlen

    def full(self):
        return len(self) >= self._maxlen

    def append(self, element):
        if len(self) == self._maxlen:
            del self[0]
        super().append(element)

    def extend(self, elements):
        for e in elements:
            self.append(e)


class BaseTrack:
    _count = 0

    @staticmethod
    def next_id():
        BaseTrack._count += 1
        return BaseTrack._count

    @staticmethod
    def reset_counter():
        BaseTrack._count = 0




def tlwh_to_xyah(tlwh):
    r = np.asarray(tlwh).copy()
    r[:2] += r[2:] / 2
    r[2] /= r[3]
    return r
