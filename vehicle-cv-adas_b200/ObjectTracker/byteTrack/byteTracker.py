"""BYTETracker with the reference's API (ObjectTracker/byteTrack/byteTracker.py:12-215).

`update(bboxes_xyxy, scores, class_ids, frame)` follows the reference's three association stages (62-185):
thresholds 0.5 / 0.1 / det 0.6 (48,73-75), match 0.8 (fused) / 0.5 (plain IoU) / 0.7 (fused) (108,130,152),
30-frame lost buffer (49), duplicate removal at IoU-distance < 0.15 (utils.py:54-69).  Each stage's cost
matrix + exact assignment is one device call (`matching.associate`); Kalman and list bookkeeping stay on the
host because they are sequential per stream.
"""
import numpy as np

from ..core import ObjectTrackBase
from . import matching
from .strack import BaseTrack, LimitedList, STrack, TrackState


def joint_stracks(a, b):
    seen, out = set(), []
    for t in list(a) + list(b):
        if t.track_id not in seen:
            seen.add(t.track_id)
            out.append(t)
    return out


def sub_stracks(a, b):
    drop = {t.track_id for t in b}
    keep = {}
    for t in a:
        keep[t.track_id] = t          # later duplicates replace earlier ones, first position kept (dict semantics)
    return [t for tid, t in keep.items() if tid not in drop]


def remove_duplicate_stracks(a, b):
    dist = matching.iou_distance(a, b)
    dup_a, dup_b = set(), set()
    for ia, ib in zip(*np.where(dist < 0.15)):
        age_a = a[ia].frame_id - a[ia].start_frame
        age_b = b[ib].frame_id - b[ib].start_frame
        if age_a > age_b:
            dup_b.add(ib)
        else:
            dup_a.add(ia)
    return [t for i, t in enumerate(a) if i not in dup_a], [t for i, t in enumerate(b) if i not in dup_b]


class BYTETrackerPy(ObjectTrackBase):
    """Python state machine (kept for A/B tests); `BYTETracker` below is the native one the product uses."""

    def __init__(self, track_thresh=0.5, track_buffer=30, match_thresh=0.8, frame_rate=30, min_box_area=10, device=0, **kwargs):
        super().__init__(**kwargs)
        self.tracked_stracks, self.lost_stracks, self.removed_stracks = [], [], []
        self.track_thresh, self.match_thresh, self.min_box_area = track_thresh, match_thresh, min_box_area
        self.frame_id = 0
        self.det_thresh = track_thresh + 0.1
        self.buffer_size = int(frame_rate / 30.0 * track_buffer)
        self.max_time_lost = self.buffer_size
        matching.DEVICE = device

    def _get_tracker_messages(self, status=TrackState.Tracked):
        pool = {TrackState.Lost: self.lost_stracks, TrackState.Removed: self.removed_stracks}.get(status, self.tracked_stracks)
        return [t.get_track_message() for t in pool]

    @staticmethod
    def _make(dets, scores, cids):
        return [STrack(STrack.tlbr_to_tlwh(b), s, c) for b, s, c in zip(dets, scores, cids)] if len(dets) > 0 else []

    def _apply(self, matches, tracks, dets, activated, refind):
        """Matched pairs of one stage: batched Kalman correction, then route to activated (was Tracked) / refind (was Lost)."""
        pairs = [(tracks[it], dets[idet]) for it, idet in matches]
        was_tracked = [t.state == TrackState.Tracked for t, _ in pairs]
        STrack.multi_update(pairs, self.frame_id)
        for (t, _), wt in zip(pairs, was_tracked):
            (activated if wt else refind).append(t)

    def update(self, bboxes, scores, class_ids, frame=None):
        self.frame_id += 1
        activated, refind, lost, removed = [], [], [], []
        bboxes, scores, class_ids = np.array(bboxes), np.array(scores), np.array(class_ids)
        high = scores > self.track_thresh
        second = np.logical_and(scores > 0.1, scores < self.track_thresh)
        detections = self._make(bboxes[high], scores[high], class_ids[high])
        detections_second = self._make(bboxes[second], scores[second], class_ids[second])

        unconfirmed = [t for t in self.tracked_stracks if not t.is_activated]
        tracked = [t for t in self.tracked_stracks if t.is_activated]

        # stage 1: tracked + lost vs high-score detections (fused cost, match_thresh)
        pool = joint_stracks(tracked, self.lost_stracks)
        STrack.multi_predict(pool)
        matches, u_track, u_det = matching.associate(pool, detections, self.match_thresh, fuse=True)
        self._apply(matches, pool, detections, activated, refind)

        # stage 2: still-tracked leftovers vs low-score detections (plain IoU, 0.5)
        r_tracked = [pool[i] for i in u_track if pool[i].state == TrackState.Tracked]
        matches, u_track2, _ = matching.associate(r_tracked, detections_second, 0.5, fuse=False)
        self._apply(matches, r_tracked, detections_second, activated, refind)
        for it in u_track2:
            t = r_tracked[it]
            if t.state != TrackState.Lost:
                t.mark_lost()
                lost.append(t)

        # stage 3: unconfirmed (one-frame-old) tracks vs leftover high detections (fused cost, 0.7)
        detections = [detections[i] for i in u_det]
        matches, u_unconf, u_det = matching.associate(unconfirmed, detections, 0.7, fuse=True)
        self._apply(matches, unconfirmed, detections, activated, activated)
        for it in u_unconf:
            unconfirmed[it].mark_removed()
            removed.append(unconfirmed[it])

        # births
        for i in u_det:
            t = detections[i]
            if t.score < self.det_thresh:
                continue
            t.activate(self.frame_id)
            t.update_crops(frame)
            activated.append(t)

        # ageing + list maintenance
        for t in self.lost_stracks:
            if self.frame_id - t.end_frame > self.max_time_lost:
                t.mark_removed()
                removed.append(t)
        self.tracked_stracks = [t for t in self.tracked_stracks if t.state == TrackState.Tracked]
        self.tracked_stracks = joint_stracks(self.tracked_stracks, activated)
        self.tracked_stracks = joint_stracks(self.tracked_stracks, refind)
        self.lost_stracks = sub_stracks(self.lost_stracks, self.tracked_stracks)
        self.lost_stracks.extend(lost)
        self.lost_stracks = sub_stracks(self.lost_stracks, self.removed_stracks)
        self.removed_stracks.extend(removed)
        self.tracked_stracks, self.lost_stracks = remove_duplicate_stracks(self.tracked_stracks, self.lost_stracks)
        return self._get_tracker_messages()

    def reset(self):
        self.frame_id = 0
        self.tracked_stracks, self.lost_stracks, self.removed_stracks = [], [], []
        BaseTrack.reset_counter()

    def DrawTrackedOnFrame(self, frame, show_box=True, show_traject=True):
        for t in [t for t in self.tracked_stracks if t.is_activated]:
            tlwh = t.tlwh
            if tlwh[2] * tlwh[3] > self.min_box_area:
                if show_box:
                    self.plot_bbox(frame, tlwh, t.class_id, t.track_id)
                if show_traject:
                    self.plot_trajectories(frame, t.trajectories, t.class_id, t.track_id)
                    self.plot_directions(frame, t.xyah, t.filter_trajectories(frame, (10, 10)), t.class_id)


class TrackView:
    """Read-only view of one native track with the attribute surface of the reference's STrack (strack.py:33-215)."""
    __slots__ = ("track_id", "is_activated", "state", "score", "class_id", "start_frame", "frame_id", "tracklet_len", "mean", "_tlwh",
                 "trajectories", "crops", "time_since_update", "location")

    def __init__(self):
        self.trajectories = LimitedList(30)
        self.crops = []
        self.time_since_update = 0
        self.location = (np.inf, np.inf)

    @property
    def tlwh(self):
        return self._tlwh.copy()

    @property
    def tlbr(self):
        r = self._tlwh.copy()
        r[2:] += r[:2]
        return r

    @property
    def xyah(self):
        return STrack.tlwh_to_xyah(self._tlwh)

    @property
    def end_frame(self):
        return self.frame_id

    def filter_trajectories(self, frame, pad=(0, 0)):
        ph, pw = pad
        return [b for b in list(self.trajectories)
                if b[0] >= pw and b[1] >= ph and b[2] <= frame.shape[1] - pw and b[3] <= frame.shape[0] - ph]

    def get_track_message(self, count):
        return {"track_id": self.track_id, "count": count, "is_activated": self.is_activated, "state": self.state, "score": self.score,
                "start_frame_number": self.start_frame, "curr_frame_number": self.frame_id, "time_since_update": self.time_since_update,
                "location": str(self.location), "crops": self.crops, "class_id": self.class_id}

    def __repr__(self):
        return f"OT_{self.track_id}_({self.start_frame}-{self.end_frame})"


class BYTETracker(ObjectTrackBase):
    """The reference's BYTETracker API (byteTracker.py:12-215) over the native tracker (csrc/tracker.cu): `update` is one
    library call per frame -- Kalman + bookkeeping in host C++, the three association stages on the device."""

    def __init__(self, track_thresh=0.5, track_buffer=30, match_thresh=0.8, frame_rate=30, min_box_area=10, device=0, **kwargs):
        super().__init__(**kwargs)
        from ... import _capi
        self._capi = _capi
        self._nt = _capi.NativeTracker(device, track_thresh, track_buffer, match_thresh, frame_rate)
        self.track_thresh, self.match_thresh, self.min_box_area = track_thresh, match_thresh, min_box_area
        self.det_thresh = track_thresh + 0.1
        self.buffer_size = int(frame_rate / 30.0 * track_buffer)
        self.max_time_lost = self.buffer_size
        self.frame_id = 0
        self._labels, self._label_list = {}, []
        self._views = {}
        self.tracked_stracks = []
        self.removed_stracks = []

    def _cid(self, label):
        k = label.item() if hasattr(label, "item") else label
        i = self._labels.get(k)
        if i is None:
            i = self._labels[k] = len(self._label_list)
            self._label_list.append(k)
        return i

    def _view(self, rec, frame=None):
        tid = int(rec["track_id"])
        v = self._views.get(tid)
        if v is None:
            v = self._views[tid] = TrackView()
            if frame is not None:                      # STrack.update_crops at birth (strack.py:131-143, byteTracker.py:167)
                tx1, ty1, tw, th = rec["tlwh"].astype(int)
                x1, y1 = max(0, tx1), max(0, ty1)
                x2, y2 = min(frame.shape[1], tx1 + tw), min(frame.shape[0], ty1 + th)
                v.crops.append(frame[y1:y2, x1:x2, :].copy())
        v.track_id, v.is_activated, v.state = tid, bool(rec["is_activated"]), int(rec["state"])
        v.score, v.class_id = float(rec["score"]), self._label_list[int(rec["class_id"])]
        v.start_frame, v.frame_id, v.tracklet_len = int(rec["start_frame"]), int(rec["frame_id"]), int(rec["tracklet_len"])
        v.mean, v._tlwh = rec["mean"].copy(), rec["tlwh"].copy()
        if int(rec["traj_frame"]) == self.frame_id:
            v.trajectories.append(rec["det_tlbr"].copy())
        return v

    def update(self, bboxes, scores, class_ids, frame=None):
        self.frame_id += 1
        ids = np.fromiter((self._cid(c) for c in class_ids), dtype=np.int32, count=len(class_ids)) if len(class_ids) else np.zeros(0, np.int32)
        recs = self._nt.update(np.asarray(bboxes, dtype=np.float64).reshape(-1, 4), np.asarray(scores, dtype=np.float64), ids)
        self.tracked_stracks = [self._view(r, frame) for r in recs]
        if len(self._views) > 4 * max(64, len(recs)):                 # forget views of long-gone tracks
            alive = {int(r["track_id"]) for r in recs} | {int(r["track_id"]) for r in self._nt.get(1)}
            self._views = {k: v for k, v in self._views.items() if k in alive}
        cnt = self._capi.NativeTracker.count()
        return [t.get_track_message(cnt) for t in self.tracked_stracks]

    def update_batch(self, frames_dets, max_out: int = 256):
        """All frames of a pipeline step in ONE library call (no interpreter work between frames).
        frames_dets: sequence of (bboxes xyxy [n,4], scores [n], class_ids [n]) per frame, in time order.
        Returns one TRACK_DTYPE record array per frame (the tracked_stracks of that frame); `messages(recs)` turns a record array into
        the reference's track messages (strack.py:207-215).  `tracked_stracks` afterwards reflects the last frame."""
        counts = np.fromiter((len(d[1]) for d in frames_dets), dtype=np.int32, count=len(frames_dets))
        tot = int(counts.sum())
        boxes = np.zeros((tot, 4), np.float64)
        scores = np.zeros(tot, np.float64)
        ids = np.zeros(tot, np.int32)
        o = 0
        for (bb, sc, cl), n in zip(frames_dets, counts):
            if n:
                boxes[o:o + n] = np.asarray(bb, dtype=np.float64).reshape(-1, 4)
                scores[o:o + n] = np.asarray(sc, dtype=np.float64)
                ids[o:o + n] = [self._cid(c) for c in (cl.tolist() if hasattr(cl, "tolist") else cl)]
            o += n
        recs = self._nt.update_batch(counts, boxes, scores, ids, max_out)
        self.frame_id += len(frames_dets)
        if recs:
            self.tracked_stracks = [self._view(r, None) for r in recs[-1]]
        return recs

    def update_batch_arrays(self, counts, xyxy, scores, class_ids, max_out: int = 256):
        """update_batch on already concatenated arrays (the pipeline's hot path): counts [F] int, xyxy [sum, 4], scores [sum],
        class_ids [sum] integer labels.  Labels are mapped to the tracker's class slots in first-seen order like `update` does."""
        cl = np.asarray(class_ids)
        if cl.size:
            uniq, first = np.unique(cl, return_index=True)
            for u in uniq[np.argsort(first)].tolist():            # register unseen labels in order of first appearance
                self._cid(u)
            lut = np.array([self._labels[u] for u in uniq.tolist()], np.int32)
            ids = lut[np.searchsorted(uniq, cl)]
        else:
            ids = np.zeros(0, np.int32)
        recs = self._nt.update_batch(counts, xyxy, scores, ids, max_out)
        self.frame_id += len(counts)
        if recs:
            self.tracked_stracks = [self._view(r, None) for r in recs[-1]]
        return recs

    def messages(self, recs):
        """Track messages (STrack.get_track_message, strack.py:207-215) of one frame's record array."""
        return [{"track_id": int(r["track_id"]), "count": int(r["pad"]), "is_activated": bool(r["is_activated"]), "state": int(r["state"]),
                 "score": float(r["score"]), "start_frame_number": int(r["start_frame"]), "curr_frame_number": int(r["frame_id"]),
                 "time_since_update": 0, "location": str((np.inf, np.inf)), "crops": [], "class_id": self._label_list[int(r["class_id"])]}
                for r in recs]

    @property
    def lost_stracks(self):
        out = []
        for r in self._nt.get(1):
            v = self._views.get(int(r["track_id"])) or TrackView()
            v.track_id, v.is_activated, v.state = int(r["track_id"]), bool(r["is_activated"]), int(r["state"])
            v.score, v.class_id = float(r["score"]), self._label_list[int(r["class_id"])]
            v.start_frame, v.frame_id, v.tracklet_len = int(r["start_frame"]), int(r["frame_id"]), int(r["tracklet_len"])
            v.mean, v._tlwh = r["mean"].copy(), r["tlwh"].copy()
            out.append(v)
        return out

    def reset(self):
        self.frame_id = 0
        self.tracked_stracks, self.removed_stracks, self._views = [], [], {}
        self._nt.reset()

    def DrawTrackedOnFrame(self, frame, show_box=True, show_traject=True):
        for t in [t for t in self.tracked_stracks if t.is_activated]:
            tlwh = t.tlwh
            if tlwh[2] * tlwh[3] > self.min_box_area:
                if show_box:
                    self.plot_bbox(frame, tlwh, t.class_id, t.track_id)
                if show_traject:
                    self.plot_trajectories(frame, t.trajectories, t.class_id, t.track_id)
                    self.plot_directions(frame, t.xyah, t.filter_trajectories(frame, (10, 10)), t.class_id)
