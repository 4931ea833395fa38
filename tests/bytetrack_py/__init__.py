"""TEST INFRASTRUCTURE (not product code): a Python ByteTrack state machine that drives the library's STANDALONE association entry
points (`adas_associate` / `adas_lap`, through adas_b200.ObjectTracker.byteTrack.matching) once per association stage.

The product tracker is the native one (`adas_b200.ObjectTracker.BYTETracker` -> csrc/tracker.cu).  This second implementation of
BYTETracker.update (reference ObjectTracker/byteTrack/byteTracker.py:62-185) exists so that the golden tracker sequences also
exercise `adas_associate` (cost matrix + exact assignment per stage) end to end."""
import numpy as np

import adas_b200  # noqa: F401
from adas_b200.ObjectTracker.byteTrack import matching
from adas_b200.ObjectTracker.byteTrack.strack import BaseTrack, LimitedList, TrackState
from adas_b200.ObjectTracker.core import ObjectTrackBase

from .strack_py import STrack


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
    """Python state machine; the product uses the native `adas_b200.ObjectTracker.BYTETracker`."""

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
