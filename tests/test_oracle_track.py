"""CPU: the oracle ByteTrack restatement (oracle/track.py) against sequences produced by the unmodified reference."""
import os

import numpy as np

import synth
from oracle import track


def test_tracker_matches_reference_sequences(golden_dir):
    g = np.load(os.path.join(golden_dir, "track.npz"))
    for seed, nobj in ((0, 8), (1, 14), (2, 4), (3, 25)):
        trk = track.Tracker()
        trk.reset()
        rows = []
        for f, (boxes, scores, labels) in enumerate(synth.track_sequence(seed, frames=45, objects=nobj)):
            trk.update(boxes, scores, labels)
            for t in trk.tracked:
                tl = t.tlwh()
                rows.append([f, t.tid, int(t.activated), t.state, tl[0], tl[1], tl[2], tl[3], float(t.score), int(str(t.cls)[5:])])
            for t in trk.lost:
                rows.append([f, t.tid, -1, t.state, 0, 0, 0, 0, 0, -1])
        got, gold = np.array(rows, np.float64), g[f"seq{seed}"]
        assert got.shape == gold.shape, (seed, got.shape, gold.shape)
        assert np.array_equal(got[:, [0, 1, 2, 3, 9]], gold[:, [0, 1, 2, 3, 9]])
        assert np.allclose(got[:, 4:9], gold[:, 4:9], rtol=0, atol=1e-9)
    lost_seen = any((g[f"seq{s}"][:, 3] == 2).any() for s in range(4))
    assert lost_seen          # the sequences exercise lost -> refind / removal paths
