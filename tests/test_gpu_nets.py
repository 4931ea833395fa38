"""GPU: whole networks and the reference-facing API through the C ABI.

Network parity is against the fp32 torch-CPU oracle sharing the seeded weights.  The product computes fp16 x fp16
-> fp32 (the reference's own `_fp16` engines do the same, coreEngine.py:168, onnxQuantization.py:38); the tolerance
on probabilities / lane logits is written next to each assert.  Everything downstream of the raw head tensor is
bit-exact and is checked as such (fused path == reference post-processing applied to the device's own raw tensor).
"""
import os

import numpy as np
import pytest
import torch

import synth
import adas_b200  # noqa: F401
from adas_b200 import _capi
from adas_b200.coreEngine import B200Engine
from gpu_util import cached_plan
from oracle import nets, post

pytestmark = pytest.mark.gpu
torch.set_num_threads(min(16, max(1, os.cpu_count() or 1)))   # oneDNN collapses under 100+ threads


def _blob(frames):
    return np.concatenate([post.yolo_prepare_input(f, 640, 640)[0] for f in frames])


def _report(name, got, ref):
    d = np.abs(got - ref)
    print(f"[parity] {name}: max_abs={d.max():.3e} mean_abs={d.mean():.3e} ref_absmax={np.abs(ref).max():.3e}")
    return float(d.max())


@pytest.mark.parametrize("impl", [1, 0])
def test_yolov5n_engine_vs_oracle(impl):
    path, sd, _ = cached_plan("yolov5", scale="n")
    eng = _capi.Engine(path, 0, max_batch=2, conv_impl=impl)
    frames = [synth.frame(s) for s in (0, 1)]
    x = _blob(frames)
    raw = eng.infer(x)[0]
    model = nets.build("yolov5", sd, scale="n")
    with torch.no_grad():
        ref = model(torch.from_numpy(x)).numpy()
    assert raw.shape == ref.shape == (2, 25200, 85)
    e_prob = _report(f"v5n impl{impl} obj/cls prob", raw[..., 4:], ref[..., 4:])
    e_box = _report(f"v5n impl{impl} box px", raw[..., :4], ref[..., :4])
    assert e_prob < 5e-3          # fp16 operands through 60 convs vs fp32: probabilities within 5e-3
    assert e_box < 0.5            # boxes within half a pixel of the 640-px input
    eng.close()


def test_yolov8l_engine_vs_oracle_and_batch_invariance():
    path, sd, _ = cached_plan("yolov8", scale="l")
    eng = _capi.Engine(path, 0, max_batch=4)
    frames = [synth.frame(s) for s in (0, 1, 2, 3)]
    x = _blob(frames)
    raw4 = eng.infer(x)[0]
    model = nets.build("yolov8", sd, scale="l")
    with torch.no_grad():
        ref = model(torch.from_numpy(x[:2])).numpy()
    assert raw4.shape == (4, 84, 8400)
    e_prob = _report("v8l cls prob", raw4[:2, 4:], ref[:, 4:])
    e_box = _report("v8l box px", raw4[:2, :4], ref[:, :4])
    assert e_prob < 5e-3
    assert e_box < 1.0
    # per-frame results are independent of the batch they ran in (deterministic tiles, no split-K)
    raw1 = eng.infer(x[2:3])[0]
    assert np.array_equal(raw1[0], raw4[2])
    # SIMT validation kernels agree with the tensor-core path to accumulation-order noise
    eng_s = _capi.Engine(path, 0, max_batch=1, conv_impl=1)
    raw_s = eng_s.infer(x[:1])[0]
    # (two fp16-operand paths with different accumulation order and SiLU evaluation: rounding noise, bounded below the oracle tolerance)
    assert _report("v8l tc vs simt prob", raw4[:1, 4:], raw_s[:, 4:]) < 3e-3
    eng.close()
    eng_s.close()


def test_yolov8l_fused_detect_matches_reference_postprocessing():
    path, sd, _ = cached_plan("yolov8", scale="l")
    eng = _capi.Engine(path, 0, max_batch=4)
    frames = np.stack([synth.frame(s) for s in (4, 5, 6, 7)])
    boxes, scores, cls, idx, counts, ncand = eng.yolo_detect(frames, 0.4, 0.45)
    # (a) pre-processing inside the fused path is the bit-exact blob, so engine_inference on it gives the same raw tensor
    x = _capi.yolo_preprocess(frames, (640, 640))
    assert np.array_equal(x, _blob(frames))
    raw = eng.infer(x)[0]
    geom = post.letterbox_geom(720, 1280, 640, 640)
    total = 0
    for b in range(4):
        r = post.yolo_postprocess(raw[b], "v8", geom, 0.4, 0.45)     # the reference's host post-processing, restated
        n = int(counts[b])
        total += n
        assert ncand[b] == r["n_cand"]
        assert np.array_equal(idx[b, :n], r["idx"])
        assert np.array_equal(boxes[b, :n], r["boxes"])
        assert np.array_equal(scores[b, :n], r["scores"])
        assert np.array_equal(cls[b, :n], r["cls"])
    assert total > 0, "synthetic operating point produced no detections"
    print("[parity] v8l fused detect: per-frame detections", counts.tolist(), "candidates", ncand.tolist())
    # (b) against the fp32 oracle end to end, restricted to candidates whose score margin exceeds the network tolerance
    model = nets.build("yolov8", sd, scale="l")
    with torch.no_grad():
        ref = model(torch.from_numpy(x[:1])).numpy()[0]
    mx_ref = ref[4:].max(0)
    mx_gpu = raw[0, 4:].max(0)
    margin = 5e-3
    sure = np.abs(mx_ref - 0.4) > margin
    assert np.array_equal((mx_ref > 0.4)[sure], (mx_gpu > 0.4)[sure])
    assert np.array_equal(ref[4:].argmax(0)[sure & (mx_ref > 0.4)], raw[0, 4:].argmax(0)[sure & (mx_ref > 0.4)])
    print(f"[parity] v8l candidate set: {int((mx_ref > 0.4).sum())} oracle candidates, {int((~sure).sum())} inside the {margin} margin")
    eng.close()


@pytest.mark.parametrize("backbone", ["18", "34"])
def test_ufldv2_engine_vs_oracle(backbone):
    path, sd, _ = cached_plan("ufldv2", backbone=backbone)
    eng = _capi.Engine(path, 0, max_batch=2)
    frames = np.stack([synth.frame(s) for s in (0, 1)])
    x = _capi.ufld_preprocess(frames, (320, 1600), 0.6)
    outs = eng.infer(x)
    model = nets.build("ufldv2", sd, backbone=backbone)
    with torch.no_grad():
        ref = [o.numpy() for o in model(torch.from_numpy(x))]
    worst = 0.0
    for name, got, r in zip(("loc_row", "loc_col", "exist_row", "exist_col"), outs, ref):
        assert got.shape == r.shape
        worst = max(worst, _report(f"ufld{backbone} {name}", got, r) / max(1.0, float(np.abs(r).max())))
    assert worst < 5e-3           # logits within 5e-3 of their dynamic range
    # fused lane detect == reference decode applied to the device's own head tensors
    pts, npts, status, coords = eng.ufld_detect(frames, want_coords=True)
    for b in range(2):
        opts, ost, ocrd = post.ufld_decode([o[b:b + 1] for o in outs], 1280, 720, post.CULANE_ROW_ANCHOR, post.CULANE_COL_ANCHOR)
        for l in range(4):
            n = int(npts[b, l])
            assert n == len(opts[l])
            assert np.allclose(coords[b, l, :n], np.array(ocrd[l]), rtol=0, atol=1e-3)
            diff = pts[b, l, :n] - np.array(opts[l], np.int32).reshape(-1, 2)
            for j in np.nonzero(diff.any(axis=1))[0]:
                assert np.abs(diff[j]).max() == 1 and abs(coords[b, l, j] - round(coords[b, l, j])) < 1e-3
        assert [bool(v) for v in status[b]] == ost
    eng.close()


def test_engine_protocol_and_detector_api(tmp_path):
    from adas_b200.ObjectDetector import YoloDetector, ObjectModelType
    from adas_b200.TrafficLaneDetector import UltrafastLaneDetectorV2, LaneModelType
    path, sd, _ = cached_plan("yolov5", scale="n")
    e = B200Engine(path)
    assert e.framework_type == "b200" and e.get_engine_input_shape() == [1, 3, 640, 640]
    shapes, names = e.get_engine_output_shape()
    assert shapes == [[1, 25200, 85]] and names == ["output0"]
    out = e.engine_inference(np.zeros((1, 3, 640, 640), np.float32))
    assert out[0].shape == (1, 25200, 85)
    with pytest.raises(Exception):
        B200Engine(str(tmp_path / "missing.b200w"))
    YoloDetector.set_defaults({"model_path": path, "model_type": ObjectModelType.YOLOV5, "classes_path": None, "box_score": 0.4,
                               "box_nms_iou": 0.45})
    det = YoloDetector(logger=None, max_batch=2)
    fr = synth.frame(3)
    det.DetectFrame(fr)
    single = det.object_info
    both = det.DetectFrames([fr, synth.frame(4)])
    assert [(r.x, r.y, r.width, r.height, r.conf, r.label) for r in single] == [(r.x, r.y, r.width, r.height, r.conf, r.label) for r in both[0]]
    for r in single:
        assert isinstance(r.tolist()[0], int)
    upath, usd, _ = cached_plan("ufldv2", backbone="18")
    lane = UltrafastLaneDetectorV2(upath, LaneModelType.UFLDV2_CULANE, None)
    lane.DetectFrame(fr)
    assert len(lane.lane_info.lanes_points) == 4 and len(lane.lane_info.lanes_status) == 4
    with pytest.raises(Exception):
        UltrafastLaneDetectorV2(upath, LaneModelType.UFLDV2_CURVELANES, None)


def test_yolov5_lite_plan_matches_decoded_plan():
    """ObjectModelType.YOLOV5_LITE end to end: a lite plan (sigmoid-only head + device lite_postprocess) gives the detections of the
    plan whose Detect layer decodes in the graph, bit for bit; engine_inference on the lite plan returns the UNdecoded tensor whose host
    lite_postprocess (oracle restatement of yoloDetector.py:36-50) reproduces the fused call; the model_type / plan pairing is enforced."""
    from adas_b200.ObjectDetector import YoloDetector, ObjectModelType
    path, sd, _ = cached_plan("yolov5", scale="n")
    lpath, _, _ = cached_plan("yolov5", scale="n", lite=True)
    cfg = {"classes_path": None, "box_score": 0.4, "box_nms_iou": 0.45}
    YoloDetector.set_defaults({"model_path": path, "model_type": ObjectModelType.YOLOV5, **cfg})
    det = YoloDetector(logger=None, max_batch=2)
    YoloDetector.set_defaults({"model_path": lpath, "model_type": ObjectModelType.YOLOV5_LITE, **cfg})
    det_l = YoloDetector(logger=None, max_batch=2)
    frames = [synth.frame(3), synth.frame(4)]
    a, b = det.DetectFrames(frames), det_l.DetectFrames(frames)
    key = lambda rs: [(r.x, r.y, r.width, r.height, r.conf, r.label) for r in rs]
    assert sum(len(x) for x in a) > 0
    for ra, rb in zip(a, b):
        assert key(ra) == key(rb)
    # raw tensors: same logits, decoded vs sigmoid-only boxes
    x = _blob(frames)
    raw, raw_l = det.engine.engine_inference(x)[0], det_l.engine.engine_inference(x)[0]
    assert np.array_equal(raw[..., 4:], raw_l[..., 4:]) and raw_l[..., :4].max() <= 1.0 and raw[..., :4].max() > 1.0
    for i in range(2):
        assert np.array_equal(post.yolo_lite_postprocess(raw_l[i]), raw[i])
        r = post.yolo_postprocess(raw_l[i], "v5lite", post.letterbox_geom(720, 1280, 640, 640), 0.4, 0.45)
        assert np.array_equal(np.array([[q.x, q.y, q.width, q.height] for q in b[i]], np.float32).reshape(-1, 4), r["boxes"])
    YoloDetector.set_defaults({"model_path": path, "model_type": ObjectModelType.YOLOV5_LITE, **cfg})
    with pytest.raises(Exception):
        YoloDetector(logger=None)            # lite model_type on a decoded plan
    YoloDetector.set_defaults({"model_path": lpath, "model_type": ObjectModelType.YOLOV5, **cfg})
    with pytest.raises(Exception):
        YoloDetector(logger=None)


@pytest.mark.parametrize("impl", ["native", "python"])
def test_bytetracker_matches_reference_golden(golden_dir, impl):
    from adas_b200.ObjectTracker import BYTETracker, BYTETrackerPy
    g = np.load(os.path.join(golden_dir, "track.npz"))
    cls_ = BYTETracker if impl == "native" else BYTETrackerPy
    for seed, nobj in ((0, 8), (1, 14), (2, 4), (3, 25)):
        trk = cls_(names=[])
        trk.reset()
        rows = []
        for f, (boxes, scores, labels) in enumerate(synth.track_sequence(seed, frames=45, objects=nobj)):
            trk.update(boxes, scores, labels, np.zeros((720, 1280, 3), np.uint8))
            for t in trk.tracked_stracks:
                tl = t.tlwh
                rows.append([f, t.track_id, int(t.is_activated), t.state, tl[0], tl[1], tl[2], tl[3], float(t.score), int(str(t.class_id)[5:])])
            for t in trk.lost_stracks:
                rows.append([f, t.track_id, -1, t.state, 0, 0, 0, 0, 0, -1])
        got, gold = np.array(rows, np.float64), g[f"seq{seed}"]
        assert got.shape == gold.shape, (seed, got.shape, gold.shape)
        assert np.array_equal(got[:, [0, 1, 2, 3, 9]], gold[:, [0, 1, 2, 3, 9]])      # frame, track id, activation, state, class: exact
        assert np.allclose(got[:, 4:9], gold[:, 4:9], rtol=0, atol=1e-6)              # Kalman boxes (fp64) and scores


def test_detect_pair_concurrent_streams_equal_separate_calls():
    """adas_detect_pair enqueues both networks on their own streams before waiting for either; the results must be bit-identical
    to the two stand-alone calls (host frames, and frames already resident on the device), over repeated graph replays."""
    ypath, _, _ = cached_plan("yolov8", scale="l")
    upath, _, _ = cached_plan("ufldv2", backbone="18")
    ye = _capi.Engine(ypath, 0, max_batch=3)
    ue = _capi.Engine(upath, 0, max_batch=3)
    frames = np.stack([synth.frame(s) for s in (21, 22, 23)])
    y_ref = ye.yolo_detect(frames, 0.4, 0.45)
    u_ref = ue.ufld_detect(frames)
    dev = torch.from_numpy(frames).cuda()
    torch.cuda.synchronize()
    for rep in range(4):         # eager -> capture -> replay -> replay
        on_dev = rep % 2 == 1
        y, u = _capi.detect_pair(ye, ue, dev.data_ptr() if on_dev else frames, 0.4, 0.45, 300, on_dev, (3, 720, 1280))
        n = y[4]
        assert np.array_equal(n, y_ref[4]) and np.array_equal(y[5], y_ref[5])
        for b in range(3):
            for k in range(4):
                assert np.array_equal(y[k][b, :n[b]], y_ref[k][b, :n[b]])
        assert np.array_equal(u[1], u_ref[1]) and np.array_equal(u[2], u_ref[2])
        for b in range(3):
            for l in range(4):
                assert np.array_equal(u[0][b, l, :u[1][b, l]], u_ref[0][b, l, :u[1][b, l]])
    assert int(y_ref[4].sum()) > 0
    ye.close()
    ue.close()


def test_pipeline_overlapped_equals_synchronous():
    """AdasPipeline.step_pipelined (two engine pairs in flight, tracker one batch behind) returns, batch for batch, exactly what the
    synchronous step() does: detections, lane points and track ids."""
    from adas_b200.pipeline import AdasPipeline
    ypath, _, _ = cached_plan("yolov8", scale="l")
    upath, _, _ = cached_plan("ufldv2", backbone="18")
    batches = [np.stack([synth.frame(40 + 2 * i + j) for j in range(2)]) for i in range(5)]

    def run(pipelined: bool):
        pipe = AdasPipeline(ypath, upath, device=0, batch=2, sets=2 if pipelined else 1, depth=3)
        out = []
        if pipelined:
            for fr in batches:
                r = pipe.step_pipelined(fr)
                if r is not None:
                    out.append(r)
            out += pipe.flush()
        else:
            out = [pipe.step(fr) for fr in batches]
        pipe.close()
        return out

    a, b = run(False), run(True)
    assert len(a) == len(b) == len(batches)
    dets = 0
    for ra, rb in zip(a, b):
        assert np.array_equal(ra.counts, rb.counts)
        for i, n in enumerate(ra.counts):
            dets += int(n)
            assert np.array_equal(ra.boxes[i, :n], rb.boxes[i, :n]) and np.array_equal(ra.scores[i, :n], rb.scores[i, :n])
            assert np.array_equal(ra.class_ids[i, :n], rb.class_ids[i, :n])
        assert np.array_equal(ra.lane_npts, rb.lane_npts) and np.array_equal(ra.lane_status, rb.lane_status)
        for i in range(ra.lane_pts.shape[0]):
            for l in range(4):
                assert np.array_equal(ra.lane_pts[i, l, :ra.lane_npts[i, l]], rb.lane_pts[i, l, :rb.lane_npts[i, l]])
        key = lambda t: (t["track_id"], t["location"], t["score"], t["class_id"], t["curr_frame_number"], t["is_activated"], t["count"])
        assert [[key(t) for t in fr] for fr in ra.tracks] == [[key(t) for t in fr] for fr in rb.tracks]
    assert dets > 0
