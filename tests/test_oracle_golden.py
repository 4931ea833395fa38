"""CPU: the oracle restatement (oracle/post.py) against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Bit-exact for indices / class ids / boxes (float32) / lane points."""
import hashlib
import os

import numpy as np
import pytest

import synth
from oracle import post


def _sha(a):
    return np.frombuffer(bytes.fromhex(hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()), np.uint8)


def test_nms_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "nms.npz"))
    n_dup = 0
    for key in g.files:
        if key in ("n0", "n1"):
            continue
        s, n, t = key.split("_")
        b, c = synth.nms_case(int(s[1:]), int(n[1:]))
        keep = post.soft_nms(b, c, float(t[1:]))
        assert np.array_equal(keep, g[key]), key
        n_dup += len(keep) != len(set(keep.tolist()))
    assert n_dup > 5          # the duplicate-emitting swap path is exercised
    assert np.array_equal(post.soft_nms(np.zeros((0, 4), np.float32), [], 0.45), g["n0"])
    assert np.array_equal(post.soft_nms(np.array([[1, 2, 3, 4]], np.float32), [0.9], 0.45), g["n1"])


@pytest.mark.parametrize("key,kind", [("v8_s0_720x1280", "v8"), ("v8_s0_480x640", "v8"), ("v8_s1_720x1280", "v8"), ("v8_s2_720x1280", "v8"),
                                      ("v8_s3_720x1280", "v8"), ("v5_s10_720x1280", "v5"), ("v5_s11_720x1280", "v5")])
def test_yolo_post_matches_reference(golden_dir, key, kind):
    g = np.load(os.path.join(golden_dir, "yolo_post.npz"))
    seed = int(key.split("_")[1][1:])
    h, w = [int(v) for v in key.split("_")[2].split("x")]
    raw = synth.yolo_v8_head(seed) if kind == "v8" else synth.yolo_v5_head(seed)
    geom = post.letterbox_geom(h, w, 640, 640)
    r = post.yolo_postprocess(raw, kind, geom, 0.4, 0.45)
    assert np.array_equal(r["boxes"], g[key + "_box"])
    assert np.array_equal(r["scores"].astype(np.float64), g[key + "_conf"])
    assert np.array_equal(r["cls"], g[key + "_cls"])
    if kind == "v8":
        blob, _ = post.yolo_prepare_input(synth.frame(seed, h, w), 640, 640)
        assert np.array_equal(_sha(blob), g[key + "_blob_sha"])


@pytest.mark.parametrize("key", ["s0_720x1280", "s0_480x640", "s1_720x1280", "s2_720x1280", "s3_720x1280"])
def test_ufld_post_matches_reference(golden_dir, key):
    g = np.load(os.path.join(golden_dir, "ufld_post.npz"))
    seed = int(key.split("_")[0][1:])
    h, w = [int(v) for v in key.split("_")[1].split("x")]
    inval = {0: (), 1: (1,), 2: (0, 3), 3: ()}[seed]
    heads = synth.ufld_heads(seed, invalid_lanes=inval)
    pts, status, _ = post.ufld_decode(heads, w, h, post.CULANE_ROW_ANCHOR, post.CULANE_COL_ANCHOR)
    for l in range(4):
        assert np.array_equal(np.array(pts[l], np.int32).reshape(-1, 2), g[f"{key}_lane{l}"]), (key, l)
    assert np.array_equal(np.array(status, np.uint8), g[key + "_status"])
    blob = post.ufld_prepare_input(synth.frame(seed, h, w), 320, 1600, 0.6)
    assert np.array_equal(_sha(blob), g[key + "_blob_sha"])


@pytest.mark.parametrize("key", ["s0_720x1280", "s0_480x640", "s1_720x1280", "s2_720x1280"])
def test_ufld_tusimple_post_matches_reference(golden_dir, key):
    """ModelConfig.init_tusimple_config (ultrafastLaneDetectorV2.py:31-37): 320x800 input, crop 0.8, 56 / 41 anchors."""
    g = np.load(os.path.join(golden_dir, "ufld_post_tusimple.npz"))
    seed = int(key.split("_")[0][1:])
    h, w = [int(v) for v in key.split("_")[1].split("x")]
    inval = {0: (), 1: (2,), 2: (0, 3)}[seed]
    heads = synth.ufld_heads(seed, ngr=100, ncr=56, ngc=100, ncc=41, invalid_lanes=inval)
    pts, status, _ = post.ufld_decode(heads, w, h, post.TUSIMPLE_ROW_ANCHOR, post.TUSIMPLE_COL_ANCHOR)
    for l in range(4):
        assert np.array_equal(np.array(pts[l], np.int32).reshape(-1, 2), g[f"{key}_lane{l}"]), (key, l)
    assert np.array_equal(np.array(status, np.uint8), g[key + "_status"])
    blob = post.ufld_prepare_input(synth.frame(seed, h, w), 320, 800, 0.8)
    assert np.array_equal(_sha(blob), g[key + "_blob_sha"])


@pytest.mark.parametrize("ds,G,R", [("tusimple", 100, 56), ("culane", 200, 18)])
def test_ufld_v1_post_matches_reference(golden_dir, ds, G, R):
    """UFLD v1 decode (ultrafastLaneDetector.py:97-136) and pre-processing (80-95) against the reference detector's own outputs."""
    g = np.load(os.path.join(golden_dir, "ufld_v1_post.npz"))
    for seed, inval in ((0, ()), (1, (2,)), (2, (0, 3))):
        for (h, w) in ((720, 1280), (480, 640)) if seed == 0 else ((720, 1280),):
            key = f"{ds}_s{seed}_{h}x{w}"
            pts, status, _ = post.ufld_v1_decode(synth.ufld_v1_head(seed, G, R, invalid_lanes=inval)[0], post.UFLD_V1[ds], 800, 288, w, h)
            for l in range(4):
                assert np.array_equal(np.array(pts[l], np.int32).reshape(-1, 2), g[f"{key}_lane{l}"]), (key, l)
            assert np.array_equal(np.array(status, np.uint8), g[key + "_status"]), key
            assert np.array_equal(_sha(post.ufld_prepare_input(synth.frame(seed, h, w), 288, 800, 1.0)), g[key + "_blob_sha"]), key


def test_association_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "track.npz"))
    for k in range(7):
        a, b, sc = g[f"assoc{k}_a"], g[f"assoc{k}_b"], g[f"assoc{k}_sc"]
        cost = post.iou_cost(a, b)
        fused = post.iou_cost(a, b, sc)
        assert np.array_equal(cost, g[f"assoc{k}_cost"])
        assert np.allclose(fused, g[f"assoc{k}_fused"], rtol=0, atol=1e-15)
        for nm, c, th in (("iou", cost, 0.5), ("fuse", fused, 0.8), ("fuse7", fused, 0.7)):
            x, _, _ = post.lapjv_extended(c, th)
            assert np.array_equal(x, g[f"assoc{k}_{nm}_x"]), (k, nm)


@pytest.mark.parametrize("key", ["v5lite_s20_720x1280", "v5lite_s20_480x640", "v5lite_s21_720x1280", "v5lite_s22_720x1280"])
def test_yolo_lite_post_matches_reference(golden_dir, key):
    """ObjectModelType.YOLOV5_LITE -- lite_postprocess (yoloDetector.py:36-50) + __process_output + NMS, golden vectors
    from the reference's own YoloDetector on a sigmoid-only head."""
    g = np.load(os.path.join(golden_dir, "yolo_lite.npz"))
    seed = int(key.split("_")[1][1:])
    h, w = [int(v) for v in key.split("_")[2].split("x")]
    raw = synth.yolo_v5_lite_head(seed)
    r = post.yolo_postprocess(raw, "v5lite", post.letterbox_geom(h, w, 640, 640), 0.4, 0.45)
    assert np.array_equal(r["boxes"], g[key + "_box"])
    assert np.array_equal(r["scores"].astype(np.float64), g[key + "_conf"])
    assert np.array_equal(r["cls"], g[key + "_cls"])
    assert len(r["idx"]) != len(set(r["idx"].tolist())) or len(r["idx"]) < r["n_cand"]       # the NMS really suppresses / duplicates


def test_yolo_lite_decode_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "yolo_lite.npz"))
    dec = post.yolo_lite_postprocess(synth.yolo_v5_lite_head(20))
    assert np.array_equal(_sha(dec), g["decoded_s20_sha"])
    assert np.array_equal(dec[::97, :6], g["decoded_s20_sample"])
