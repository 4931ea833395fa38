"""Target for compute-sanitizer (memcheck / racecheck): a few small launches of every hand-written kernel family through the C ABI --
the warp-specialised tcgen05 conv kernel (slab, plain, stride-2, staged and direct epilogues), a chain launch (gemm_chain.cu: four
same-shape layers with residuals in one persistent launch), the stem conv (stem_conv.cu), FC stream, NMS, lane decode, the fused
three-stage association through mapped host memory.  usage: compute-sanitizer --tool racecheck --target-processes all python tools/sanitize_target.py"""
import os, sys, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
os.environ.setdefault("ADAS_B200_AUTOTUNE", "0")
os.environ.setdefault("ADAS_B200_NO_GRAPH", "1")
import numpy as np
import adas_b200
from adas_b200 import _capi, plan
import synth
from gpu_util import to_padded

rng = np.random.default_rng(0)
CASES = [  # B cin cout H W k s act residual tile
    (1, 64, 128, 24, 24, 3, 1, 1, "post", (128, 2)),
    (1, 64, 64, 16, 24, 3, 1, 2, "pre", (64, 3)),
    (1, 128, 256, 16, 16, 1, 1, 1, None, (256, 2)),
    (1, 64, 128, 32, 32, 3, 2, 1, None, (128, 2)),
    (1, 64, 80, 16, 16, 1, 1, 0, None, (80, 1)),
]
for i, (B, cin, cout, H, W, k, s, act, res, tile) in enumerate(CASES):
    pb = plan.PlanBuilder(plan.MODEL_YOLOV5, 3, H, W)
    xin = pb.new_padded(H, W, cin)
    w = (rng.standard_normal((cout, cin, k, k)) * 0.05).astype(np.float32)
    pd = k // 2
    Ho, Wo = (H + 2 * pd - k) // s + 1, (W + 2 * pd - k) // s + 1
    rv = pb.new_padded(Ho, Wo, cout) if res else None
    out = pb.conv(xin, w, np.zeros(cout, np.float32), k, s, act, res=rv, res_pre_act=(res == "pre"), tile=tile, pad=pd)
    path = os.path.join(tempfile.gettempdir(), f"san_{i}.b200w")
    pb.write(path)
    eng = _capi.Engine(path, 0, max_batch=B)
    eng.write_buffer(xin.buf, to_padded(rng.standard_normal((B, cin, H, W)).astype(np.float32), cin))
    eng.run(B)
    eng.read_buffer(out.buf, B)
    eng.close()
    print("conv case", i, "ok", flush=True)
# chain launch (EXPERIMENTAL kernel, off by default in the engine): four 3x3 64->64 layers, every second one with a residual = the input of the layer before it (C2f bottleneck pattern)
os.environ["ADAS_B200_CHAIN"] = "1"
pb = plan.PlanBuilder(plan.MODEL_YOLOV5, 3, 24, 24)
x0 = pb.new_padded(24, 24, 64)
cur = x0
for li in range(2):
    w1 = (rng.standard_normal((64, 64, 3, 3)) * 0.05).astype(np.float32)
    w2 = (rng.standard_normal((64, 64, 3, 3)) * 0.05).astype(np.float32)
    t = pb.conv(cur, w1, np.zeros(64, np.float32), 3, 1, 1)
    cur = pb.conv(t, w2, np.zeros(64, np.float32), 3, 1, 1, res=cur)
path = os.path.join(tempfile.gettempdir(), "san_chain.b200w"); pb.write(path)
eng = _capi.Engine(path, 0, max_batch=2)
descs = [eng.time_step(2, i, 1)[2] for i in range(eng.num_steps(2))]
assert any("chain of" in d for d in descs), descs
eng.write_buffer(x0.buf, to_padded(rng.standard_normal((2, 64, 24, 24)).astype(np.float32), 64))
eng.run(2); eng.run(2)
eng.read_buffer(cur.buf, 2)
eng.close()
os.environ.pop("ADAS_B200_CHAIN", None)
print("chain ok", flush=True)
# stem conv straight from the image (3x3 and 7x7, the latter reaching beyond the one-pixel halo)
for k, pad, cout in ((3, 1, 64), (7, 3, 64), (6, 2, 16)):
    pb = plan.PlanBuilder(plan.MODEL_YOLOV8, 3, 36, 50)
    out = pb.conv(pb.image, (rng.standard_normal((cout, 3, k, k)) * 0.1).astype(np.float32), np.zeros(cout, np.float32), k, 2, 1, pad=pad)
    path = os.path.join(tempfile.gettempdir(), f"san_stem{k}.b200w"); pb.write(path)
    eng = _capi.Engine(path, 0, max_batch=2)
    eng.write_buffer(pb.image.buf, to_padded(rng.standard_normal((2, 3, 36, 50)).astype(np.float32), 4))
    eng.run(2); eng.read_buffer(out.buf, 2); eng.close()
print("stem ok", flush=True)
# FC stream + swap-AB FC
pb = plan.PlanBuilder(plan.MODEL_UFLDV2, 3, 8, 8)
xi = pb.new_dense(1, 512); h = pb.new_dense(1, 256); o = pb.new_dense(1, 136, f32=True)
pb.fc(xi, 512, (rng.standard_normal((256, 512)) * 0.05).astype(np.float32), np.zeros(256, np.float32), 2, h)
pb.fc(h, 256, (rng.standard_normal((136, 256)) * 0.05).astype(np.float32), np.zeros(136, np.float32), 0, o)
path = os.path.join(tempfile.gettempdir(), "san_fc.b200w"); pb.write(path)
eng = _capi.Engine(path, 0, max_batch=3); eng.write_buffer(xi, rng.standard_normal((3, 512)).astype(np.float16)); eng.run(3); eng.close()
print("fc ok", flush=True)
# post-processing kernels and the tracker
raw = np.stack([synth.yolo_v8_head(s, n_hot=60) for s in (0, 1)])
_capi.yolo_postprocess(raw, 0, 80, (640, 640), (720, 1280), 0.4, 0.45)
heads = np.stack([np.concatenate([x.ravel() for x in synth.ufld_heads(0)])])
from oracle import post
_capi.ufld_postprocess(heads, (200, 72, 100, 81, 4), (1280, 720), post.CULANE_ROW_ANCHOR, post.CULANE_COL_ANCHOR)
trk = _capi.NativeTracker(0)
trk.reset()
for boxes, scores, labels in synth.track_sequence(0, frames=6, objects=6):
    trk.update(np.asarray(boxes, float), scores, np.asarray([int(str(l)[5:]) if isinstance(l, str) else int(l) for l in labels], np.int32))
print("post + tracker ok", flush=True)
