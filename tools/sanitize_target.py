"""Target for compute-sanitizer (memcheck / racecheck): a few small launches of every hand-written kernel family through the C ABI --
the warp-specialised tcgen05 conv kernel (slab, plain, stride-2, staged and direct epilogues), FC stream, NMS, lane decode, the fused
three-stage association.  usage: compute-sanitizer --tool racecheck --target-processes all python tools/sanitize_target.py"""
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
trk = _capi.NativeTracker(0)
trk.reset()
for boxes, scores, labels in synth.track_sequence(0, frames=6, objects=6):
    trk.update(np.asarray(boxes, float), scores, np.asarray([int(str(l)[5:]) if isinstance(l, str) else int(l) for l in labels], np.int32))
print("post + tracker ok", flush=True)
