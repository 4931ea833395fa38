"""Scratch timing script for the first GPU runs (not the contract bench)."""
import sys, time, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import adas_b200
from adas_b200 import _capi
from gpu_util import cached_plan
import synth
for kind, kw, B in (("yolov8", dict(scale="l"), 8), ("ufldv2", dict(backbone="34"), 8), ("yolov5", dict(scale="n"), 8)):
    t = time.time(); path, sd, pb = cached_plan(kind, **kw); print(kind, "plan build s", round(time.time()-t,1), flush=True)
    for impl in (0,):
        eng = _capi.Engine(path, 0, max_batch=B, conv_impl=impl)
        for mask, name in ((0xFFFFFFFF, "all ops"), (1 << 1, "gemm only"), (1 << 2, "im2col only")):
            ms, n = eng.time_ops(B, mask, 5)
            gf = pb.flops_per_img * B / 1e9
            print(f"{kind} B={B} impl={impl} {name}: {ms:.3f} ms/pass ({n} launches), {gf/ms:.1f} TFLOP/s -> {B/ms*1e3:.0f} img/s", flush=True)
        eng.close()
