"""Target for ncu: replays one plan eagerly (ADAS_B200_NO_GRAPH=1) so every kernel is a separate launch.
usage: python tools/profile_target.py yolov8|ufldv2|yolov5 [batch] [passes]"""
import os, sys
os.environ["ADAS_B200_NO_GRAPH"] = "1"
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import adas_b200
from adas_b200 import _capi
from gpu_util import cached_plan
kind = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 8; passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kw = {"yolov8": dict(scale="l"), "ufldv2": dict(backbone="34"), "yolov5": dict(scale="n")}[kind]
path, sd, pb = cached_plan(kind, **kw)
eng = _capi.Engine(path, 0, max_batch=B)
eng.run(B)                                  # autotune + first touch happen outside the profiled range
if os.environ.get("ADAS_B200_PROFILE_RANGE"):
    import torch
    torch.cuda.profiler.start()
for _ in range(passes):
    eng.run(B)
if os.environ.get("ADAS_B200_PROFILE_RANGE"):
    torch.cuda.profiler.stop()
print("launches", _capi.launch_count(), "ops per pass", len(pb.ops))
