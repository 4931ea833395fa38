import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import adas_b200
from adas_b200 import _capi
from gpu_util import cached_plan
path, sd, pb = cached_plan("yolov8", scale="l")
B = 8
eng = _capi.Engine(path, 0, max_batch=B)
ms, n = eng.time_ops(B, 1 << 1, 5)
print("dbg", os.environ.get("ADAS_B200_DBG"), f"gemm only: {ms:.3f} ms/pass ({n} launches) {pb.flops_per_img*B/1e9/ms:.1f} TFLOP/s")
