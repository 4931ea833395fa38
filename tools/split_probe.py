"""Does splitting a batch over several streams (smaller kernels, back-filled SMs) beat one batch-8 program?  Scratch."""
import sys, time, os, threading
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import adas_b200
from adas_b200 import _capi
from gpu_util import cached_plan

ypath, _, _ = cached_plan("yolov8", scale="l")
upath, _, _ = cached_plan("ufldv2", backbone="34")
IT = 40

def loop(eng, b, n):
    for _ in range(n):
        eng.run(b)

def timed(engs_b):
    for e, b in engs_b:
        loop(e, b, 3)
    ts = [threading.Thread(target=loop, args=(e, b, IT)) for e, b in engs_b]
    t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    return (time.perf_counter() - t0) / IT * 1e3

y8 = _capi.Engine(ypath, 0, max_batch=8)
u8 = _capi.Engine(upath, 0, max_batch=8)
print("yolo b8 alone       ms", round(timed([(y8, 8)]), 3), flush=True)
print("ufld b8 alone       ms", round(timed([(u8, 8)]), 3), flush=True)
print("yolo b8 | ufld b8   ms", round(timed([(y8, 8), (u8, 8)]), 3), flush=True)
y4 = [_capi.Engine(ypath, 0, max_batch=4) for _ in range(2)]
u4 = [_capi.Engine(upath, 0, max_batch=4) for _ in range(2)]
print("yolo 2x b4          ms", round(timed([(e, 4) for e in y4]), 3), flush=True)
print("ufld 2x b4          ms", round(timed([(e, 4) for e in u4]), 3), flush=True)
print("yolo 2x b4 | ufld 2x b4 ms", round(timed([(e, 4) for e in y4] + [(e, 4) for e in u4]), 3), flush=True)
print("yolo 2x b4 | ufld b8 ms", round(timed([(e, 4) for e in y4] + [(u8, 8)]), 3), flush=True)
y2 = [_capi.Engine(ypath, 0, max_batch=2) for _ in range(4)]
print("yolo 4x b2          ms", round(timed([(e, 2) for e in y2]), 3), flush=True)
print("yolo 4x b2 | ufld b8 ms", round(timed([(e, 2) for e in y2] + [(u8, 8)]), 3), flush=True)
