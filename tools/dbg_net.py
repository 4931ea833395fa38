"""Whole-network graph-replay time under the GEMM debug bits (ADAS_B200_DBG): scratch decomposition tool."""
import sys, time, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import adas_b200
from adas_b200 import _capi
from gpu_util import cached_plan
for kind, kw in (("yolov8", dict(scale="l")), ("ufldv2", dict(backbone="34"))):
    path, _, _ = cached_plan(kind, **kw)
    B = int(os.environ.get("B", "8"))
    e = _capi.Engine(path, 0, max_batch=B)
    for _ in range(4): e.run(B)
    t0 = time.perf_counter()
    for _ in range(40): e.run(B)
    print(f"dbg={os.environ.get('ADAS_B200_DBG','0'):>3} {kind} b{B} graph replay {(time.perf_counter()-t0)/40*1e3:.3f} ms", flush=True)
    e.close()
