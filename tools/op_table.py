"""Per-layer table of a plan at real clocks: every launch replayed alone (CUDA events, L2-warm), GEMM shapes and tile choices.
usage: python tools/op_table.py yolov8|ufldv2|yolov5 [batch] [iters]   (env switches of the library apply)"""
import os, re, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import adas_b200
from adas_b200 import _capi
from gpu_util import cached_plan
kind = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 8; iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
kw = {"yolov8": dict(scale="l"), "ufldv2": dict(backbone="34"), "yolov5": dict(scale="n")}[kind]
path, sd, pb = cached_plan(kind, **kw)
eng = _capi.Engine(path, 0, max_batch=B)
eng.run(B)
n = eng.num_steps(B)
names = {1: "gemm", 2: "im2col", 3: "maxpool", 4: "upsample", 5: "layernorm", 6: "stempack", 7: "stemconv", 31: "(folded)"}
tot = 0.0; tot_g = 0.0; rows = []
for i in range(n):
    ms, t, d = eng.time_step(B, i, iters)
    tot += ms
    tf = ""
    if t == 1:
        tot_g += ms
        m = re.match(r"M=(\d+) N=(\d+) K=(\d+)", d)
        if m:
            M, N, K = map(int, m.groups())
            tf = f"{2.0 * M * N * K / ms / 1e9:7.1f} TF(incl halo)"
    rows.append((ms, i, names.get(t, str(t)), d, tf))
    print(f"{i:3d} {names.get(t, str(t)):9s} {ms * 1e3:8.1f} us {tf} {d}", flush=True)
flops = (pb.flops_per_img - pb.stem_flops_per_img) * B      # the stem conv is not a GEMM launch
print(f"TOTAL {kind} b{B}: sum of isolated launches {tot * 1e3:.1f} us (gemm {tot_g * 1e3:.1f} us) -> {flops / tot_g / 1e9:.1f} TFLOP/s algorithmic over GEMM time")
ms_all, nl = eng.time_ops(B, 0xFFFFFFFF, 10)
ms_g, ng = eng.time_ops(B, 1 << 1, 10)
print(f"back-to-back: all {ms_all * 1e3:.1f} us ({nl} launches), gemm only {ms_g * 1e3:.1f} us ({ng}) -> {flops / ms_g / 1e9:.1f} TFLOP/s")
eng.close()
