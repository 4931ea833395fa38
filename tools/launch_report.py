"""Summarise an ncu launch-list CSV (gpu__time_duration.sum per launch) against the plan's GEMM ops.
usage: python tools/launch_report.py <csv> yolov8|ufldv2|yolov5 <batch>"""
import csv, sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import adas_b200
from adas_b200 import plan
fn, kind, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = [(r[4].split('(')[0], r[8], float(r[14]) / 1e3) for r in csv.reader(open(fn)) if len(r) >= 15 and r[0].isdigit()]
W = plan.synth_weights(kind, 0)
pb = {"yolov8": lambda: plan.build_yolov8(W, "l"), "ufldv2": lambda: plan.build_ufldv2(W, "34"), "yolov5": lambda: plan.build_yolov5(W, "n")}[kind]()
tot = sum(r[2] for r in rows)
print(f"{fn}: {len(rows)} launches, {tot:.1f} us total")
by = {}
for r in rows:
    by[r[0]] = by.get(r[0], 0) + r[2]
for k, v in sorted(by.items(), key=lambda x: -x[1]):
    print(f"  {k:28s} {v:9.1f} us {100 * v / tot:5.1f}%")
out = []
for (typ, p, f), (name, grid, us) in zip(pb.ops, rows):
    if typ == 1:
        a_buf, a_coff, Kc, ntaps, wt, bt, N, act, rb, rc, rp, ob, oc, masked, tr, BN = p[:16]
        M = B * pb.buffers[ob if p[16] else a_buf][0] if not tr else N
        Nn = N if not tr else B
        out.append((us, M, Nn, Kc * ntaps, ntaps, grid, 2 * M * Nn * Kc * ntaps / us / 1e6))
print("  GEMM launches by time: us, M, N, K, taps, grid, TFLOP/s (incl. halo rows)")
for o in sorted(out, key=lambda x: -x[0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    print("   %8.1f  M=%7d N=%4d K=%5d taps=%d grid=%-14s %7.1f" % o)
# group by (M,N,K)
grp = {}
for o in out:
    k = (o[1], o[2], o[3], o[4])
    g = grp.setdefault(k, [0, 0.0])
    g[0] += 1; g[1] += o[0]
print("  grouped: count, total us, M, N, K, taps")
for k, g in sorted(grp.items(), key=lambda x: -x[1][1])[:25]:
    print("   %3d %8.1f  M=%7d N=%4d K=%5d taps=%d  %7.1f TFLOP/s" % (g[0], g[1], k[0], k[1], k[2], k[3], 2 * g[0] * k[0] * k[1] * k[2] / g[1] / 1e6))
