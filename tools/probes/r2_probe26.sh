#!/bin/bash
# round-2 probe 26: warp_perspective kernel (bit-exact vs cv2 / the reference's golden hashes), timing of the batch-8 bird view
O=gpurun_out/probe26; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s -k "warp or bird_view or lane_geometry" > $O/pytest_warp.txt 2>&1; echo "rc=$?"
grep -E "passed|failed|^E  |FAILED" $O/pytest_warp.txt | tail -n 10
python - <<'PY'
import sys, time; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, adas_b200, synth
from adas_b200 import _capi
from adas_b200.TrafficLaneDetector.ufldDetector.perspectiveTransformation import PerspectiveTransformation
import cv2
fr = np.stack([synth.frame(s) for s in range(8)])
t = PerspectiveTransformation((1280, 720))
_capi.warp_perspective(fr, t.M, (1280, 720))
t0 = time.perf_counter()
for _ in range(5): _capi.warp_perspective(fr, t.M, (1280, 720))
print("device warp of 8 frames incl. H2D+D2H: %.2f ms per call" % ((time.perf_counter() - t0) / 5 * 1e3))
t0 = time.perf_counter()
for f in fr: cv2.warpPerspective(f, t.M, (1280, 720), flags=cv2.INTER_LINEAR)
print("cv2 (host) 8 frames: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
PY
