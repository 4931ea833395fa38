#!/bin/bash
# round-2 probe 8: e2e frames test, sanitizer runs (memcheck / racecheck / synccheck) on every kernel family
O=gpurun_out/probe8; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_nets.py -m gpu -q --timeout 600 -s -k "frames_to or pipeline or bytetracker or lite" > $O/pytest_nets.txt 2>&1
grep -E "parity\]|passed|failed|^E  |FAILED" $O/pytest_nets.txt | tail -20
python tools/sanitize_target.py > $O/san_plain.txt 2>&1; tail -n 3 $O/san_plain.txt
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --log-file $O/$tool.txt python tools/sanitize_target.py > $O/${tool}_run.txt 2>&1
  echo "== $tool rc=$?"; tail -n 4 $O/$tool.txt
done
