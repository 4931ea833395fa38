#!/bin/bash
# round-2 probe 20: full GPU suite, stem timing, sanitizers over every kernel family (chain, stem, mapped-memory association included)
O=gpurun_out/probe20; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s > $O/pytest_all.txt 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^E  |FAILED|Timeout|skipped|onnx-plan" $O/pytest_all.txt | tail -n 12
timeout 600 python tools/op_table.py yolov8 8 > $O/optable_yolov8.txt 2>$O/optable_yolov8.err; head -n 1 $O/optable_yolov8.txt | cut -c1-150; tail -n 2 $O/optable_yolov8.txt
python tools/sanitize_target.py > $O/san_plain.txt 2>&1; tail -n 3 $O/san_plain.txt
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --target-processes all --log-file $O/$tool.txt python tools/sanitize_target.py > $O/${tool}_run.txt 2>&1
  echo "== $tool rc=$?"; tail -n 3 $O/$tool.txt
done
