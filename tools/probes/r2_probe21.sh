#!/bin/bash
# round-2 probe 21: lane geometry kernel (rows K + 8f-1) and the ONNX engine test
O=gpurun_out/probe21; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s -k "lane or birdview or onnx or area" > $O/pytest_lane.txt 2>&1; echo "rc=$?"
grep -E "passed|failed|^E  |FAILED|Timeout|skipped|parity\]|lane geometry\]" $O/pytest_lane.txt | tail -n 30
