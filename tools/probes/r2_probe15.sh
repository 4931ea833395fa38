#!/bin/bash
# round-2 probe 15: full GPU suite after the GEMM cleanup / TuSimple / plan validation changes
O=gpurun_out/probe15; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s > $O/pytest_all.txt 2>&1; echo "rc=$?"
grep -E "^\[parity\]|passed|failed|^E  |FAILED|Timeout|skipped" $O/pytest_all.txt | tail -n 40
