#!/bin/bash
# round-2 probe 28: UFLD v1 (plan, engine, decode kernel, wrapper) + full suite
O=gpurun_out/probe28; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s -k "v1" > $O/pytest_v1.txt 2>&1; echo "v1 rc=$?"
grep -E "passed|failed|^E  |FAILED|parity\]" $O/pytest_v1.txt | tail -n 14
true
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
