#!/bin/bash
# round-2 probe 22: full GPU suite with the lane-existence operating point; bench
O=gpurun_out/probe22; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s > $O/pytest_all.txt 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^E  |FAILED|Timeout|skipped|lane coordinates|frames ->|lane geometry\]" $O/pytest_all.txt | tail -n 14
timeout 900 python bench.py > $O/bench.json 2>$O/bench.err; python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('bench',d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['roofline']['frac'],d['tracks_alive'],d['cpu_baseline'],d['other_configs'])"
