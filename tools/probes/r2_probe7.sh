#!/bin/bash
# round-2 probe 7: parity operating point (gain 16), vectorised tracker hand-off, C-driven NCCL gather
O=gpurun_out/probe7; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 -s > $O/pytest_all.txt 2>&1
echo "all rc=$?" >> $O/pytest_all.txt
grep -E "parity\]|passed|failed|^E  |FAILED" $O/pytest_all.txt | tail -40
python bench.py --steps 20 --warmup 5 > $O/bench.json 2>$O/bench.err; python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['roofline']['frac'],d['tracks_alive'])"
tail -n 3 $O/bench.err
