#!/bin/bash
# round-2 probe 32: batches in flight -- engine pairs (sets) x queue depth
O=gpurun_out/probe32; mkdir -p $O
for cfg in "2 3" "3 4" "2 3" "3 4" "4 5" "1 2"; do set -- $cfg
timeout 600 python bench.py --steps 100 --warmup 5 --cpu-frames 0 --other-configs 0 --sets $1 --depth $2 > $O/bench_s$1_d$2.json 2>$O/bench.err; python -c "
import json;d=json.loads(open('$O/bench_s$1_d$2.json').read().strip().splitlines()[-1]);print('sets $1 depth $2',d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'])"
done
