#!/bin/bash
# round-2 probe 18: association kernel through mapped host memory (no copies / no stream sync per frame) -- tracker tests, bench x2
O=gpurun_out/probe18; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k "track or frames_to or comm or pipeline" > $O/pytest_track.txt 2>&1; tail -n 3 $O/pytest_track.txt
for i in 1 2; do
timeout 600 python bench.py --steps 100 --warmup 5 --cpu-frames 0 --other-configs 0 > $O/bench_$i.json 2>$O/bench_$i.err; python -c "
import json;d=json.loads(open('$O/bench_$i.json').read().strip().splitlines()[-1]);print('bench',d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['roofline']['frac'],d['tracks_alive'])"
done
ADAS_B200_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --cpu-frames 0 --other-configs 0 2>&1 | grep -i "trace" | tail -n 5
