#!/bin/bash
# round-2 probe 30: what the driver runs at round end -- GPU suite, smoke, bench (both arms), on the final build
O=gpurun_out/probe30; mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 1 $O/smoke.txt
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_reference.json 2>$O/bench_reference.err; echo "ref rc=$?"; cut -c1-300 $O/bench_reference.json
timeout 900 python bench.py > $O/bench_n1.json 2>$O/bench_n1.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]);print('bench',d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['host_tracker_breakdown_ms_per_step'],d['roofline']['frac'],d['gpu_launches'],d['cpu_baseline']['value'],d['cpu_baseline']['cores'],[ (k,v['frac_of_peak']) for k,v in d['other_configs'].items()], d['clocks'])"
