#!/bin/bash
# round-2 probe 17: pipelined stem_conv.cu -- kernel test, op table head, bench A/B (direct stem vs im2col route), 100 steps, alternating
O=gpurun_out/probe17; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "stem" > $O/pytest_stem.txt 2>&1; tail -n 2 $O/pytest_stem.txt
timeout 600 python tools/op_table.py yolov8 8 > $O/optable_yolov8.txt 2>$O/optable_yolov8.err; head -n 2 $O/optable_yolov8.txt | cut -c1-150; tail -n 2 $O/optable_yolov8.txt
for i in 1 2 3; do for mode in direct im2col; do
if [ $mode = im2col ]; then export ADAS_B200_STEMCONV=0 ADAS_B200_PLAN_CACHE=/tmp/plans_nostem; else unset ADAS_B200_STEMCONV ADAS_B200_PLAN_CACHE; fi
timeout 600 python bench.py --steps 100 --warmup 5 --cpu-frames 0 --other-configs 0 > $O/bench_${mode}_$i.json 2>$O/bench_${mode}_$i.err; python -c "
import json;d=json.loads(open('$O/bench_${mode}_$i.json').read().strip().splitlines()[-1]);print('stem $mode',d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['roofline']['frac'],d['clocks']['sm_mhz'],d['clocks_e2e']['sm_mhz'])"
done; done
