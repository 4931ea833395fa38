#!/bin/bash
# round-2 probe 16: stem_conv.cu (direct stem from the image) -- kernel tests, full suite, stem A/B (YOLO im2col vs direct, UFLD pack vs direct), bench
O=gpurun_out/probe16; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s > $O/pytest_all.txt 2>&1; echo "rc=$?"
grep -E "passed|failed|^E  |FAILED|Timeout|skipped" $O/pytest_all.txt | tail -n 25
for net in yolov8 ufldv2; do
  timeout 600 python tools/op_table.py $net 8 > $O/optable_${net}_direct.txt 2>$O/optable_${net}_direct.err; head -n 4 $O/optable_${net}_direct.txt | cut -c1-150; tail -n 2 $O/optable_${net}_direct.txt
done
ADAS_B200_UFLD_STEM=direct ADAS_B200_PLAN_CACHE=/tmp/plans_direct timeout 600 python tools/op_table.py ufldv2 8 > $O/optable_ufldv2_stemdirect.txt 2>&1; head -n 4 $O/optable_ufldv2_stemdirect.txt | cut -c1-150; tail -n 2 $O/optable_ufldv2_stemdirect.txt
ADAS_B200_STEMCONV=0 ADAS_B200_PLAN_CACHE=/tmp/plans_nostem timeout 600 python tools/op_table.py yolov8 8 > $O/optable_yolov8_im2col.txt 2>&1; head -n 4 $O/optable_yolov8_im2col.txt | cut -c1-150; tail -n 2 $O/optable_yolov8_im2col.txt
for mode in direct im2col direct; do
if [ $mode = im2col ]; then export ADAS_B200_STEMCONV=0 ADAS_B200_PLAN_CACHE=/tmp/plans_nostem; else unset ADAS_B200_STEMCONV ADAS_B200_PLAN_CACHE; fi
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_$mode.json 2>$O/bench_$mode.err; python -c "
import json;d=json.loads(open('$O/bench_$mode.json').read().strip().splitlines()[-1]);print('stem $mode',d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['roofline']['frac'],d['tracks_alive'])"
done
