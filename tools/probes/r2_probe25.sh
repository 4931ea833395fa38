#!/bin/bash
# round-2 probe 25: third staging buffer where it costs no operand stage (one TMA store may still read while the next chunk is written)
O=gpurun_out/probe25; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 > $O/pytest_kernels.txt 2>&1; tail -n 2 $O/pytest_kernels.txt
for net in yolov8 ufldv2; do
  timeout 600 python tools/op_table.py $net 8 > $O/optable_${net}_stg3.txt 2>&1; tail -n 2 $O/optable_${net}_stg3.txt
  ADAS_B200_STG2=1 timeout 600 python tools/op_table.py $net 8 > $O/optable_${net}_stg2.txt 2>&1; tail -n 2 $O/optable_${net}_stg2.txt
done
for i in 1 2; do for mode in stg3 stg2; do
if [ $mode = stg2 ]; then export ADAS_B200_STG2=1; else unset ADAS_B200_STG2; fi
timeout 600 python bench.py --steps 100 --warmup 5 --cpu-frames 0 --other-configs 0 > $O/bench_${mode}_$i.json 2>$O/bench_${mode}_$i.err; python -c "
import json;d=json.loads(open('$O/bench_${mode}_$i.json').read().strip().splitlines()[-1]);print('$mode',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'])"
done; done
