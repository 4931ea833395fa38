#!/bin/bash
# round-2 probe 6: fused tracker association + batched tracker call, FC stream rewrite, SiLU numerics A/B for the 1e-3 contract
O=gpurun_out/probe6; mkdir -p $O
for v in "V3=1" "ADAS_B200_DBG=128" "ADAS_B200_GEMM=v2"; do
  echo "== $v" >> $O/parity_ab.txt
  env $v timeout 600 python -m pytest tests/test_gpu_nets.py -m gpu -q -s -k "yolov8l_engine_vs_oracle or yolov5n_engine" 2>&1 | grep -E "parity|passed|failed|assert" >> $O/parity_ab.txt
done
cat $O/parity_ab.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 -s > $O/pytest_all.txt 2>&1
echo "all rc=$?" >> $O/pytest_all.txt
grep -E "parity|passed|failed|Error|assert|^E " $O/pytest_all.txt | tail -40
python tools/op_table.py ufldv2 8 > $O/optable_ufld_b8.txt 2>$O/optable_ufld_b8.err; tail -n 4 $O/optable_ufld_b8.txt | cut -c1-120
python bench.py --steps 20 --warmup 5 > $O/bench.json 2>$O/bench.err; python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['roofline']['frac'])"
