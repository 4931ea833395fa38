#!/bin/bash
# round-2 probe 12: chain launches kept only where they time faster (default) -- decisions, op tables, bench A/B in one box
O=gpurun_out/probe12; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q --timeout 300 -s -k "chain_launches" > $O/pytest_chain.txt 2>&1
grep -E "chain\]|passed|failed|^E  |FAILED|Timeout" $O/pytest_chain.txt | tail -n 8
for net in yolov8 ufldv2; do
  ADAS_B200_CHAIN_LOG=1 timeout 600 python tools/op_table.py $net 8 > $O/optable_${net}_auto.txt 2>$O/optable_${net}_auto.err; tail -n 2 $O/optable_${net}_auto.txt
  grep "chain\]" $O/optable_${net}_auto.err
  ADAS_B200_CHAIN=1 timeout 600 python tools/op_table.py $net 8 > $O/optable_${net}_chain.txt 2>$O/optable_${net}_chain.err; tail -n 2 $O/optable_${net}_chain.txt
  ADAS_B200_CHAIN=0 timeout 600 python tools/op_table.py $net 8 > $O/optable_${net}_nochain.txt 2>&1; tail -n 2 $O/optable_${net}_nochain.txt
done
for mode in 2 0 1 2 0; do
ADAS_B200_CHAIN=$mode timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_$mode.json 2>$O/bench_$mode.err; python -c "
import json;d=json.loads(open('$O/bench_$mode.json').read().strip().splitlines()[-1]);print('chain mode $mode',d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['roofline']['frac'],d['tracks_alive'])"
done
