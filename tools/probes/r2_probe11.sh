#!/bin/bash
# round-2 probe 11: chain launches (gemm_chain.cu) -- correctness vs per-layer launches, then A/B timing in the same box
O=gpurun_out/probe11; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q --timeout 300 -s -k "chain_launches" > $O/pytest_chain.txt 2>&1
echo "chain rc=$?" >> $O/pytest_chain.txt
grep -E "chain\]|passed|failed|^E  |FAILED|Timeout|rc=" $O/pytest_chain.txt | tail -20
if grep -q "passed" $O/pytest_chain.txt && ! grep -q "failed" $O/pytest_chain.txt; then
  for net in yolov8 ufldv2; do
    timeout 600 python tools/op_table.py $net 8 > $O/optable_${net}_chain.txt 2>$O/optable_${net}_chain.err; tail -n 2 $O/optable_${net}_chain.txt
    ADAS_B200_CHAIN=0 timeout 600 python tools/op_table.py $net 8 > $O/optable_${net}_nochain.txt 2>&1; tail -n 2 $O/optable_${net}_nochain.txt
  done
  timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_all.txt 2>&1; tail -n 3 $O/pytest_all.txt
  timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2>$O/bench.err; python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['roofline']['frac'],d['tracks_alive'])"
  ADAS_B200_CHAIN=0 timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_nochain.json 2>$O/bench_nochain.err; python -c "
import json;d=json.loads(open('$O/bench_nochain.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['roofline']['frac'],d['tracks_alive'])"
fi
