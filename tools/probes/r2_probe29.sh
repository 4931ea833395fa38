#!/bin/bash
# round-2 probe 29 (gpurun --gpus 2): two engines on two GPUs in one process; N=2 bench of the final build
O=gpurun_out/probe29; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q --timeout 600 -s -k "two_devices" > $O/pytest_two.txt 2>&1; echo "rc=$?"; tail -n 3 $O/pytest_two.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 50 --warmup 5 --cpu-frames 0 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"
python -c "
import json;d=json.loads(open('$O/bench_n2.json').read().strip().splitlines()[-1]);print('n2',d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d.get('gather'))"
