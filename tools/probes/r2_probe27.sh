#!/bin/bash
# round-2 probe 27 (gpurun --gpus 8): the scaling run the driver does -- bench at N=8 with the per-step gather issued from C
O=gpurun_out/probe27; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 30 --warmup 5 --cpu-frames 0 > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$?"
python -c "
import json;d=json.loads(open('$O/bench_n8.json').read().strip().splitlines()[-1]);print('n8',d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['roofline']['frac'],d.get('gather'))"
tail -n 3 $O/bench_n8.err
