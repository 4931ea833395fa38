#!/bin/bash
# round-2 probe 14 (gpurun --gpus 2): the per-step NCCL gather issued from the library (comm.cu) -- bench at N=2 with and without it
O=gpurun_out/probe14; mkdir -p $O
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 "$@"; }
run > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"
ADAS_B200_NO_GATHER=1 run > $O/bench_n2_nogather.json 2> $O/bench_n2_nogather.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
for f in bench_n1 bench_n2 bench_n2_nogather; do python -c "
import json;d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]);print('$f',d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['roofline']['frac'],d.get('gather'))"; done
tail -n 5 $O/bench_n2.err
