#!/bin/bash
# round-2 probe 23: evidence set of the final build -- bench line, ncu launch list + DRAM traffic of the bench command itself,
# full captures (source on) of the 3x3 256->256 @40x40 GEMM launch and of a chain launch, per-layer tables at batch 8 and 32
O=gpurun_out/probe23; mkdir -p $O
timeout 900 python bench.py > $O/bench_n1.json 2>$O/bench_n1.err; python -c "
import json;d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]);print('bench',d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['roofline']['frac'],d['roofline']['traffic'])"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
CMD="python bench.py --profile-steps 2"
timeout 1500 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $O/bench_steps_metrics.csv $CMD > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/traffic_report.py $O/bench_steps_metrics.csv 2 8 $O/bench_traffic.json "ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv $CMD" | tee $O/traffic_summary.txt
for net in yolov8 ufldv2; do for b in 8 32; do
  timeout 600 python tools/op_table.py $net $b > $O/optable_${net}_b$b.txt 2>$O/optable_${net}_b$b.err; tail -n 2 $O/optable_${net}_b$b.txt
done; done
# full captures: the 40x40 3x3 256->256 launch (op index from the table) and the first chain launch
ADAS_B200_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_gemm_v3 -s 22 -c 1 -f -o $O/full_gemm_40x40 python tools/profile_target.py yolov8 8 1 > $O/ncu_full_gemm.log 2>&1
ADAS_B200_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_chain_v3 -c 1 -f -o $O/full_chain python tools/profile_target.py yolov8 8 1 > $O/ncu_full_chain.log 2>&1
for r in full_gemm_40x40 full_chain; do ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null; ncu -i $O/$r.ncu-rep --page details > $O/${r}_details.txt 2>/dev/null; done
python - <<'PY'
import csv
for r in ("full_gemm_40x40","full_chain"):
    rows=list(csv.reader(open(f'gpurun_out/probe23/{r}_raw.csv')))
    h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
    d=dict(zip(h,v))
    print(r, d.get('Kernel Name','')[:40], d.get('Grid Size'), 'dur', d.get('gpu__time_duration.sum'), 'tensor%', d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'), 'dram rd', d.get('dram__bytes_read.sum'), 'wr', d.get('dram__bytes_write.sum'))
PY
ls -la $O
