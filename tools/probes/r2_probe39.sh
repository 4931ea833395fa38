#!/bin/bash
# round-2 probe 39: evidence set of the DEFAULT build (chain launches off): ncu launch list + DRAM traffic of the bench, per-layer tables
O=gpurun_out/probe39; mkdir -p $O
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
CMD="python bench.py --profile-steps 2"
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $O/bench_steps_metrics.csv $CMD > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/traffic_report.py $O/bench_steps_metrics.csv 2 8 $O/bench_traffic.json "ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv $CMD" | head -n 8
for net in yolov8 ufldv2; do for b in 8 32; do
  timeout 600 python tools/op_table.py $net $b > $O/optable_${net}_b$b.txt 2>$O/optable_${net}_b$b.err; tail -n 2 $O/optable_${net}_b$b.txt
done; done
