#!/bin/bash
# round-2 probe 13: DRAM traffic of whole bench steps (ncu, caches not flushed), launch list of the same command, 2 steps
O=gpurun_out/probe13; mkdir -p $O
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
CMD="python bench.py --profile-steps 2"
timeout 1500 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $O/bench_steps_metrics.csv $CMD > $O/ncu_bench.log 2>&1
echo "ncu rc=$?"; tail -n 3 $O/ncu_bench.log
python tools/traffic_report.py $O/bench_steps_metrics.csv 2 8 $O/bench_traffic.json "ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv $CMD" | tee $O/traffic_summary.txt
