#!/bin/bash
# round-2 probe 19: tracker time breakdown (library vs association round trips), ncu full capture of the stem conv kernel
O=gpurun_out/probe19; mkdir -p $O
timeout 600 python bench.py --steps 100 --warmup 5 --cpu-frames 0 --other-configs 0 > $O/bench.json 2>$O/bench.err; python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('bench',d['value'],d['ms_per_step'],d['e2e']['value'],d['host_tracker_ms_per_step'],d['host_tracker_breakdown_ms_per_step'],d['roofline']['frac'])"
ADAS_B200_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:stem_conv -c 1 -f -o $O/full_stem python tools/profile_target.py yolov8 8 1 > $O/ncu_stem.log 2>&1
ncu -i $O/full_stem.ncu-rep --page raw --csv > $O/full_stem_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/probe19/full_stem_raw.csv')))
h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
want=["gpu__time_duration.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","dram__throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_tensor","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct","smsp__average_warp_latency_issue_stalled_long_scoreboard","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__inst_executed.sum","dram__bytes_read.sum","dram__bytes_write.sum","launch__occupancy_limit_registers","launch__registers_per_thread","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio","smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_drain_per_issue_active.ratio","smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio"]
for i,n in enumerate(h):
    if any(n.startswith(w) for w in want): print(n, '=', v[i], rows[1][i] if len(rows)>2 else '')
PY
