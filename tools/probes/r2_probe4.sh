#!/bin/bash
# round-2 probe 4: ncu evidence for the v3 kernel -- per-launch DRAM traffic / tensor-pipe share for a whole YOLOv8l + UFLD pass,
# full captures (with source) of the layer-0 GEMM, a 3x3 256->256 @40x40, a 1x1 2048->512, a stride-2 3x3 and the 256->256 @80x80 head conv
O=gpurun_out/probe4; mkdir -p $O
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,sm__cycles_active.avg,launch__grid_size"
for net in yolov8 ufldv2; do
  ADAS_B200_PROFILE_RANGE=1 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $O/metrics_${net}_b8.csv python tools/profile_target.py $net 8 1 > $O/metrics_${net}.log 2>&1
done
for s in 0 25 27 39 87; do
  ADAS_B200_PROFILE_RANGE=1 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_gemm_v3 -s $s -c 1 -f -o $O/full_yolo_g$s python tools/profile_target.py yolov8 8 1 > $O/full_$s.log 2>&1
done
ls -la $O
