#!/bin/bash
# round-2 probe 1: per-layer tables (product / decomposition bits), untested epilogue variants
O=gpurun_out/probe1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
python tools/op_table.py yolov8 8 > $O/optable_yolo_b8.txt 2>$O/optable_yolo_b8.err
python tools/op_table.py ufldv2 8 > $O/optable_ufld_b8.txt 2>$O/optable_ufld_b8.err
python tools/op_table.py yolov8 32 10 > $O/optable_yolo_b32.txt 2>$O/optable_yolo_b32.err
for d in 0 16 48 18 64; do
  ADAS_B200_AUTOTUNE=0 ADAS_B200_DBG=$d python tools/op_table.py yolov8 8 > $O/optable_yolo_b8_noat_dbg$d.txt 2>&1
done
ADAS_B200_HOIST=1 python tools/op_table.py yolov8 8 > $O/optable_yolo_b8_hoist.txt 2>&1
ADAS_B200_EPI16=1 python tools/op_table.py yolov8 8 > $O/optable_yolo_b8_epi16.txt 2>&1
ADAS_B200_MC=3 ADAS_B200_AT_LOG=1 python tools/op_table.py yolov8 8 > $O/optable_yolo_b8_mc3.txt 2>$O/optable_yolo_b8_mc3.err
ADAS_B200_EPI16=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "conv_parity or stem or fc" > $O/pytest_epi16.txt 2>&1
ADAS_B200_HOIST=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "conv_parity or stem or fc" > $O/pytest_hoist.txt 2>&1
tail -3 $O/optable_yolo_b8.txt $O/optable_ufld_b8.txt $O/pytest_epi16.txt $O/pytest_hoist.txt
