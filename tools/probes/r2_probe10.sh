#!/bin/bash
# round-2 probe 10: A/B in one box -- residual through TMA vs per-lane loads (YOLO), overlapping-row TMA map probe
O=gpurun_out/probe10; mkdir -p $O
./tools/tma_overlap_probe > $O/tma_overlap.txt 2>&1; head -n 40 $O/tma_overlap.txt
python tools/op_table.py yolov8 8 > $O/optable_yolo_b8.txt 2>$O/optable_yolo_b8.err; tail -n 2 $O/optable_yolo_b8.txt
ADAS_B200_NO_RES_TMA=1 python tools/op_table.py yolov8 8 > $O/optable_yolo_b8_nores.txt 2>&1; tail -n 2 $O/optable_yolo_b8_nores.txt
python tools/op_table.py yolov8 8 > $O/optable_yolo_b8_again.txt 2>&1; tail -n 2 $O/optable_yolo_b8_again.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --timeout 240 > $O/pytest_kernels.txt 2>&1; tail -n 2 $O/pytest_kernels.txt
