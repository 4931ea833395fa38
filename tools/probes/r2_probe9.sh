#!/bin/bash
# round-2 probe 9: residual tiles through TMA + three staging buffers
O=gpurun_out/probe9; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --timeout 240 > $O/pytest_kernels.txt 2>&1; tail -n 3 $O/pytest_kernels.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 -s > $O/pytest_all.txt 2>&1
grep -E "parity\]|passed|failed|^E  |FAILED" $O/pytest_all.txt | tail -30
python tools/op_table.py yolov8 8 > $O/optable_yolo_b8.txt 2>$O/optable_yolo_b8.err; tail -n 2 $O/optable_yolo_b8.txt
python tools/op_table.py ufldv2 8 > $O/optable_ufld_b8.txt 2>$O/optable_ufld_b8.err; tail -n 2 $O/optable_ufld_b8.txt
ADAS_B200_NO_RES_TMA=1 python tools/op_table.py ufldv2 8 > $O/optable_ufld_b8_nores.txt 2>&1; tail -n 2 $O/optable_ufld_b8_nores.txt
