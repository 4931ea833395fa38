#!/bin/bash
# round-2 probe 2: v3 kernel (staged TMA-store epilogue, 16 epilogue warps) -- parity, whole-suite, per-layer tables, sanitizer
O=gpurun_out/probe2; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --timeout 240 > $O/pytest_kernels.txt 2>&1
echo "kernels rc=$?" >> $O/pytest_kernels.txt
tail -n 5 $O/pytest_kernels.txt
if grep -q "failed\|error\|Timeout" $O/pytest_kernels.txt; then
  # show which tile shapes fail, one by one, without -x
  timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "tile_shapes or concat or conv_parity" > $O/pytest_kernels_all.txt 2>&1
  tail -n 40 $O/pytest_kernels_all.txt
fi
python tools/op_table.py yolov8 8 > $O/optable_yolo_b8.txt 2>$O/optable_yolo_b8.err
tail -n 2 $O/optable_yolo_b8.txt
python tools/op_table.py ufldv2 8 > $O/optable_ufld_b8.txt 2>$O/optable_ufld_b8.err
tail -n 2 $O/optable_ufld_b8.txt
ADAS_B200_AT_LOG=1 python tools/op_table.py yolov8 8 > $O/optable_yolo_b8_atlog.txt 2>$O/atlog_yolo_b8.err
ADAS_B200_NO_TMA_STORE=1 python tools/op_table.py yolov8 8 > $O/optable_yolo_b8_nostage.txt 2>&1
tail -n 2 $O/optable_yolo_b8_nostage.txt
python tools/op_table.py yolov8 32 10 > $O/optable_yolo_b32.txt 2>&1
tail -n 2 $O/optable_yolo_b32.txt
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > $O/pytest_all.txt 2>&1
echo "all rc=$?" >> $O/pytest_all.txt
tail -n 8 $O/pytest_all.txt
timeout 600 compute-sanitizer --tool memcheck --log-file $O/memcheck.txt python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tile_shapes and (128-2 or 256-2 or 64-2)" > $O/memcheck_run.txt 2>&1
tail -n 3 $O/memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --log-file $O/racecheck.txt python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tile_shapes and 128-2" > $O/racecheck_run.txt 2>&1
tail -n 3 $O/racecheck.txt
