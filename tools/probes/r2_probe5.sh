#!/bin/bash
# round-2 probe 5: v3 + weight prefetch + silu2 + fast division + FC stream + lite + parity tests
O=gpurun_out/probe5; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q --timeout 600 -s > $O/pytest_all.txt 2>&1
echo "all rc=$?" >> $O/pytest_all.txt
grep -E "parity|passed|failed|Error|error|assert" $O/pytest_all.txt | tail -40
python tools/op_table.py yolov8 8 > $O/optable_yolo_b8.txt 2>$O/optable_yolo_b8.err; tail -n 2 $O/optable_yolo_b8.txt
python tools/op_table.py ufldv2 8 > $O/optable_ufld_b8.txt 2>$O/optable_ufld_b8.err; tail -n 2 $O/optable_ufld_b8.txt
ADAS_B200_NO_WPREFETCH=1 python tools/op_table.py yolov8 8 > $O/optable_yolo_b8_nopf.txt 2>&1; tail -n 2 $O/optable_yolo_b8_nopf.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -n 2 $O/smoke.txt
python bench.py --steps 20 --warmup 5 > $O/bench.json 2>$O/bench.err; tail -c 1500 $O/bench.json
