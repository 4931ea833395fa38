#!/bin/bash
# (gpurun --gpus 2) two engines on two GPUs in one process
mkdir -p gpurun_out/probe29
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q --timeout 600 -k "two_devices" > gpurun_out/probe29/pytest_two_b.txt 2>&1; tail -n 3 gpurun_out/probe29/pytest_two_b.txt
