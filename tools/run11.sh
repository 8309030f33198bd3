#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_framework.py -x -q -m gpu > gpurun_out/r2_t_fw.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_t_fw.log
