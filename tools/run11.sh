#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_graphs.py tests/test_gpu_framework.py -x -q -m gpu > gpurun_out/r2_t_train.log 2>&1; echo "t rc=$?"
tail -25 gpurun_out/r2_t_train.log
