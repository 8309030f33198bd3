#!/bin/sh
timeout 900 python -m pytest tests/test_gpu_partition.py -q -m gpu > gpurun_out/r2_pytest_part.log 2>&1
tail -3 gpurun_out/r2_pytest_part.log
