#!/bin/sh
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu.log
