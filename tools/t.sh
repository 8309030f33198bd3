#!/bin/sh
timeout -s KILL 70 python -m pytest tests/test_gpu_dist.py -q -m gpu -k nccl_world2 > gpurun_out/r2_t2gpu_b.log 2>&1; echo "t rc=$?" >> gpurun_out/r2_t2gpu_b.log
tail -4 gpurun_out/r2_t2gpu_b.log
