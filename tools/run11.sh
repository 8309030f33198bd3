#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -m gpu -k "gather_dense or generic_mpnn or dense" > gpurun_out/r2_t_gd.log 2>&1; echo "t rc=$?"
tail -25 gpurun_out/r2_t_gd.log
