#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_model.py -x -q -m gpu -k "attention or unbuilt" > gpurun_out/r2_t_attn.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2_t_attn.log
