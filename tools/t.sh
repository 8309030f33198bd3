#!/bin/sh
python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -q -m gpu -x -k "concat or two_source or interleave or ordered" > gpurun_out/r2_pytest_part.log 2>&1
tail -30 gpurun_out/r2_pytest_part.log
