#!/bin/sh
python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py tests/test_gpu_kernels.py -q -m gpu -x -k "attention or readout or unsupported or unbuilt" > gpurun_out/r2_pytest_part.log 2>&1
tail -30 gpurun_out/r2_pytest_part.log
