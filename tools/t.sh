#!/bin/sh
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_model.py -q -m gpu -x -k "reset_after or unsupported or generic_width or unbuilt" > gpurun_out/r2_pytest_part.log 2>&1
tail -25 gpurun_out/r2_pytest_part.log
