#!/bin/bash
mkdir -p gpurun_out
for upt in 32 16; do
  echo "UPT=$upt"
  IGN_PROJ_UPT=$upt timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gru_seq or write_only" 2>&1 | tail -2
  IGN_PROJ_UPT=$upt timeout -s KILL 300 python tools/ordered_update_bench.py 2>&1 | grep -E "hoisted"
done
