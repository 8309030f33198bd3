#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gru_seq or write_only" > gpurun_out/r2_t_proj.log 2>&1; echo "t rc=$?"
tail -3 gpurun_out/r2_t_proj.log
timeout -s KILL 300 python tools/ordered_update_bench.py 2>&1 | grep -E "hoisted|steps 6"
cp tools/libignnition_b200_prof.so ignnition_b200/libignnition_b200.so
timeout -s KILL 300 python tools/ordered_update_bench.py > gpurun_out/r2_proj_phases.txt 2>&1
grep -E "gru_seq_proj walkers" gpurun_out/r2_proj_phases.txt | tail -1
