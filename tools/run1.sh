#!/bin/bash
# first GPU pass of round 2: the new fused kernel's tests, then the whole suite, then config 5 on one GPU
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.txt 2>&1
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "agg_gru_cell_tc or csr_rank" > gpurun_out/r2_t1.log 2>&1; echo "t1 rc=$?" >> gpurun_out/r2_t1.log
tail -5 gpurun_out/r2_t1.log
timeout -s KILL 600 python -m pytest tests/test_gpu_partition.py -x -q -m gpu > gpurun_out/r2_t2.log 2>&1; echo "t2 rc=$?" >> gpurun_out/r2_t2.log
tail -5 gpurun_out/r2_t2.log
timeout -s KILL 900 python tools/mpnn_bench.py > gpurun_out/r2_mpnn1.json 2> gpurun_out/r2_mpnn1.err; echo "mpnn rc=$?"
tail -3 gpurun_out/r2_mpnn1.err; cat gpurun_out/r2_mpnn1.json
timeout -s KILL 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t3.log 2>&1; echo "t3 rc=$?" >> gpurun_out/r2_t3.log
tail -5 gpurun_out/r2_t3.log
