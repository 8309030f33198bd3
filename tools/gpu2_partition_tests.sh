#!/bin/bash
# two GPUs: partitioned-graph parity tests (all exchanges) + NCCL data-parallel training test, config 5 with the copy exchange
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_partition.py tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/r2_t2gpu.log 2>&1; echo "t rc=$?" >> gpurun_out/r2_t2gpu.log
tail -4 gpurun_out/r2_t2gpu.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/mpnn_bench.py --steps 8 --sweep 4x1x0 > gpurun_out/r2_mpnn_n2_sweep.json 2> gpurun_out/r2_mpnn_n2_sweep.err; echo "rc=$?"
grep sweep gpurun_out/r2_mpnn_n2_sweep.json
