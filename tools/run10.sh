#!/bin/bash
# eight GPUs: copy exchange by chunk count and number of copy streams (one graph build)
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 tools/mpnn_bench.py --steps 8 --sweep 4x1x1.7,4x1x1.4,4x1x2.0,3x1x1.7,3x1x2.2,5x1x1.5,4x2x1.7,4x1x1.7 > gpurun_out/r2_mpnn_n8_sweep.json 2> gpurun_out/r2_mpnn_n8_sweep.err; echo "rc=$?"
cat gpurun_out/r2_mpnn_n8_sweep.json | grep sweep
tail -n 3 gpurun_out/r2_mpnn_n8_sweep.err
