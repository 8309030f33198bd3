#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "agg_gru_cell_tc or csr_rank" > gpurun_out/r2_t1.log 2>&1; echo "t1 rc=$?" >> gpurun_out/r2_t1.log
tail -4 gpurun_out/r2_t1.log
timeout -s KILL 900 python tools/mpnn_bench.py --nodes 4000000 --edges 80000000 --steps 3 > gpurun_out/r2_mpnn_small.json 2> gpurun_out/r2_mpnn_small.err && \
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:agg_gru_tc_kernel -s 4 -c 1 -o gpurun_out/r2_agg_v2 python tools/mpnn_bench.py --nodes 4000000 --edges 80000000 --steps 3 > gpurun_out/r2_ncu_agg.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/r2_mpnn_small.json
tail -3 gpurun_out/r2_ncu_agg.log
