#!/bin/sh
for n in 3 32; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:small_graph_kernel -s 2 -c 1 -o gpurun_out/r2_small_graph_b$n -f python tools/small_graph_eager.py $n > gpurun_out/ncu_small.log 2>&1
tail -1 gpurun_out/ncu_small.log
done
ls -la gpurun_out/*.ncu-rep | tail -3
