#!/bin/bash
# two GPUs: peer-store rate by method, idle and under HBM load
mkdir -p gpurun_out
timeout -s KILL 300 tools/p2p_rate 320 > gpurun_out/r2_p2p_rate.txt 2>&1; echo "p2p rc=$?"
cat gpurun_out/r2_p2p_rate.txt
