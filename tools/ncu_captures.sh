#!/bin/bash
# one GPU: launch list of the default inference step + full captures of the dominant kernels of the final build
mkdir -p gpurun_out
python bench.py --no-also --steps 2 --warmup 1 > gpurun_out/r2_final_plain.json 2>/dev/null; echo "plain rc=$?"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_infer.csv python bench.py --no-also --steps 2 --warmup 1 > gpurun_out/r2_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:"gru_seq_proj_kernel|mlp_head_tc_kernel|segment_reduce_kernel|gru_cell_tc_kernel|project_kernel" -s 10 -c 5 -o gpurun_out/r2_final_infer -f python bench.py --no-also --steps 2 --warmup 1 > gpurun_out/r2_ncu_full.log 2>&1; echo "full rc=$?"
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:agg_gru_tc_kernel -s 2 -c 1 -o gpurun_out/r2_final_agg -f python tools/mpnn_bench.py --nodes 4000000 --edges 80000000 --steps 3 > gpurun_out/r2_ncu_agg.log 2>&1; echo "agg rc=$?"
ls -la gpurun_out/r2_final_*.ncu-rep
