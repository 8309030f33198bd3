#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_partition.py -x -q -m gpu -k "agg_gru_cell_tc or csr_rank or partition or routing" > gpurun_out/r2_t1.log 2>&1; echo "t1 rc=$?" >> gpurun_out/r2_t1.log
tail -3 gpurun_out/r2_t1.log
for dbg in 0 1; do
IGN_AGG_DBG=$dbg timeout -s KILL 900 python tools/mpnn_bench.py --nodes 4000000 --edges 80000000 --steps 3 > gpurun_out/r2_mpnn_dbg$dbg.json 2> gpurun_out/r2_mpnn_dbg$dbg.err
python - <<PY
import json
for l in open('gpurun_out/r2_mpnn_dbg$dbg.json'):
    d=json.loads(l); print('dbg$dbg', d['workload'], 'fused', round(d['fused_update']['avg_launch_ms'],3), round(d['fused_update']['achieved_gbs']), 'pair', round(d['unfused_pair']['segment_reduce_ms'],3), round(d['unfused_pair']['gru_cell_ms'],3))
PY
done
timeout -s KILL 900 python tools/mpnn_bench.py --steps 5 > gpurun_out/r2_mpnn_v3.json 2> gpurun_out/r2_mpnn_v3.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_mpnn_v3.json'):
    d=json.loads(l); print(d['workload'], 'iter ms', round(d['ms_per_iteration'],3), 'fused', round(d['fused_update']['avg_launch_ms'],3), round(d['fused_update']['achieved_gbs']), 'pair', round(d['unfused_pair']['segment_reduce_ms'],3), round(d['unfused_pair']['gru_cell_ms'],3))
PY
timeout -s KILL 600 python bench.py --steps 10 --no-also --cpu-samples 4 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_a.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])
for k in d['kernels']: print(k['name'], k['launches_per_step'], round(k['ms_total'],3), round(k['ms_avg'],4))
PY
timeout -s KILL 900 python tools/mpnn_bench.py --nodes 4000000 --edges 80000000 --steps 3 > /dev/null 2>&1 && \
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:agg_gru_tc_kernel -s 4 -c 1 -o gpurun_out/r2_agg_v5 python tools/mpnn_bench.py --nodes 4000000 --edges 80000000 --steps 3 > gpurun_out/r2_ncu_agg.log 2>&1
echo "ncu rc=$?"
