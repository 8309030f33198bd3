#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "agg_gru_cell_tc or csr_rank" > gpurun_out/r2_t1.log 2>&1; echo "t1 rc=$?" >> gpurun_out/r2_t1.log
tail -4 gpurun_out/r2_t1.log
timeout -s KILL 600 python -m pytest tests/test_gpu_partition.py -x -q -m gpu > gpurun_out/r2_t2.log 2>&1; echo "t2 rc=$?" >> gpurun_out/r2_t2.log
tail -4 gpurun_out/r2_t2.log
timeout -s KILL 900 python tools/mpnn_bench.py --variant uniform,skew > gpurun_out/r2_mpnn1.json 2> gpurun_out/r2_mpnn1.err; echo "mpnn rc=$?"
tail -3 gpurun_out/r2_mpnn1.err; cat gpurun_out/r2_mpnn1.json
timeout -s KILL 600 python bench.py --steps 10 --no-also --cpu-samples 4 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_a.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])
for k in d['kernels']: print(k['name'], k['launches_per_step'], round(k['ms_total'],3), round(k['ms_avg'],4))
PY
IGN_NO_FUSED_TC=1 timeout -s KILL 600 python bench.py --steps 10 --no-also --cpu-samples 4 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_b.json').read().strip().splitlines()[-1])
print('unfused', d['value'], d['ms_per_step'])
for k in d['kernels']: print(k['name'], k['launches_per_step'], round(k['ms_total'],3), round(k['ms_avg'],4))
PY
