#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
timeout -s KILL 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print('main', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'roof', round(d['roofline']['frac'],3))
for a in d['also'] or []:
    if 'error' in a: print('ERR', a); continue
    if 'ms_per_iteration' in a: print(a['workload'], a['exchange'], round(a['ms_per_iteration'],3), round(a['mp_edges_per_s_per_iteration']/1e9,2)); continue
    print(a['workload'], a['mode'], a['samples_per_gpu'], 'value', round(a['value']), 'ms', round(a['ms_per_step'],4), 'e2e', round(a['e2e']['value']), 'launches/step', a['launches_per_step'], a.get('parity') and (a['parity'].get('graph_replay_equals_eager_forward'), a['parity']['max_rel_err_vs_fp64_oracle']))
PY
