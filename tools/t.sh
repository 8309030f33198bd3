#!/bin/sh
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 600 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['roofline']['frac'])
for a in l.get('also',[]):
    print(a.get('workload'), a.get('samples_per_gpu'), a.get('mode'), a.get('ms_per_step'), a.get('value'), a.get('launches_per_step'), a.get('error'))
PY
