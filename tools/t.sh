#!/bin/sh
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 300 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['roofline']['frac'], l.get('clocks'))
for a in l.get('also',[]):
    print(a.get('workload'), a.get('samples_per_gpu'), a.get('mode'), a.get('ms_per_step'), a.get('value'), a.get('error'))
PY
