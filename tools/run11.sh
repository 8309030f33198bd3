#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout -s KILL 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"
tail -n 3 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print('main', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'roof', round(d['roofline']['frac'],3), d['roofline']['traffic'])
print(d['parity'])
for a in d['also'] or []:
    if 'error' in a: print('ERR', a); continue
    if 'ms_per_iteration' in a: print(a['workload'], a['exchange'], round(a['ms_per_iteration'],3), round(a['mp_edges_per_s_per_iteration']/1e9,2)); continue
    if 'parse_only_samples_per_s' in a: print(a['workload'], round(a['value']), 'parse only', round(a['parse_only_samples_per_s']), a['host_threads'], a['host_cores']); continue
    print(a['workload'], a['mode'], a['samples_per_gpu'], 'value', round(a['value']), 'ms', round(a['ms_per_step'],4), 'e2e', round(a['e2e']['value']), 'launches/step', a['launches_per_step'])
PY
