#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 10 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench4 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n4.json').read().strip().splitlines()[-1])
print('main', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
for a in d['also'] or []:
    if 'error' in a: print('ERR', a); continue
    if 'ms_per_iteration' in a: print(a['workload'], a['exchange'], round(a['ms_per_iteration'],3), round(a['mp_edges_per_s_per_iteration']/1e9,2)); continue
    if 'parse_only_samples_per_s' in a: print(a['workload'], round(a['value'])); continue
    print(a['workload'], a['mode'], a['samples_per_gpu'], a['scaling'], 'value', round(a['value']), 'ms', round(a['ms_per_step'],4))
PY
tail -n 3 gpurun_out/r2_bench_n4.err
