#!/bin/bash
# eight GPUs: the default bench line (weak-scaling inference + every `also` leg)
mkdir -p gpurun_out
t0=$(date +%s)
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench8 rc=$? wall $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
print('main', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
for a in d['also'] or []:
    if 'error' in a: print('ERR', a); continue
    if 'ms_per_iteration' in a: print(a['workload'], a['exchange'], round(a['ms_per_iteration'],3), round(a['mp_edges_per_s_per_iteration']/1e9,2), a.get('exchange_detail',{}).get('floor_over_achieved')); continue
    if 'parse_only_samples_per_s' in a: print(a['workload'], round(a['value']), 'parse only', round(a['parse_only_samples_per_s']), a['host_threads'], a['host_cores']); continue
    print(a['workload'], a['mode'], a['samples_per_gpu'], a['scaling'], 'value', round(a['value']), 'ms', round(a['ms_per_step'],4), 'e2e', round(a['e2e']['value']))
PY
tail -n 4 gpurun_out/r2_bench_n8.err
