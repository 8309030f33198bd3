#!/bin/sh
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 400 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, l['e2e']['value'])
for a in l.get('also',[]):
    print(a.get('workload'), a.get('samples_per_gpu'), a.get('mode'), a.get('ms_per_step'), a.get('value'), a.get('error'))
PY
