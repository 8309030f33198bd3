#!/bin/bash
# eight GPUs: the default bench line (weak-scaling inference + every `also` leg), config 5 by exchange and chunk count
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29601 bench.py --gpus 8 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench8 rc=$?"
rm -f gpurun_out/r2_mpnn_n8.json
for ch in 4 16; do
IGN_EXCHANGE_CHUNKS=$ch run 2961$((ch % 10)) tools/mpnn_bench.py --exchange copy --steps 5 >> gpurun_out/r2_mpnn_n8.json 2>> gpurun_out/r2_mpnn_n8.err; echo "mpnn8 chunks=$ch rc=$?"
done
run 29621 tools/mpnn_bench.py --exchange nccl --steps 5 >> gpurun_out/r2_mpnn_n8.json 2>> gpurun_out/r2_mpnn_n8.err; echo "mpnn8 nccl rc=$?"
run 29622 tools/mpnn_bench.py --variant local --exchange boundary,copy --steps 5 >> gpurun_out/r2_mpnn_n8.json 2>> gpurun_out/r2_mpnn_n8.err; echo "mpnn8 local rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_mpnn_n8.json'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['workload'], d['exchange'], 'iter ms', round(d['ms_per_iteration'],3), 'G edges/s', round(d['mp_edges_per_s_per_iteration']/1e9,2), 'kernel', round(d['fused_update']['avg_launch_ms'],3), d.get('exchange_detail'), d['state_checksum'])
d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
print('main', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
for a in d['also'] or []:
    if 'error' in a: print('ERR', a); continue
    if 'ms_per_iteration' in a: print(a['workload'], a['exchange'], round(a['ms_per_iteration'],3), round(a['mp_edges_per_s_per_iteration']/1e9,2)); continue
    print(a['workload'], a['mode'], a['samples_per_gpu'], a['scaling'], 'value', round(a['value']), 'ms', round(a['ms_per_step'],4), 'e2e', round(a['e2e']['value']))
PY
tail -3 gpurun_out/r2_bench_n8.err gpurun_out/r2_mpnn_n8.err
