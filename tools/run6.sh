#!/bin/bash
# two GPUs: partitioned-graph parity tests (all exchanges), config 5 with the copy-engine exchange by chunk count
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_partition.py -x -q -m gpu > gpurun_out/r2_t2gpu.log 2>&1; echo "t rc=$?" >> gpurun_out/r2_t2gpu.log
tail -4 gpurun_out/r2_t2gpu.log
rm -f gpurun_out/r2_mpnn_n2.json
for ch in 4 8 16; do
IGN_EXCHANGE_CHUNKS=$ch timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$ch tools/mpnn_bench.py --exchange copy --steps 5 >> gpurun_out/r2_mpnn_n2.json 2>> gpurun_out/r2_mpnn_n2.err; echo "mpnn2 chunks=$ch rc=$?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r2_mpnn_n2.json'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['workload'], d['exchange'], 'iter ms', round(d['ms_per_iteration'],3), 'G edges/s', round(d['mp_edges_per_s_per_iteration']/1e9,2), 'kernel', round(d['fused_update']['avg_launch_ms'],3), d.get('exchange_detail'), d['state_checksum'])
PY
