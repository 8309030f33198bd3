#!/bin/bash
# two GPUs: partitioned-graph parity tests, NCCL data-parallel training test, config 5 with the three exchanges
mkdir -p gpurun_out
nvidia-smi -L
timeout -s KILL 900 python -m pytest tests/test_gpu_partition.py tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/r2_t2gpu.log 2>&1; echo "t rc=$?" >> gpurun_out/r2_t2gpu.log
tail -6 gpurun_out/r2_t2gpu.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/mpnn_bench.py --exchange peer,nccl --steps 5 > gpurun_out/r2_mpnn_n2.json 2> gpurun_out/r2_mpnn_n2.err; echo "mpnn2 rc=$?"
tail -3 gpurun_out/r2_mpnn_n2.err
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/mpnn_bench.py --variant local --exchange boundary,peer --steps 5 >> gpurun_out/r2_mpnn_n2.json 2>> gpurun_out/r2_mpnn_n2.err; echo "mpnn2b rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_mpnn_n2.json'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['workload'], d['exchange'], 'iter ms', round(d['ms_per_iteration'],3), 'G edges/s', round(d['mp_edges_per_s_per_iteration']/1e9,2), 'kernel', round(d['fused_update']['avg_launch_ms'],3), d.get('exchange_detail'), d['state_checksum'])
PY
