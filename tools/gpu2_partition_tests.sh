#!/bin/bash
# two GPUs: partitioned-graph parity tests (all exchanges) + NCCL data-parallel training test, config 5 variant C
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_partition.py tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/r2_t2gpu.log 2>&1; echo "t rc=$?" >> gpurun_out/r2_t2gpu.log
tail -4 gpurun_out/r2_t2gpu.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/mpnn_bench.py --variant local --exchange boundary,copy --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['workload'], d['exchange'], 'iter ms', round(d['ms_per_iteration'],3), 'G edges/s', round(d['mp_edges_per_s_per_iteration']/1e9,2), d['exchange_detail']['bytes_received_per_gpu_per_iteration'], d['state_checksum'])
"
