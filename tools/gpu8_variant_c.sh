#!/bin/bash
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 tools/mpnn_bench.py --variant local --exchange boundary,copy --steps 8 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['workload'], d['exchange'], 'iter ms', round(d['ms_per_iteration'],3), 'G edges/s', round(d['mp_edges_per_s_per_iteration']/1e9,2), 'kernel', round(d['fused_update']['avg_launch_ms'],3), d['exchange_detail']['bytes_received_per_gpu_per_iteration'], d['state_checksum'])
"
