#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_partition.py -x -q -m gpu -k "agg_gru_cell_tc or one_rank" 2>&1 | tail -3
timeout -s KILL 300 python tools/mpnn_bench.py --nodes 4000000 --edges 80000000 --steps 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('fused kernel ms', round(d['fused_update']['avg_launch_ms'],3), 'GB/s', round(d['fused_update']['achieved_gbs']), 'pair', round(d['unfused_pair']['segment_reduce_ms'],3), round(d['unfused_pair']['gru_cell_ms'],3))
"
IGN_AGG_DBG=1 timeout -s KILL 300 python tools/mpnn_bench.py --nodes 4000000 --edges 80000000 --steps 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('gather-only ms', round(d['fused_update']['avg_launch_ms'],3))
"
