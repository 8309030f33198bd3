#!/bin/bash
mkdir -p gpurun_out
cp ignnition_b200/libignnition_b200.so /tmp/lib_real.so
for nb in 3 5 7 9 12; do
  cp tools/libign_aggprobe_$nb.so ignnition_b200/libignnition_b200.so
  IGN_AGG_DBG=1 timeout -s KILL 300 python tools/mpnn_bench.py --nodes 4000000 --edges 80000000 --steps 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('NB=$nb gather-only kernel ms', round(d['fused_update']['avg_launch_ms'],3), 'GB/s', round(d['fused_update']['achieved_gbs']), 'segment_reduce ms', round(d['unfused_pair']['segment_reduce_ms'],3))
"
done
cp /tmp/lib_real.so ignnition_b200/libignnition_b200.so
