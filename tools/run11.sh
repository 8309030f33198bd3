#!/bin/bash
for thr in 0 100000000; do for b in 256 1024 2048 4096; do IGN_BWD_STEPS_MIN_ROWS=$thr python bench.py --workload routenet_nsfnet_b4096 --batch $b --train --no-also --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('thr $thr train batch', d['config']['samples_per_gpu'], 'paths', d['config']['paths_per_gpu'], round(d['ms_per_step'],3), 'ms', [(k['name'],k['launches_per_step'],round(k['ms_total'],3)) for k in d['kernels'][:2]])"; done; done
