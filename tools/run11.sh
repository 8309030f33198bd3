#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "match_autograd" 2>&1 | tail -4
python bench.py --workload routenet_synth50_b256 --train --no-also --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('synth50 train', d['config'], round(d['ms_per_step'],3), round(d['value']))"
