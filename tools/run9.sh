#!/bin/bash
mkdir -p gpurun_out
python tools/parity_probe.py > gpurun_out/r2_parity_probe2.jsonl 2> gpurun_out/r2_parity_probe.err
python tools/parity_probe.py --workload qsize_nsfnet_b4096 >> gpurun_out/r2_parity_probe2.jsonl 2>> gpurun_out/r2_parity_probe.err
tail -3 gpurun_out/r2_parity_probe.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_parity_probe2.jsonl'):
    d=json.loads(l); print(d['workload'], d['weights'])
    for k,v in d['max_rel_dev_vs_fp64'].items(): print('   %-22s'%k, {a:'%.2e'%b for a,b in v.items()})
PY
