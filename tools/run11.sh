#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gru_seq or write_only" > gpurun_out/r2_t_proj.log 2>&1; echo "t rc=$?"
tail -3 gpurun_out/r2_t_proj.log
timeout -s KILL 300 python tools/ordered_update_bench.py 2>&1 | grep -E "hoisted|steps 6"
python bench.py --no-also > gpurun_out/r2_bench_proj.json 2> gpurun_out/r2_bench_proj.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_proj.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), [(k['name'],k['launches_per_step'],round(k['ms_total'],3)) for k in d['kernels'][:5]])
PY
