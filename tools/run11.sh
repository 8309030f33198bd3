#!/bin/bash
mkdir -p gpurun_out
python bench.py --no-also > gpurun_out/r2_bench_proj.json 2> gpurun_out/r2_bench_proj.err; echo rc=$?
IGN_GRU_SEQ_PROJ=0 python bench.py --no-also > gpurun_out/r2_bench_noproj.json 2> gpurun_out/r2_bench_noproj.err; echo rc=$?
python - <<'PY'
import json
for f in ['r2_bench_proj','r2_bench_noproj']:
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), [(k['name'],k['launches_per_step'],round(k['ms_total'],3)) for k in d['kernels'][:4]], d['parity']['max_rel_err_vs_fp64_oracle'], d['parity']['state_max_rel_err_vs_fp64_oracle'])
PY
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
