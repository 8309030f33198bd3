#!/bin/bash
# one GPU: sanitizer logs (memcheck, racecheck, synccheck) of the smoke model and of the new tcgen05 kernels at small sizes
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_${tool}_smoke.log 2>&1; echo "$tool smoke rc=$?"
  tail -n 4 gpurun_out/r2_sanitizer_${tool}_smoke.log
done
for tool in memcheck racecheck; do
  timeout -s KILL 1200 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "test_gru_seq_proj and (127 or 129 or 3000) or test_agg_gru_cell_tc and 129 or test_gather_dense and 1000" > gpurun_out/r2_sanitizer_${tool}_kernels.log 2>&1; echo "$tool kernels rc=$?"
  tail -n 6 gpurun_out/r2_sanitizer_${tool}_kernels.log
done
