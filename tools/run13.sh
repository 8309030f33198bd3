#!/bin/bash
cp ignnition_b200/libignnition_b200.so /tmp/lib_real.so
for w in 1 2 3; do
  cp tools/libign_projw_$w.so ignnition_b200/libignnition_b200.so
  echo "walkers per SM: $w"
  timeout -s KILL 120 python tools/ordered_update_bench.py 2>&1 | grep -E "gru_seq_proj walkers|hoisted" | tail -2 | cut -c1-330
done
cp /tmp/lib_real.so ignnition_b200/libignnition_b200.so
