#!/bin/sh
# Builds tools/libignnition_b200_prof.so: the library with the phase-profiling variants of the
# tensor-core walker (-DIGN_WALK_PROFILE), the fused readout (-DIGN_MLP_PROFILE), the GRU cell (-DIGN_CELL_PROFILE)
# the step-synchronous backward kernel (-DIGN_BWD_PROFILE; read with tools/ordered_bwd_bench.py) and the one-launch
# loop of small graphs (-DIGN_SG_PROFILE: per-stage timestamps of CTA 0 behind the barrier counter; tools/small_graph_stages.py).
# Use on the GPU box:  cp tools/libignnition_b200_prof.so ignnition_b200/libignnition_b200.so
set -e
cd "$(dirname "$0")/../ignnition_b200/csrc"
make -s
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr -I../../include"
nvcc $FLAGS -DIGN_WALK_PROFILE -c gru_seq_tc.cu -o /tmp/ign_gru_seq_tc_prof.o
nvcc $FLAGS -DIGN_PROJ_PROFILE -c gru_seq_proj_tc.cu -o /tmp/ign_gru_seq_proj_tc_prof.o
nvcc $FLAGS -DIGN_MLP_PROFILE -c mlp_head_tc.cu -o /tmp/ign_mlp_head_tc_prof.o
nvcc $FLAGS -DIGN_CELL_PROFILE -c gru_cell_tc.cu -o /tmp/ign_gru_cell_tc_prof.o
nvcc $FLAGS -DIGN_BWD_PROFILE -c gru_step_bwd_tc.cu -o /tmp/ign_gru_step_bwd_tc_prof.o
nvcc $FLAGS -DIGN_SG_PROFILE -c small_graph.cu -o /tmp/ign_small_graph_prof.o
OBJS=$(ls *.o | grep -v -e '^gru_seq_tc.o$' -e '^gru_seq_proj_tc.o$' -e '^mlp_head_tc.o$' -e '^gru_cell_tc.o$' -e '^gru_step_bwd_tc.o$' -e '^small_graph.o$')
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../tools/libignnition_b200_prof.so $OBJS \
  /tmp/ign_gru_seq_tc_prof.o /tmp/ign_gru_seq_proj_tc_prof.o /tmp/ign_mlp_head_tc_prof.o /tmp/ign_gru_cell_tc_prof.o /tmp/ign_gru_step_bwd_tc_prof.o /tmp/ign_small_graph_prof.o -lcudart
echo built tools/libignnition_b200_prof.so
