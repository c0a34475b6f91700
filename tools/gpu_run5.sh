#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for d in 32 8; do
  for lay in blk soa; do
  PDEIP_LIB=$PWD/pde_inverse_problem_b200/libpdeip_trace.so timeout 300 python tools/tensor_phase_trace.py $d $lay > gpurun_out/trace_${lay}_d$d.txt 2>&1
  done
done
