#!/bin/bash
# Accuracy / speed study of the lo weight halves in the tensor residual kernel (profiles/r01_summary_tensor_v2.md).
# Needs the variant libraries next to libpdeip.so, built from csrc/ with
#   nvcc ... -DPDEIP_TC_NLO_FWD=<f> -DPDEIP_TC_NLO_BWD=<b> -c residual_tensor.cu -o rt.o ; nvcc -shared -o ../libpdeip_nlo<f><b>.so build/*.o(rt.o instead of residual_tensor.o)
# (f, b) = streams of the forward (P1, P2) / backward (P3, P4) layer GEMMs that get the lo halves: 32 = all, 11 = the default.
for v in 32 12 11 21; do
  echo "=== NLO fwd/bwd = $v"
  PDEIP_LIB=$PWD/pde_inverse_problem_b200/libpdeip_nlo$v.so timeout 200 python tools/tensor_errors.py 2>&1 | tail -7
  PDEIP_LIB=$PWD/pde_inverse_problem_b200/libpdeip_nlo$v.so timeout 200 python bench.py --particles 606208 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', j['ms_per_step'], j['kernels']['mlp_residual'], j['loss'])"
done
