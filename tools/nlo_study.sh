for v in 32 12 11 21; do
  echo "=== NLO fwd/bwd = $v"
  PDEIP_LIB=$PWD/pde_inverse_problem_b200/libpdeip_nlo$v.so timeout 200 python tools/tensor_errors.py 2>&1 | tail -7
  PDEIP_LIB=$PWD/pde_inverse_problem_b200/libpdeip_nlo$v.so timeout 200 python bench.py --particles 606208 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', j['ms_per_step'], j['kernels']['mlp_residual'], j['loss'])"
done
